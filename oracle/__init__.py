"""CPU oracle for the video-transformer hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package
(`videotransformer_pytorch_b200/`) may import from here; only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s CPU-baseline legs do.

Contents
--------
* ``vt_oracle``   – plain-torch (CPU, fp32/fp64) restatement of the reference's
  TimeSformer / ViViT forward (autograd supplies the backward), written without
  einops and driven by a reference-format ``state_dict``.
  Pinned: ``oracle/make_golden.py`` imports the real modules from
  ``/root/reference`` in the build container and checks/commits golden vectors
  under ``tests/golden``.
* ``hog_oracle``  – numpy fp64 restatement of ``skimage.feature.hog`` exactly as
  called at reference ``dataset.py:39-45``.  scikit-image is not installed in
  this image and the reference holds no HOG test vectors: **parity unpinned**.
* ``mask_oracle`` – restatement of ``mask_generator.py:23-107`` (CubeMaskGenerator)
  pinned bit-exactly against the reference class.
"""
