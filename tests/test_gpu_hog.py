"""HOG kernel vs the numpy fp64 oracle: bins bit-exact, values to fp32 tolerance.  -m gpu"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_hog_matches_oracle():
    from oracle import hog_oracle as HO
    from videotransformer_pytorch_b200 import hog
    rng = np.random.default_rng(0)
    frames = rng.integers(0, 256, size=(3, 224, 224, 3), dtype=np.uint8)
    frames[1, :, :, :] = (np.linspace(0, 255, 224)[None, :, None] * np.ones((224, 1, 3))).astype(np.uint8)  # ramps
    frames[2, 50:100, 60:120] = 255
    feat, bins = hog.hog_features(torch.from_numpy(frames).cuda(), want_bins=True)
    for f in range(3):
        assert np.array_equal(bins[f].cpu().numpy(), HO.extract_hog_bins(frames[f])), f
        ref = HO.extract_hog_features(frames[f])
        got = feat[f].cpu().numpy().astype(np.float64)
        assert np.abs(got - ref).max() < 2e-5, (f, np.abs(got - ref).max())
