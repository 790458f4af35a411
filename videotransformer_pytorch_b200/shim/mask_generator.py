"""`import mask_generator` of the reference (data_trainer.py) -> bit-identical generators."""
from videotransformer_pytorch_b200.mask_generator import CubeMaskGenerator, RandomMaskGenerator  # noqa: F401
