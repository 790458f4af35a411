"""`from mixup import Mixup` of the reference (model_trainer.py:15) -> RNG-stream identical Mixup that can also defer the
blend of uint8 clips to the patch-operand kernel."""
from videotransformer_pytorch_b200.mixup import (MixedClip, Mixup, cutmix_bbox_and_lam, mixup_target, one_hot,  # noqa: F401
                                                 rand_bbox)
