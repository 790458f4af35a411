"""The reference's UNMODIFIED model_trainer.py imported with the shim directory first on sys.path: its
`from transformer import ...` / `from video_transformer import ...` / `from mixup import Mixup` resolve to this package,
`VideoTransformer.__init__` (model_trainer.py:40-104) builds the B200 modules, and a forward runs (kernel table = CPU
emulation).  Build-container only: needs /root/reference; its absent third-party imports (pytorch_lightning, torchmetrics,
timm, matplotlib) are stubbed — none of them is on the hot path."""
import importlib
import os
import sys
import types

import pytest
import torch

REF = '/root/reference'
SHIM = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'videotransformer_pytorch_b200', 'shim')


@pytest.fixture
def reference_trainer():
    if not os.path.isdir(REF):
        pytest.skip('/root/reference not present (build container only)')
    saved_path = list(sys.path)
    saved_mods = {k: sys.modules.get(k) for k in ('transformer', 'video_transformer', 'mixup', 'mask_generator', 'model_trainer',
                                                   'utils', 'optimizer', 'weight_init')}

    def stub(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    class Accuracy:
        def __init__(self, *a, **k):
            pass

    pl = stub('pytorch_lightning', LightningModule=torch.nn.Module)
    stub('pytorch_lightning.utilities')
    stub('pytorch_lightning.utilities.distributed', rank_zero_only=lambda f: f)
    stub('torchmetrics', Accuracy=Accuracy)
    stub('timm'); stub('timm.loss', SoftTargetCrossEntropy=torch.nn.CrossEntropyLoss)
    stub('matplotlib'); stub('matplotlib.pyplot')
    for k in saved_mods:
        sys.modules.pop(k, None)
    sys.path[:0] = [SHIM, REF]
    try:
        yield importlib.import_module('model_trainer')
    finally:
        sys.path[:] = saved_path
        for k, v in saved_mods.items():
            sys.modules.pop(k, None)
            if v is not None:
                sys.modules[k] = v


def test_reference_trainer_builds_b200_modules_through_the_shim(reference_trainer, emu):
    import videotransformer_pytorch_b200 as pkg
    mt = reference_trainer
    assert mt.TimeSformer is pkg.TimeSformer and mt.ViViT is pkg.ViViT and mt.MaskFeat is pkg.MaskFeat
    assert mt.ClassificationHead is pkg.ClassificationHead and mt.Mixup is pkg.Mixup
    assert 'reference' in mt.__file__                      # the trainer itself is the reference's file, unmodified
    cfg = types.SimpleNamespace(objective='supervised', arch='timesformer', pretrain_pth=None, weights_from='imagenet',
                                img_size=32, num_frames=2, attention_type='divided_space_time', num_class=5,
                                eval_metrics='finetune', mixup=False)
    vt = mt.VideoTransformer(cfg, trainer=None, ckpt_dir='.', do_eval=False, do_test=False)
    assert isinstance(vt.model, pkg.TimeSformer) and isinstance(vt.cls_head, pkg.ClassificationHead)
    assert vt.no_weight_decay_keywords() == vt.model.no_weight_decay_keywords()
    # the reference's training_step core (model_trainer.py:204-208) on a tiny clip; embed_dims 768 is fixed by the ctor
    x = torch.randn(2, 2, 3, 32, 32)
    y = torch.tensor([1, 3])
    preds = vt.cls_head(vt.model(x))
    loss = vt.loss_fn(preds, y)
    loss.backward()
    assert preds.shape == (2, 5) and torch.isfinite(loss)
    assert all(p.grad is not None for p in vt.model.parameters())
    # the reference's own optimizer grouping walks named_parameters() of these modules (optimizer.py:49)
    names = [n for n, _ in vt.named_parameters()]
    assert 'model.transformer_layers.layers.0.attentions.0.temporal_fc.weight' in names and 'cls_head.cls_head.weight' in names
