"""Timing of the TimeSformer-B attention variants (space_only, joint_space_time) and of the fused clip + optimizer step on
one B200.  fwd+bwd through graph.GraphedTrainStep, CUDA events.

    python tools/variants_bench.py [--batch 8] [--steps 5]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from videotransformer_pytorch_b200 import ClassificationHead, TimeSformer
from videotransformer_pytorch_b200.graph import GraphedTrainStep
from videotransformer_pytorch_b200.optim import FusedAdamW, FusedSGD

ap = argparse.ArgumentParser()
ap.add_argument('--batch', type=int, default=8)
ap.add_argument('--steps', type=int, default=5)
a = ap.parse_args()
dev = torch.device('cuda')


class Net(torch.nn.Module):
    def __init__(self, attention_type):
        super().__init__()
        self.m = TimeSformer(num_frames=8, img_size=224, patch_size=16, embed_dims=768, num_heads=12,
                             num_transformer_layers=12, attention_type=attention_type)
        self.h = ClassificationHead(400, 768)

    def forward(self, x, y):
        return torch.nn.functional.cross_entropy(self.h(self.m(x)), y)


def timed(fn, n):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


torch.manual_seed(0)
x = torch.randn(a.batch, 8, 3, 224, 224, device=dev)
y = torch.randint(0, 400, (a.batch,), device=dev)
for kind in ('joint_space_time', 'space_only', 'divided_space_time'):
    net = Net(kind).to(dev).train()
    step = GraphedTrainStep(net, (x, y))
    for _ in range(2):
        loss = step(x, y)
    ms = timed(lambda: step(x, y), a.steps)
    print(f'TimeSformer-B {kind:19s} 8x224 batch {a.batch} fwd+bwd (graph): {ms:7.2f} ms/step = {a.batch / ms * 1e3:6.1f} clips/s; '
          f'loss {float(loss):.4f}; {step.kernels_per_replay} kernels per replay')
    if kind == 'divided_space_time':
        params = [p for p in net.parameters()]
        for name, opt in (('SGD-nesterov', FusedSGD(params, lr=1e-3, momentum=0.9, nesterov=True, weight_decay=1e-4)),
                          ('AdamW', FusedAdamW(params, lr=1e-4, weight_decay=0.05))):
            opt.step(clip_grad=1.0)
            ms_o = timed(lambda: opt.step(clip_grad=1.0), 10)
            n = sum(p.numel() for p in params)
            print(f'  fused clip + {name} over {n / 1e6:.1f} M parameters ({len(params)} tensors): {ms_o * 1e3:.0f} us/step')
        ref = torch.optim.SGD(params, lr=1e-3, momentum=0.9, nesterov=True, weight_decay=1e-4)

        def ref_step():
            for p in params:                      # clip_gradients as the reference writes it (model_trainer.py:155-170)
                nrm = torch.norm(p.grad.detach(), 2)
                coef = 1.0 / (nrm + 1e-6)
                if coef < 1:
                    p.grad.data.mul_(coef)
            ref.step()
        ref_step()
        print(f'  reference flow (per-parameter torch.norm + host compare, then torch.optim.SGD): {timed(ref_step, 3) * 1e3:.0f} us/step')
    del step, net
    torch.cuda.empty_cache()
