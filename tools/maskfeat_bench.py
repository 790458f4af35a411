"""MaskFeat (MViT-B, reference 2-stage Q-pool config) 16x224^2 fwd+bwd timing on one B200 (BASELINE config 5 shape):
full MaskFeat.forward with cube masks from CubeMaskGenerator and HOG targets from the HOG kernel.  Eager launches,
CUDA-event timing, plus a CUPTI kernel-time breakdown of one step.

    python tools/maskfeat_bench.py [--batch 8] [--steps 5] [--profile]
"""
import argparse
import os
import random
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from videotransformer_pytorch_b200 import MaskFeat, _lib
from videotransformer_pytorch_b200.hog import hog_targets
from videotransformer_pytorch_b200.mask_generator import CubeMaskGenerator

ap = argparse.ArgumentParser()
ap.add_argument('--batch', type=int, default=8)
ap.add_argument('--steps', type=int, default=5)
ap.add_argument('--warmup', type=int, default=2)
ap.add_argument('--profile', action='store_true')
ap.add_argument('--graph', action='store_true', help='also time the step replayed as one CUDA graph')
a = ap.parse_args()
dev = torch.device('cuda')
torch.manual_seed(0)
random.seed(0)
B = a.batch
model = MaskFeat(pool_q_stride_size=[[1, 1, 2, 2], [3, 1, 2, 2]], feature_dim=2 * 2 * 2 * 3 * 9).to(dev).train()
gen = CubeMaskGenerator((8, 14, 14), min_num_patches=16)
x = torch.randn(B, 16, 3, 224, 224, device=dev)
u8 = torch.randint(0, 256, (B, 16, 224, 224, 3), dtype=torch.uint8, device=dev)
masks, markers = [], []
for _ in range(B):
    m, cm = gen()
    masks.append(torch.as_tensor(m))
    markers.append(cm)
mask = torch.stack(masks).to(dev)
target = torch.stack([hog_targets(u8[i], markers[i]) for i in range(B)])


def step():
    for p in model.parameters():
        p.grad = None
    pred, loss = model(x, target, mask, markers)
    loss.backward()
    return loss


# the captured step goes first: autograd's AccumulateGrad nodes are bound to the stream they are first used on, and an
# eager step on the default stream would pin them there (graph.py)
if a.graph:
    from videotransformer_pytorch_b200.graph import GraphedTrainStep

    class Step(torch.nn.Module):
        def __init__(self, m):
            super().__init__()
            self.m = m

        def forward(self, x, target, mask, cmask):
            return self.m.forward_with_center_mask(x, target, mask, cmask)[1]

    try:
        cmask = model.center_frame_mask(mask, markers)
        gstep = GraphedTrainStep(Step(model), (x, target, mask.float(), cmask))
        for _ in range(2):
            gl = gstep(x, target, mask.float(), cmask)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.steps * 2):
            gstep(x, target, mask.float(), cmask)
        e1.record()
        torch.cuda.synchronize()
        gms = e0.elapsed_time(e1) / (a.steps * 2)
        print(f'MaskFeat MViT-B 16x224 batch {B} fwd+bwd (CUDA graph): {gms:.2f} ms/step = {B / gms * 1e3:.1f} clips/s; '
              f'loss {float(gl):.5f}; {gstep.kernels_per_replay} kernels per replay')
    except Exception as ex:       # report and keep the eager numbers
        import traceback
        traceback.print_exc()
        print('graph capture failed:', ex)
for _ in range(a.warmup):
    loss = step()
torch.cuda.synchronize()
print('loss', float(loss), 'finite grads', all(bool(torch.isfinite(p.grad).all()) for p in model.parameters()))
n0 = _lib.launch_count()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.steps):
    step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.steps
print(f'MaskFeat MViT-B 16x224 batch {B} fwd+bwd (eager): {ms:.2f} ms/step = {B / ms * 1e3:.1f} clips/s; '
      f'{(_lib.launch_count() - n0) // a.steps} kernel launches/step; peak mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB')
if a.profile:
    from torch.profiler import ProfilerActivity, profile
    shapes = []
    real_gemm = _lib.K.gemm

    def logged_gemm(a_, b_, M, N, Kd, **kw):
        shapes.append((M, N, Kd, int(kw.get('a_mn', False)), int(kw.get('b_mn', False)), kw.get('epi', 'bf16'),
                       'aux' if kw.get('aux') is not None else ''))
        return real_gemm(a_, b_, M, N, Kd, **kw)

    _lib.K.gemm = logged_gemm
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        step()
        torch.cuda.synchronize()
    _lib.K.gemm = real_gemm
    rows = sorted(prof.key_averages(), key=lambda r: -r.device_time_total)[:25]
    tot = sum(r.device_time_total for r in prof.key_averages())
    print(f'kernel time total {tot / 1e3:.2f} ms')
    for r in rows:
        print(f'{r.device_time_total / 1e3:9.3f} ms  {r.count:5d}x  {r.key[:110]}')
    # the 30 longest individual launches (which pooling / attention launches carry the time)
    singles = sorted((e for e in prof.events() if e.device_time_total > 0 and e.device_type.name == 'CUDA'),
                     key=lambda e: -e.device_time_total)[:30]
    print('longest single launches:')
    for e in singles:
        print(f'  {e.device_time_total:8.1f} us  {e.name[:100]}')
    # per-shape GEMM table: every K.gemm call launches exactly one *gemm*_tcgen05_kernel, in call order
    evs = sorted((e for e in prof.events() if 'tcgen05_kernel' in e.name and 'gemm' in e.name), key=lambda e: e.time_range.start)
    if len(evs) == len(shapes):
        agg = {}
        for sh, e in zip(shapes, evs):
            c = agg.setdefault(sh, [0, 0.0])
            c[0] += 1
            c[1] += e.device_time_total
        print('GEMM shapes (M, N, K, a_mn, b_mn, epilogue): count, total us, TFLOP/s')
        for sh, (cnt, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
            fl = 2.0 * sh[0] * sh[1] * sh[2] * cnt
            print(f'  {str(sh):62s} {cnt:3d}x {us:9.1f} us  {fl / us / 1e6:7.1f}')
    else:
        print(f'gemm table skipped: {len(evs)} kernel events vs {len(shapes)} calls')
