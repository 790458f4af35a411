"""Isolated timing of every GEMM configuration of the TimeSformer-B step (batch 8): TFLOP/s per shape,
with L2 flushed between repetitions, next to torch.matmul (cuBLAS) on the same operands."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from videotransformer_pytorch_b200 import _lib

K = _lib.K
dev = torch.device('cuda')
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def bench(fn, reps=8):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        torch.cuda._sleep(400000)      # keep the GPU busy while the host prepares the launch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def case(name, M, N, Kd, a_mn=False, b_mn=False, epi='bf16', split=False, resid=False, bns=(0,), splits=(0,), clusters=(1, 3)):
    a = torch.randn((Kd, M) if a_mn else (M, Kd), device=dev).bfloat16()
    b = torch.randn((Kd, N) if b_mn else (N, Kd), device=dev).bfloat16()
    bias = torch.randn(N, device=dev) if epi in ('bf16', 'gelu') or resid else None
    aux = None
    kw = {}
    if resid:
        aux = torch.randn(M, N, device=dev)
        perm = torch.randperm(M, device=dev).to(torch.int32)
        kw = dict(aux=aux, aux_row=perm, out_row=perm, row_scale=torch.rand(M, device=dev))
    if epi == 'dgelu':
        kw = dict(aux=torch.randn(M, N, device=dev).bfloat16())
    out = torch.empty((M, N), dtype=torch.float32 if epi == 'f32' else torch.bfloat16, device=dev)
    out2 = torch.empty_like(out) if epi == 'gelu' else None
    fl = 2.0 * M * N * Kd
    A = a.t() if a_mn else a
    Bm = b if b_mn else b.t()
    t_ref = bench(lambda: torch.matmul(A, Bm))
    res = []
    for bn in bns:
        for sp in splits:
            for cl in clusters:
                t = bench(lambda: K.gemm(a, b, M, N, Kd, a_mn=a_mn, b_mn=b_mn, epi=epi, bias=bias, out=out, out2=out2,
                                         split_ok=split, force_bn=bn, force_splits=sp, force_cluster=cl, **kw))
                res.append(f'bn={bn or "auto"},sp={sp or "auto"},cl={cl}: {t * 1e3:6.1f}us {fl / t / 1e9:5.0f}TF')
    print(f'{name:28s} M={M:6d} N={N:5d} K={Kd:6d} | cuBLAS {t_ref * 1e3:7.1f}us {fl / t_ref / 1e9:6.0f}TF | ' + ' | '.join(res), flush=True)


print('== forward')
case('qkv temporal', 12544, 2304, 768, bns=(0,))
case('proj (bf16,rowscale)', 12544, 768, 768, bns=(0,))
case('temporal_fc (f32 resid map)', 12544, 768, 768, epi='f32', resid=True, bns=(0,))
case('qkv spatial', 12608, 2304, 768, bns=(0,))
case('fc1 gelu', 12552, 3072, 768, epi='gelu', bns=(0,))
case('fc2 (f32 resid)', 12552, 768, 3072, epi='f32', resid=True, bns=(0,))
print('== dgrad (B MN-major)')
case('d proj', 12544, 768, 768, b_mn=True, bns=(0,))
case('d qkv', 12544, 768, 2304, b_mn=True, bns=(0,))
case('d fc2 dgelu', 12552, 3072, 768, b_mn=True, epi='dgelu', bns=(0,))
case('d fc1', 12552, 768, 3072, b_mn=True, bns=(0,))
print('== wgrad (A,B MN-major, split-K)')
case('w 768x768', 768, 768, 12544, a_mn=True, b_mn=True, epi='f32', split=True, bns=(0,), splits=(0,))
case('w qkv 2304x768', 2304, 768, 12544, a_mn=True, b_mn=True, epi='f32', split=True, bns=(0,), splits=(0,))
case('w fc1 3072x768', 3072, 768, 12552, a_mn=True, b_mn=True, epi='f32', split=True, bns=(0,), splits=(0,))
case('w fc2 768x3072', 768, 3072, 12552, a_mn=True, b_mn=True, epi='f32', split=True, bns=(0,), splits=(0,))
print('== K sweep at M=12544 N=768 (epilogue share)')
for kd in (768, 3072):
    case(f'K={kd}', 12544, 768, kd, bns=(256,))
