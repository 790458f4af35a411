"""Timing of the remainder-row split (vt_gemm_rows.cu): python tools/rows_probe.py — each FFN GEMM shape at M = 12552 with
VT_ROWS_SPLIT off / on, and at M = 12544 (what the tensor-core kernel alone costs without the 8 extra rows)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from videotransformer_pytorch_b200 import _lib

K = _lib.K
dev = 'cuda'
g = torch.Generator().manual_seed(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record()
        torch.cuda.synchronize()
        tot += s.elapsed_time(e)
    return tot / iters * 1e3


rel = lambda a, b: float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))
CASES = [('fc1 (bf16)', 3072, 768, False, 'bf16'), ('fc2 (f32 resid)', 768, 3072, False, 'f32'),
         ('d fc2 (bf16)', 3072, 768, True, 'bf16'), ('d fc1 (bf16)', 768, 3072, True, 'bf16'),
         ('mvit fc1', 1536, 384, False, 'bf16'), ('mvit fc2 (f32 resid)', 384, 1536, False, 'f32'), ('mvit d qkv', 384, 1152, True, 'bf16')]
for name, N, Kd, b_mn, epi in CASES:
    M = 12552
    a = torch.randn(M, Kd, generator=g).bfloat16().to(dev)
    b = ((torch.randn(Kd, N, generator=g) if b_mn else torch.randn(N, Kd, generator=g)) * 0.05).bfloat16().to(dev)
    bias = torch.randn(N, generator=g).to(dev)
    aux = torch.randn(M, N, generator=g).to(dev) if epi == 'f32' else None
    kw = dict(b_mn=b_mn, epi=epi, bias=bias, aux=aux)
    out = torch.empty(M, N, dtype=torch.float32 if epi == 'f32' else torch.bfloat16, device=dev)
    res = {}
    for mode in ('0', '1'):
        os.environ['VT_ROWS_SPLIT'] = mode
        res[mode] = timed(lambda: K.gemm(a, b, M, N, Kd, out=out, **kw))
        res['out' + mode] = out.clone()
    os.environ['VT_ROWS_SPLIT'] = '0'
    M0 = 12544
    kw0 = dict(kw, aux=None if aux is None else aux[:M0])
    t0 = timed(lambda: K.gemm(a[:M0], b, M0, N, Kd, out=out[:M0], **kw0))
    ref = (a[M0:].float() @ (b.float() if b_mn else b.float().t()) + bias) + (aux[M0:] if aux is not None else 0)
    print(f'{name:22s} M=12552 one kernel {res["0"]:7.1f} us | split {res["1"]:7.1f} us | M=12544 alone {t0:7.1f} us | last rows vs fp32 '
          f'{rel(res["out1"][M0:].float(), ref):.2e} | head rows equal {bool(torch.equal(res["out0"][:M0], res["out1"][:M0]))}')
