"""Bring-up probe for the tcgen05 GEMM: runs one configuration and prints where it deviates.
usage: python tools/gemm_probe.py M N K a_mn b_mn [bn] [splits]"""
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from videotransformer_pytorch_b200 import _lib

M, N, Kd, a_mn, b_mn = [int(v) for v in sys.argv[1:6]]
bn = int(sys.argv[6]) if len(sys.argv) > 6 else 0
splits = int(sys.argv[7]) if len(sys.argv) > 7 else 0
cluster = int(sys.argv[8]) if len(sys.argv) > 8 else 0
torch.manual_seed(0)
a = torch.randn((Kd, M) if a_mn else (M, Kd)).cuda().bfloat16()
b = torch.randn((Kd, N) if b_mn else (N, Kd)).cuda().bfloat16()
out = _lib.K.gemm(a, b, M, N, Kd, a_mn=bool(a_mn), b_mn=bool(b_mn), epi='f32', force_bn=bn, split_ok=splits > 0,
                  force_splits=splits, force_cluster=cluster)
torch.cuda.synchronize()
A = a.float().t() if a_mn else a.float()
B = b.float() if b_mn else b.float().t()
ref = A @ B
err = (out - ref).abs()
rel = float((out - ref).norm() / ref.norm())
print(f'M={M} N={N} K={Kd} a_mn={a_mn} b_mn={b_mn} bn={bn} splits={splits} cluster={cluster}: rel={rel:.3e} max_abs={float(err.max()):.3e} '
      f'nan={int(torch.isnan(out).sum())} zeros={int((out == 0).sum())}')
if rel > 1e-3:
    bad_rows = (err.max(dim=1).values > 1e-2).nonzero().flatten()
    bad_cols = (err.max(dim=0).values > 1e-2).nonzero().flatten()
    print(' bad rows:', bad_rows[:16].tolist(), '... count', bad_rows.numel())
    print(' bad cols:', bad_cols[:16].tolist(), '... count', bad_cols.numel())
    print(' out[0,:8]', out[0, :8].tolist())
    print(' ref[0,:8]', ref[0, :8].tolist())
    # partial-K hypotheses
    for kk in range(16, Kd + 1, 16):
        r2 = A[:, :kk] @ B[:kk]
        e2 = float((out - r2).norm() / (r2.norm() + 1e-9))
        if e2 < 1e-3:
            print(f' matches partial sum over first {kk} of K')
