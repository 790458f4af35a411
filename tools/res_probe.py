"""Probe of the TMA residual epilogue: python tools/res_probe.py {plain|temporal|spatial} [cluster]  (one case per process)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ['VT_TMA_RES'] = '1'
os.environ.setdefault('VT_TMA_RES_SPATIAL', '1')
import torch
from videotransformer_pytorch_b200 import _lib, ops

K = _lib.K
kind = sys.argv[1]
cluster = int(sys.argv[2]) if len(sys.argv) > 2 else 1
dev = 'cuda'
g = torch.Generator().manual_seed(0)
B, T, P, D, Kd = 2, 8, 196, 768, 128
S = 1 + P * T
R = B * S
rel = lambda a, b: float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))
if kind == 'plain':
    M = 4096 if len(sys.argv) < 4 else int(sys.argv[3])
    a, w = torch.randn(M, Kd, generator=g).bfloat16().to(dev), torch.randn(D, Kd, generator=g).bfloat16().to(dev)
    aux = torch.randn(M, D, generator=g).to(dev)
    out = torch.empty(M, D, device=dev)
    K.gemm(a, w, M, D, Kd, epi='f32', aux=aux, out=out, force_cluster=cluster)
    torch.cuda.synchronize()
    print(kind, 'cluster', cluster, 'M', M, 'rel err', rel(out, a.float() @ w.float().t() + aux))
else:
    maps = ops.token_maps(B, T, P, 'cuda:0')
    aff = ops.affine_row_maps(B, T, P, D)[kind]
    x2 = torch.randn(R, D, generator=g).to(dev)
    w = torch.randn(D, Kd, generator=g).bfloat16().to(dev)
    if kind == 'temporal':
        M, rows, orow, arow = B * P * T, R, maps['temporal'], maps['temporal']
    else:
        M, rows, orow, arow = B * T * (P + 1), R + B * T, maps['sp_out'], maps['sp_aux']
    a = torch.randn(M, Kd, generator=g).bfloat16().to(dev)
    got = torch.full((rows, D), -7.0, device=dev)
    K.gemm(a, w, M, D, Kd, epi='f32', aux=x2, aux_row=arow, out=got, out_row=orow, row_map=aff, force_cluster=cluster)
    torch.cuda.synchronize()
    r = a.float() @ w.float().t()
    add = x2[arow.long().clamp(min=0)] * (arow >= 0)[:, None]
    full = torch.full((rows, D), -7.0, device=dev)
    full[orow.long()] = r + add
    print(kind, 'cluster', cluster, 'rel err', rel(got, full))
