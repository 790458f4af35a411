"""ncu target: a handful of step-shaped GEMM launches (warm-up first, then profiled range)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from videotransformer_pytorch_b200 import _lib
K = _lib.K
dev = torch.device('cuda')
M = 12552
xn = torch.randn(M, 768, device=dev).bfloat16()
w1 = (torch.randn(3072, 768, device=dev) * 0.03).bfloat16()
w2 = (torch.randn(768, 3072, device=dev) * 0.03).bfloat16()
wq = (torch.randn(2304, 768, device=dev) * 0.03).bfloat16()
b1 = torch.randn(3072, device=dev); b2 = torch.randn(768, device=dev); bq = torch.randn(2304, device=dev)
res = torch.randn(M, 768, device=dev)
g = torch.randn(M, 768, device=dev).bfloat16()

def run():
    qkv = K.gemm(xn, wq, M, 2304, 768, bias=bq, epi='bf16')                                  # 1 plain bf16
    z, h = K.gemm(xn, w1, M, 3072, 768, bias=b1, epi='gelu')                                  # 2 gelu
    y = K.gemm(h, w2, M, 768, 3072, bias=b2, epi='f32', aux=res)                              # 3 f32 + residual
    dz = K.gemm(g, w2, M, 3072, 768, b_mn=True, epi='dgelu', aux=z)                           # 4 dgelu
    dw = K.gemm(dz, xn, 3072, 768, M, a_mn=True, b_mn=True, epi='f32', split_ok=True)        # 5 wgrad split-K (+reduce)
    return y, dw

for _ in range(3):
    run()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
run()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
