"""Per-phase clock64 stamps of the GEMM kernel (diagnostics field of vt_gemm_params)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from videotransformer_pytorch_b200 import _lib
K = _lib.K
dev = torch.device('cuda')
for (M, N, Kd) in [(12544, 768, 768), (12544, 2304, 768), (12544, 768, 3072)]:
    a = torch.randn(M, Kd, device=dev).bfloat16(); b = torch.randn(N, Kd, device=dev).bfloat16()
    bias = torch.randn(N, device=dev)
    dbg = torch.zeros(148, 16, dtype=torch.int64, device=dev)
    for _ in range(3):
        K.gemm(a, b, M, N, Kd, bias=bias, epi='bf16', debug=dbg)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda._sleep(400000); e0.record(); K.gemm(a, b, M, N, Kd, bias=bias, epi='bf16', debug=dbg); e1.record(); torch.cuda.synchronize()
    d = dbg.cpu().double()
    t0 = d[:, 0:1]
    rel = (d[:, :8] - t0)
    names = ['entry', 'setup done', 'first operands', 'tile0 MMAs issued', 'last tile MMAs issued', 'tile0 drained', 'last tile drained', 'exit']
    print(f'M={M} N={N} K={Kd}: kernel {e0.elapsed_time(e1) * 1e3:.1f} us (events); cycles since CTA entry, median / max over CTAs')
    for i, n in enumerate(names):
        print(f'   {n:24s} {rel[:, i].median().item():10.0f} {rel[:, i].max().item():10.0f}')
