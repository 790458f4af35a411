import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from videotransformer_pytorch_b200 import _lib
K = _lib.K; lib = _lib.load_library()
dev = torch.device('cuda')
Bp, N, H, hd = 64, 197, 12, 64
qkv = (torch.randn(Bp, N, 3, H, hd, device=dev) * 0.7).bfloat16()
dctx = torch.randn(Bp * N, H * hd, device=dev).bfloat16()
for _ in range(3):
    ctx, lse, _ = K.attn_fwd(qkv, Bp, N, H, hd, 0.125)
    K.attn_bwd(qkv, ctx, dctx, lse, Bp, N, H, hd, 0.125)
torch.cuda.synchronize()
dbg = torch.zeros(Bp * H, 32, dtype=torch.int64, device=dev)
lib.vt_debug_buffer(ctypes.c_void_p(dbg.data_ptr()))
def show(title, names_c, names_e):
    d = dbg.cpu().double()
    t0 = d[:, 0:1]
    print(title)
    for i, n in names_c:
        v = d[:, i] - t0[:, 0]
        print(f'   ctl  {n:28s} {v.median().item():9.0f} {v.max().item():9.0f}')
    for i, n in names_e:
        v = d[:, 16 + i] - t0[:, 0]
        print(f'   warp0 {n:27s} {v.median().item():9.0f} {v.max().item():9.0f}')
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda._sleep(400000); e0.record(); ctx, lse, _ = K.attn_fwd(qkv, Bp, N, H, hd, 0.125); e1.record(); torch.cuda.synchronize()
show(f'attn_tc_fwd: {e0.elapsed_time(e1)*1e3:.1f} us; cycles since CTA entry (median / max over 768 CTAs)',
     [(1, 'TMA issued'), (2, 'Q,K landed'), (3, 'V landed'), (4, 'P0 ready -> PV0'), (5, 'P1 ready -> PV1'), (15, 'exit')],
     [(1, 'S ready'), (2, 'max pass done'), (3, 'P written'), (4, 'O ready'), (5, 'O stored'), (15, 'exit')])
dbg.zero_()
torch.cuda._sleep(400000); e0.record(); K.attn_bwd(qkv, ctx, dctx, lse, Bp, N, H, hd, 0.125); e1.record(); torch.cuda.synchronize()
show(f'attn_tc_bwd: {e0.elapsed_time(e1)*1e3:.1f} us',
     [(1, 'TMA issued'), (2, 'loads landed')] + [(3 + 2 * i, f'it{i} S,dP issued') for i in range(4)] + [(4 + 2 * i, f'it{i} P,dS ready') for i in range(4)] + [(11, 'all MMAs issued'), (15, 'exit')],
     [(1, 'delta done')] + [(2 + 2 * i, f'it{i} S,dP complete') for i in range(4)] + [(3 + 2 * i, f'it{i} P,dS written') for i in range(4)] + [(12, 'dQ ready'), (13, 'dQ stored'), (15, 'exit')])
lib.vt_debug_buffer(None)
