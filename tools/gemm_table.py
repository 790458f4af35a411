"""Isolated timing of every GEMM configuration of the TimeSformer-B step (batch 8): TFLOP/s per shape,
with L2 flushed between repetitions, next to torch.matmul (cuBLAS) on the same operands.

    python tools/gemm_table.py [quick]

Columns: cuBLAS (plain matmul of the same operands, no epilogue), then vt_gemm in the configuration the step uses (auto) and
with the single-CTA / CTA-pair kernel forced; residual rows use the real temporal / spatial row maps; `notail` = narrow tail
units off, `generic` = per-thread residual epilogue (VT_NO_TMA_RES) for comparison."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from videotransformer_pytorch_b200 import _lib, ops

# the round-2 paths are switched on here explicitly (they ship behind environment switches until confirmed on hardware)
for _name in ('VT_TMA_RES', 'VT_TAIL_UNITS', 'VT_TMA_GELU'):
    os.environ.setdefault(_name, '1')
K = _lib.K
dev = torch.device('cuda')
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
B, T, P, D = 8, 8, 196, 768
S = 1 + P * T
QUICK = len(sys.argv) > 1 and sys.argv[1] == 'quick'


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


def case(name, M, N, Kd, a_mn=False, b_mn=False, epi='bf16', split=False, resid=None, rowscale=False, variants=None):
    a = torch.randn((Kd, M) if a_mn else (M, Kd), device=dev).bfloat16()
    b = torch.randn((Kd, N) if b_mn else (N, Kd), device=dev).bfloat16()
    bias = torch.randn(N, device=dev) if (epi in ('bf16', 'gelu') or resid) and not split else None
    kw = {}
    out_rows = M
    if rowscale:
        kw['row_scale'] = torch.rand(M, device=dev)
    if resid == 'plain':
        kw.update(aux=torch.randn(M, N, device=dev))
    elif resid in ('temporal', 'spatial'):
        maps = ops.token_maps(B, T, P, str(dev))
        aff = ops.affine_row_maps(B, T, P, D)[resid]
        x2 = torch.randn(B * S, D, device=dev)
        out_rows = B * S + (B * T if resid == 'spatial' else 0)
        kw.update(aux=x2, aux_row=maps['temporal' if resid == 'temporal' else 'sp_aux'],
                  out_row=maps['temporal' if resid == 'temporal' else 'sp_out'], row_map=aff)
    if epi == 'dgelu':
        kw = dict(aux=torch.randn(M, N, device=dev).bfloat16())
    out = torch.empty((out_rows, N), dtype=torch.float32 if epi == 'f32' else torch.bfloat16, device=dev)
    out2 = torch.empty_like(out) if epi == 'gelu' else None
    fl = 2.0 * M * N * Kd
    A = a.t() if a_mn else a
    Bm = b if b_mn else b.t()
    t_ref = bench(lambda: torch.matmul(A, Bm))
    res = []
    variants = variants or [('auto', {}), ('single', dict(force_cluster=1)), ('pair', dict(force_cluster=3))]
    for label, extra in variants:
        env = extra.pop('env', None) if isinstance(extra, dict) else None      # (name, value) applied for this variant only
        call_kw = dict(kw)
        call_kw.update(extra)
        saved = None
        if env:
            saved = os.environ.get(env[0])
            os.environ[env[0]] = env[1]
        try:
            t = bench(lambda: K.gemm(a, b, M, N, Kd, a_mn=a_mn, b_mn=b_mn, epi=epi, bias=bias, out=out, out2=out2,
                                     split_ok=split, **call_kw))
        finally:
            if env:
                if saved is None:
                    os.environ.pop(env[0], None)
                else:
                    os.environ[env[0]] = saved
                extra['env'] = env
        res.append(f'{label}: {t * 1e3:6.1f}us {fl / t / 1e9:5.0f}TF')
    print(f'{name:30s} M={M:6d} N={N:5d} K={Kd:6d} | cuBLAS {t_ref * 1e3:6.1f}us {fl / t_ref / 1e9:5.0f}TF | ' + ' | '.join(res), flush=True)


std = [('auto', {}), ('single', dict(force_cluster=1)), ('pair', dict(force_cluster=3))]
tail = std + [('auto-notail', dict(force_tail=1))]
resv = std + [('generic', dict(env=('VT_TMA_RES', '0')))]
print('== forward')
case('qkv temporal', 12544, 2304, 768)
case('proj temporal (bf16,rowscale)', 12544, 768, 768, rowscale=True)
case('temporal_fc (f32 resid, map)', 12544, 768, 768, epi='f32', resid='temporal', variants=resv)
case('qkv spatial', 12608, 2304, 768, variants=tail)
case('proj spatial (f32 resid, map)', 12608, 768, 768, epi='f32', resid='spatial', rowscale=True, variants=resv + [('auto-notail', dict(force_tail=1))])
case('fc1 (bf16)', 12552, 3072, 768, variants=tail)
case('fc2 (f32 resid)', 12552, 768, 3072, epi='f32', resid='plain', rowscale=True, variants=resv + [('auto-notail', dict(force_tail=1))])
case('fc1 gelu epilogue (z,h by TMA)', 12552, 3072, 768, epi='gelu', variants=tail + [('generic', dict(env=('VT_TMA_GELU', '0')))])
print('== dgrad (B MN-major)')
case('d proj / d temporal_fc', 12544, 768, 768, b_mn=True, rowscale=True)
case('d qkv temporal', 12544, 768, 2304, b_mn=True)
case('d proj spatial', 12608, 768, 768, b_mn=True, variants=tail)
case('d qkv spatial', 12608, 768, 2304, b_mn=True, variants=tail)
case('d fc2 (bf16)', 12552, 3072, 768, b_mn=True, variants=tail)
case('d fc1', 12552, 768, 3072, b_mn=True, variants=tail)
case('d fc2 dgelu epilogue', 12552, 3072, 768, b_mn=True, epi='dgelu', variants=[('generic', {}), ('tma', dict(env=('VT_TMA_DGELU', '1')))])
print('== wgrad (A,B MN-major, split-K)')
case('w 768x768', 768, 768, 12544, a_mn=True, b_mn=True, epi='f32', split=True)
case('w qkv 2304x768', 2304, 768, 12544, a_mn=True, b_mn=True, epi='f32', split=True)
case('w fc1 3072x768', 3072, 768, 12552, a_mn=True, b_mn=True, epi='f32', split=True)
case('w fc2 768x3072', 768, 3072, 12552, a_mn=True, b_mn=True, epi='f32', split=True)
