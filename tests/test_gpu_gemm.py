"""tcgen05 GEMM (vt_gemm) vs torch fp32 matmul on the same bf16-rounded operands.  -m gpu"""

import pytest
import torch

pytestmark = pytest.mark.gpu


def K():
    from videotransformer_pytorch_b200 import _lib
    return _lib.K


def mk(shape, seed, scale=1.0):
    g = torch.Generator(device='cpu').manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).cuda()


def ref_mm(a, b, a_mn, b_mn):
    A = a.float().t() if a_mn else a.float()
    B = b.float() if b_mn else b.float().t()
    return A @ B


def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


SHAPES = [(128, 128, 64), (128, 256, 64), (256, 256, 128), (384, 512, 192), (200, 136, 72), (1000, 768, 768),
          (12544, 2304, 768)]


@pytest.mark.parametrize('M,N,Kd', SHAPES)
@pytest.mark.parametrize('a_mn,b_mn', [(False, False), (False, True), (True, True), (True, False)])
def test_gemm_plain_f32(M, N, Kd, a_mn, b_mn):
    if (a_mn and M % 8) or (b_mn and N % 8) or Kd % 8:
        pytest.skip('leading dims must be multiples of 8')
    a = mk((Kd, M) if a_mn else (M, Kd), 1).bfloat16()
    b = mk((Kd, N) if b_mn else (N, Kd), 2).bfloat16()
    out = K().gemm(a, b, M, N, Kd, a_mn=a_mn, b_mn=b_mn, epi='f32')
    torch.cuda.synchronize()
    r = ref_mm(a, b, a_mn, b_mn)
    assert rel(out, r) < 1e-5, rel(out, r)


@pytest.mark.parametrize('bn', [128, 192, 256])
def test_gemm_bf16_bias_rowscale(bn):
    M, N, Kd = 640, 512, 256
    a, b = mk((M, Kd), 3).bfloat16(), mk((N, Kd), 4).bfloat16()
    bias, rs = mk((N,), 5), mk((M,), 6).abs()
    out = K().gemm(a, b, M, N, Kd, epi='bf16', bias=bias, row_scale=rs, force_bn=bn)
    r = (ref_mm(a, b, False, False) + bias) * rs[:, None]
    assert out.dtype == torch.bfloat16
    assert rel(out, r) < 4e-3


def test_gemm_f32_residual_rowmaps():
    M, N, Kd, R = 300, 256, 128, 400
    a, b = mk((M, Kd), 7).bfloat16(), mk((N, Kd), 8).bfloat16()
    bias, rs = mk((N,), 9), mk((M,), 10)
    aux = mk((R, N), 11)
    perm = torch.randperm(R, generator=torch.Generator().manual_seed(0))[:M]
    out_row = perm.to(torch.int32).cuda()
    out_row[5] = -1                      # skipped row
    aux_row = torch.randint(0, R, (M,), generator=torch.Generator().manual_seed(1)).to(torch.int32).cuda()
    aux_row[7] = -1                      # no addend
    out = torch.full((R, N), 123.0, device='cuda')
    K().gemm(a, b, M, N, Kd, epi='f32', bias=bias, row_scale=rs, aux=aux, aux_row=aux_row, out=out, out_row=out_row)
    r = (ref_mm(a, b, False, False) + bias) * rs[:, None]
    add = aux[aux_row.long().clamp(min=0)]
    add[7] = 0
    r = r + add
    exp = torch.full((R, N), 123.0, device='cuda')
    ok = out_row >= 0
    exp[out_row[ok].long()] = r[ok]
    assert rel(out, exp) < 1e-5


def test_gemm_gelu_and_dgelu():
    M, N, Kd = 256, 512, 128
    a, b = mk((M, Kd), 12, 0.3).bfloat16(), mk((N, Kd), 13, 0.3).bfloat16()
    bias = mk((N,), 14)
    z, h = K().gemm(a, b, M, N, Kd, epi='gelu', bias=bias)
    zr = ref_mm(a, b, False, False) + bias
    assert rel(z, zr) < 4e-3
    assert rel(h, torch.nn.functional.gelu(zr)) < 4e-3
    # dgelu: out = acc * gelu'(z)
    g = mk((M, Kd), 15, 0.3).bfloat16()
    out = K().gemm(g, b, M, N, Kd, epi='dgelu', aux=z)
    zz = z.float().requires_grad_(True)
    torch.nn.functional.gelu(zz).sum().backward()
    assert rel(out, ref_mm(g, b, False, False) * zz.grad) < 4e-3


@pytest.mark.parametrize('path', ['tma-reduce-add', 'workspace'])
@pytest.mark.parametrize('cluster', [1, 3], ids=['single-cta', 'cta-pair'])
@pytest.mark.parametrize('splits', [2, 5, 16])
def test_gemm_splitk_wgrad(splits, cluster, path, monkeypatch):
    """split-K partial tiles: reduce-added into the zeroed output by TMA (default) or summed from an fp32 workspace."""
    if path == 'workspace':
        monkeypatch.setenv('VT_SPLITK_WORKSPACE', '1')
    Mtok, Nout, Kin = 2048 + 64, 384, 256
    dy, x = mk((Mtok, Nout), 16).bfloat16(), mk((Mtok, Kin), 17).bfloat16()
    out = torch.full((Nout, Kin), 7.0, device='cuda')            # stale contents must not leak into the result
    K().gemm(dy, x, Nout, Kin, Mtok, a_mn=True, b_mn=True, epi='f32', split_ok=True, force_splits=splits,
             force_cluster=cluster, out=out)
    r = dy.float().t() @ x.float()
    assert rel(out, r) < 1e-5


def test_gemm_splitk_odd_tile_edges():
    """in-place split-K on a shape whose tiles hang over both output edges (TMA clips the reduce-add box)."""
    Mtok, Nout, Kin = 1000, 200, 136
    dy, x = mk((Mtok, Nout), 26).bfloat16(), mk((Mtok, Kin), 27).bfloat16()
    out = K().gemm(dy, x, Nout, Kin, Mtok, a_mn=True, b_mn=True, epi='f32', split_ok=True, force_splits=4)
    assert rel(out, dy.float().t() @ x.float()) < 1e-5


def test_gemm_wgrad_heuristic_split_big():
    Mtok, Nout, Kin = 12552, 768, 3072
    dy, x = mk((Mtok, Nout), 18, 0.1).bfloat16(), mk((Mtok, Kin), 19, 0.1).bfloat16()
    out = K().gemm(dy, x, Nout, Kin, Mtok, a_mn=True, b_mn=True, epi='f32', split_ok=True)
    r = dy.float().t() @ x.float()
    assert rel(out, r) < 1e-5


def test_gemm_rejects_bad_args():
    a, b = mk((64, 64), 1).bfloat16(), mk((64, 64), 2).bfloat16()
    with pytest.raises(RuntimeError):
        K().gemm(a.float(), b, 64, 64, 64)
    with pytest.raises(RuntimeError):
        K().gemm(a, b, 64, 60, 64)        # N not a multiple of 8 / shape mismatch


@pytest.mark.parametrize('M,N,Kd', [(12552, 768, 3072), (12608, 2304, 768), (1000, 576, 192)])
def test_gemm_bn192_and_auto_config(M, N, Kd):
    a, b = mk((M, Kd), 21, 0.2).bfloat16(), mk((N, Kd), 22, 0.2).bfloat16()
    r = ref_mm(a, b, False, False)
    for bn in (0, 192):
        out = K().gemm(a, b, M, N, Kd, epi='f32', force_bn=bn)
        assert rel(out, r) < 1e-5, bn


@pytest.mark.parametrize('shape', [(1000, 3072), (12552, 3072), (2344, 3072), (5, 8), (1, 8)])
def test_gelu_kernels(shape):
    """sizes below, between and above one / two grid strides (the kernels take two vectors per thread per iteration)"""
    torch.manual_seed(3)
    z = (torch.randn(*shape) * 1.5).cuda().bfloat16()
    h = K().gelu(z)
    zf = z.float().requires_grad_(True)
    ref = torch.nn.functional.gelu(zf)
    assert rel(h, ref) < 3e-3
    dh = torch.randn(*shape).cuda().bfloat16()
    ref.backward(dh.float())
    assert rel(K().dgelu(dh, z), zf.grad) < 3e-3


@pytest.mark.parametrize('cluster', [1, 2, 3])
@pytest.mark.parametrize('M,N,Kd,a_mn,b_mn,bn', [(128, 256, 128, False, False, 256), (100, 128, 64, False, False, 128),
                                                  (1000, 768, 768, False, False, 0), (12544, 768, 768, False, True, 0),
                                                  (12552, 3072, 768, False, False, 256), (12608, 768, 768, False, True, 192),
                                                  (2304, 768, 12544, True, True, 0), (640, 576, 320, True, True, 192),
                                                  (640, 384, 320, True, False, 128)])
def test_gemm_cluster_multicast(M, N, Kd, a_mn, b_mn, bn, cluster):
    if cluster == 3 and bn == 192:
        bn = 256          # the CTA-pair kernel has BN 128 / 256
    a = mk((Kd, M) if a_mn else (M, Kd), 31, 0.3).bfloat16()
    b = mk((Kd, N) if b_mn else (N, Kd), 32, 0.3).bfloat16()
    out = K().gemm(a, b, M, N, Kd, a_mn=a_mn, b_mn=b_mn, epi='f32', force_bn=bn, force_cluster=cluster,
                   split_ok=a_mn and b_mn)
    torch.cuda.synchronize()
    assert rel(out, ref_mm(a, b, a_mn, b_mn)) < 1e-5


@pytest.mark.parametrize('epi', ['bf16', 'f32res', 'gelu', 'dgelu'])
def test_gemm_pair_kernel_epilogues(epi):
    M, N, Kd = 1000, 512, 256
    a, b = mk((M, Kd), 41, 0.3).bfloat16(), mk((N, Kd), 42, 0.3).bfloat16()
    bias = mk((N,), 43)
    r = ref_mm(a, b, False, False)
    if epi == 'bf16':
        out = K().gemm(a, b, M, N, Kd, epi='bf16', bias=bias, force_cluster=3)
        assert rel(out, r + bias) < 4e-3
    elif epi == 'f32res':
        aux = mk((M, N), 44)
        perm = torch.randperm(M).to(torch.int32).cuda()
        out = torch.zeros(M, N, device='cuda')
        K().gemm(a, b, M, N, Kd, epi='f32', bias=bias, aux=aux, aux_row=perm, out_row=perm, out=out, force_cluster=3)
        exp = torch.zeros(M, N, device='cuda')
        exp[perm.long()] = r + bias + aux[perm.long()]
        assert rel(out, exp) < 1e-5
    elif epi == 'gelu':
        z, h = K().gemm(a, b, M, N, Kd, epi='gelu', bias=bias, force_cluster=3)
        assert rel(z, r + bias) < 4e-3 and rel(h, torch.nn.functional.gelu(r + bias)) < 4e-3
    else:
        z = mk((M, N), 45).bfloat16()
        out = K().gemm(a, b, M, N, Kd, epi='dgelu', aux=z, force_cluster=3)
        zz = z.float().requires_grad_(True)
        torch.nn.functional.gelu(zz).sum().backward()
        assert rel(out, r * zz.grad) < 4e-3


# ---- round 2: fp32 residual epilogue on TMA (affine row maps through a tensor map of the token stream), narrow tail units ---
def _ops():
    from videotransformer_pytorch_b200 import ops
    return ops


@pytest.mark.parametrize('cluster', [1, 3], ids=['single-cta', 'cta-pair'])
@pytest.mark.parametrize('bn', [0, 128, 192, 256])
@pytest.mark.parametrize('M,N,Kd', [(1000, 768, 256), (12552, 768, 768), (130, 96, 192), (4096, 256, 64)])
def test_residual_epilogue_tma_plain_rows(M, N, Kd, bn, cluster, monkeypatch):
    """y = s(m) (A B^T + bias) + aux without row maps (FC2, joint proj): TMA residual path == generic per-thread path."""
    if cluster == 3 and bn == 192:
        pytest.skip('the pair kernel has no 192-wide tile')
    a, b = mk((M, Kd), 30).bfloat16(), mk((N, Kd), 31).bfloat16()
    bias, rs, aux = mk((N,), 32), mk((M,), 33), mk((M, N), 34)
    out = torch.full((M, N), 55.0, device='cuda')
    monkeypatch.setenv('VT_TMA_RES', '1')
    K().gemm(a, b, M, N, Kd, epi='f32', bias=bias, row_scale=rs, aux=aux, out=out, force_bn=bn, force_cluster=cluster)
    r = (ref_mm(a, b, False, False) + bias) * rs[:, None] + aux
    assert rel(out, r) < 1e-5
    monkeypatch.setenv('VT_TMA_RES', '0')
    old = torch.empty_like(out)
    K().gemm(a, b, M, N, Kd, epi='f32', bias=bias, row_scale=rs, aux=aux, out=old, force_bn=bn, force_cluster=cluster)
    assert rel(out, old) < 1e-6


@pytest.mark.parametrize('cluster', [1, 3], ids=['single-cta', 'cta-pair'])
@pytest.mark.parametrize('B,T,P,D', [(2, 8, 196, 768), (3, 4, 9, 128), (1, 2, 50, 256), (2, 8, 196, 96)])
def test_residual_epilogue_tma_temporal_and_spatial_maps(B, T, P, D, cluster, monkeypatch):
    """The divided space-time scatters as TMA boxes (temporal '(b p) t', spatial '(b t) (1+p)' with the per-frame cls
    replicas going to side rows; 32-row groups that straddle a period go row by row) against the same GEMM driven by the
    out_row / aux_row arrays alone (generic epilogue)."""
    ops = _ops()
    maps = ops.token_maps(B, T, P, 'cuda:0')
    aff = ops.affine_row_maps(B, T, P, D)
    S = 1 + P * T
    R = B * S
    Kd = 128
    x2 = mk((R, D), 40)
    w, bias = mk((D, Kd), 41).bfloat16(), mk((D,), 42)
    for name, Mrows, out_rows, out_row, aux_row in (('temporal', B * P * T, R, maps['temporal'], maps['temporal']),
                                                    ('spatial', B * T * (P + 1), R + B * T, maps['sp_out'], maps['sp_aux'])):
        a = mk((Mrows, Kd), 43).bfloat16()
        rs = mk((Mrows,), 44)
        got = torch.full((out_rows, D), -7.0, device='cuda')
        monkeypatch.setenv('VT_TMA_RES', '1')
        monkeypatch.setenv('VT_TMA_RES_SPATIAL', '1')
        K().gemm(a, w, Mrows, D, Kd, epi='f32', bias=bias, row_scale=rs, aux=x2, aux_row=aux_row, out=got, out_row=out_row,
                 row_map=aff[name], force_cluster=cluster)
        monkeypatch.setenv('VT_TMA_RES', '0')
        exp = torch.full((out_rows, D), -7.0, device='cuda')
        K().gemm(a, w, Mrows, D, Kd, epi='f32', bias=bias, row_scale=rs, aux=x2, aux_row=aux_row, out=exp, out_row=out_row,
                 force_cluster=cluster)
        assert rel(got, exp) < 1e-6, name
        # rows the map never names (the cls row of every sample) keep their old contents
        assert bool((got[torch.arange(B, device='cuda') * S] == -7.0).all()), name
        r = (a.float() @ w.float().t() + bias) * rs[:, None]
        add = x2[aux_row.long().clamp(min=0)] * (aux_row >= 0)[:, None]
        full = torch.full((out_rows, D), -7.0, device='cuda')
        full[out_row.long()] = r + add
        assert rel(got, full) < 1e-5, name


@pytest.mark.parametrize('cluster', [0, 1, 3], ids=['auto', 'single-cta', 'cta-pair'])
@pytest.mark.parametrize('M,N,Kd', [(12552, 768, 768), (12608, 2304, 256), (12552, 3072, 128), (136, 512, 192), (264, 256, 64)])
@pytest.mark.parametrize('form', ['fwd_bf16', 'dgrad_bf16', 'fwd_f32_residual'])
def test_narrow_tail_units(M, N, Kd, form, cluster, monkeypatch):
    """A last row of tiles with <= 64 valid rows runs as narrow units (M = 12552 / 12608 of the FFN / spatial pass): same
    results as with the feature off, and against fp32 torch."""
    a = mk((M, Kd), 50).bfloat16()
    bias, rs = mk((N,), 51), mk((M,), 52)
    if form == 'dgrad_bf16':
        b = mk((Kd, N), 53).bfloat16()                       # W [n_out = Kd, k_in = N] read MN-major
        kw = dict(b_mn=True, epi='bf16', row_scale=rs)
        r = (a.float() @ b.float()) * rs[:, None]
    elif form == 'fwd_bf16':
        b = mk((N, Kd), 53).bfloat16()
        kw = dict(epi='bf16', bias=bias)
        r = a.float() @ b.float().t() + bias
    else:
        b = mk((N, Kd), 53).bfloat16()
        aux = mk((M, N), 54)
        kw = dict(epi='f32', bias=bias, row_scale=rs, aux=aux)
        r = (a.float() @ b.float().t() + bias) * rs[:, None] + aux
    monkeypatch.setenv('VT_TAIL_UNITS', '1')
    got = K().gemm(a, b, M, N, Kd, force_cluster=cluster, force_tail=2, **kw)
    off = K().gemm(a, b, M, N, Kd, force_cluster=cluster, force_tail=1, **kw)
    tol = 1e-5 if form == 'fwd_f32_residual' else 4e-3
    assert rel(got, r) < tol
    assert torch.equal(got, off)


@pytest.mark.parametrize('M,N,Kd', [(12552, 768, 3072), (1032, 200, 2048), (1025, 768, 2304), (1040, 384, 4096), (12552, 3072, 768)])
@pytest.mark.parametrize('form', ['fwd_bf16', 'dgrad_bf16', 'fwd_f32_residual', 'dgrad_f32_plain'])
def test_remainder_rows_split(M, N, Kd, form, monkeypatch):
    """M a few rows past a multiple of 128 (12552 = 98 x 128 + 8): the full row tiles run on the tensor cores, the last rows
    as CUDA-core dot products (vt_gemm_rows.cu).  Checked against fp32 torch, and row for row against the one-kernel path."""
    a = mk((M, Kd), 60).bfloat16()
    bias, rs = mk((N,), 61), mk((M,), 62)
    if form == 'dgrad_bf16':
        b = mk((Kd, N), 63).bfloat16()
        kw = dict(b_mn=True, epi='bf16', row_scale=rs)
        r = (a.float() @ b.float()) * rs[:, None]
    elif form == 'dgrad_f32_plain':
        b = mk((Kd, N), 63).bfloat16()
        kw = dict(b_mn=True, epi='f32', bias=bias)
        r = a.float() @ b.float() + bias
    elif form == 'fwd_bf16':
        b = mk((N, Kd), 63).bfloat16()
        kw = dict(epi='bf16', bias=bias)
        r = a.float() @ b.float().t() + bias
    else:
        b = mk((N, Kd), 63).bfloat16()
        aux = mk((M, N), 64)
        kw = dict(epi='f32', bias=bias, row_scale=rs, aux=aux)
        r = (a.float() @ b.float().t() + bias) * rs[:, None] + aux
    monkeypatch.setenv('VT_ROWS_SPLIT', '1')
    got = K().gemm(a, b, M, N, Kd, **kw)
    monkeypatch.setenv('VT_ROWS_SPLIT', '0')
    one = K().gemm(a, b, M, N, Kd, **kw)
    f32 = form in ('fwd_f32_residual', 'dgrad_f32_plain')
    tol = 1e-5 if f32 else 4e-3
    assert rel(got, r) < tol
    m0 = M // 128 * 128
    assert rel(got[:m0], one[:m0]) < 1e-6                      # the full row tiles: tensor-core kernel either way
    assert rel(got[m0:], r[m0:].to(got.device)) < tol          # the remainder rows on their own
    assert rel(got[m0:], one[m0:]) < (2e-5 if f32 else 8e-3)


@pytest.mark.parametrize('cluster', [1, 3], ids=['single-cta', 'cta-pair'])
@pytest.mark.parametrize('M,N,Kd', [(12552, 3072, 768), (1000, 512, 128), (130, 96, 64)])
def test_gelu_and_dgelu_epilogues_on_tma(M, N, Kd, cluster, monkeypatch):
    """FC1 with z / h = gelu(z) leaving as two TMA boxes, and the FC2 data gradient with gelu'(z) multiplied in from a
    TMA-loaded z box: both equal the generic epilogues bit for bit and match torch."""
    a, b = mk((M, Kd), 60, 0.3).bfloat16(), mk((N, Kd), 61, 0.3).bfloat16()
    bias = mk((N,), 62)
    monkeypatch.setenv('VT_TMA_GELU', '1')
    z, h = K().gemm(a, b, M, N, Kd, epi='gelu', bias=bias, force_cluster=cluster)
    zr = ref_mm(a, b, False, False) + bias
    assert rel(z, zr) < 4e-3 and rel(h, torch.nn.functional.gelu(zr)) < 4e-3
    monkeypatch.setenv('VT_TMA_GELU', '0')
    z0, _ = K().gemm(a, b, M, N, Kd, epi='gelu', bias=bias, force_cluster=cluster)
    # h is taken from the bf16-rounded z: identical to the stand-alone GELU kernel on z (the generic fused epilogue rounds later)
    assert torch.equal(z, z0) and torch.equal(h, K().gelu(z))
    g = mk((M, Kd), 63, 0.3).bfloat16()
    w = mk((Kd, N), 64, 0.3).bfloat16()                       # [n_out = Kd, k_in = N], read MN-major
    monkeypatch.setenv('VT_TMA_DGELU', '0')
    d0 = K().gemm(g, w, M, N, Kd, b_mn=True, epi='dgelu', aux=z, force_cluster=cluster)
    monkeypatch.setenv('VT_TMA_DGELU', '1')
    d1 = K().gemm(g, w, M, N, Kd, b_mn=True, epi='dgelu', aux=z, force_cluster=cluster)
    zz = z.float().requires_grad_(True)
    torch.nn.functional.gelu(zz).sum().backward()
    assert rel(d1, (g.float() @ w.float()) * zz.grad) < 4e-3
    assert torch.equal(d0, d1)


@pytest.mark.parametrize('res', ['0', '1'], ids=['generic-epilogue', 'tma-residual'])
@pytest.mark.parametrize('M,N,Kd', [(12544, 768, 768), (1000, 256, 64), (12552, 768, 3072)])
def test_second_bias_after_row_scale(M, N, Kd, res, monkeypatch):
    """out = s(m) (A B^T + bias) + bias2 + aux: the bias of a second linear layer folded into the GEMM (merged proj +
    temporal_fc), through the generic, the TMA-residual and the remainder-row paths."""
    monkeypatch.setenv('VT_TMA_RES', res)
    a, b = mk((M, Kd), 70).bfloat16(), mk((N, Kd), 71).bfloat16()
    bias, bias2, rs, aux = mk((N,), 72), mk((N,), 73), mk((M,), 74), mk((M, N), 75)
    out = K().gemm(a, b, M, N, Kd, epi='f32', bias=bias, bias2=bias2, row_scale=rs, aux=aux)
    r = (ref_mm(a, b, False, False) + bias) * rs[:, None] + bias2 + aux
    assert rel(out, r) < 1e-5


@pytest.mark.parametrize('M,N,Kd', [(768, 768, 12544), (2304, 768, 12608), (96, 448, 20000)])
def test_split_k_into_a_prezeroed_output(M, N, Kd):
    """Weight-gradient form (both operands MN-major, split-K by TMA reduce-add): with out_zeroed the library skips its own
    memset and accumulates into what the caller zeroed (gradient arena / DDP bucket); without it stale contents are harmless."""
    a, b = mk((Kd, M), 80).bfloat16(), mk((Kd, N), 81).bfloat16()
    r = a.float().t() @ b.float()
    stale = torch.full((M, N), 9.0, device='cuda')
    K().gemm(a, b, M, N, Kd, a_mn=True, b_mn=True, epi='f32', split_ok=True, out=stale)
    assert rel(stale, r) < 1e-5
    zeroed = torch.zeros((M, N), device='cuda')
    K().gemm(a, b, M, N, Kd, a_mn=True, b_mn=True, epi='f32', split_ok=True, out=zeroed, out_zeroed=True)
    assert rel(zeroed, r) < 1e-5
