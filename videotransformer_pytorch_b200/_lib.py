"""ctypes binding of libvt_b200.so (include/vt_b200.h) + the tensor-level kernel API used by ops.py.

There is deliberately NO fallback: if the shared library is missing or a kernel rejects its arguments a
RuntimeError is raised.  `K` is the process-wide kernel table; tests may swap it for the CPU emulation in
tests/emu_kernels.py to exercise the host-side logic without a GPU (the product never does).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libvt_b200.so')

EPI = {'bf16': 0, 'f32': 1, 'gelu': 2, 'dgelu': 3}

c_i32, c_i64, c_f32, c_vp = C.c_int32, C.c_int64, C.c_float, C.c_void_p


class GemmParams(C.Structure):
    _fields_ = [('a', c_vp), ('b', c_vp), ('lda', c_i64), ('ldb', c_i64),
                ('M', c_i32), ('N', c_i32), ('K', c_i32),
                ('a_mn_major', c_i32), ('b_mn_major', c_i32), ('epilogue', c_i32),
                ('bias', c_vp), ('out', c_vp), ('out2', c_vp), ('aux', c_vp),
                ('ldo', c_i64), ('ldo2', c_i64), ('ldaux', c_i64),
                ('out_row', c_vp), ('aux_row', c_vp), ('row_scale', c_vp),
                ('workspace', c_vp), ('workspace_bytes', c_i64),
                ('force_splits', c_i32), ('force_bn', c_i32), ('force_cluster', c_i32), ('debug', c_vp)]


class LnFwdParams(C.Structure):
    _fields_ = [('x', c_vp), ('ldx', c_i64), ('in_row', c_vp), ('gamma', c_vp), ('beta', c_vp),
                ('y', c_vp), ('mean', c_vp), ('rstd', c_vp), ('rows', c_i32), ('D', c_i32),
                ('eps', c_f32), ('y_fp32', c_i32)]


class LnBwdParams(C.Structure):
    _fields_ = [('dy', c_vp), ('dy_fp32', c_i32), ('x', c_vp), ('ldx', c_i64), ('in_row', c_vp),
                ('mean', c_vp), ('rstd', c_vp), ('gamma', c_vp),
                ('dres', c_vp), ('dx', c_vp), ('lddx', c_i64), ('dx_aux', c_vp), ('out_row', c_vp),
                ('partials', c_vp), ('rows', c_i32), ('D', c_i32)]


class ReduceParams(C.Structure):
    _fields_ = [('inp', c_vp), ('out', c_vp), ('stride', c_i64), ('S', c_i32), ('n', c_i64),
                ('accumulate', c_i32), ('scale', c_f32)]


class ColsumParams(C.Structure):
    _fields_ = [('inp', c_vp), ('ld', c_i64), ('M', c_i32), ('N', c_i32), ('out', c_vp), ('workspace', c_vp), ('counters', c_vp)]


class CastParams(C.Structure):
    _fields_ = [('src', c_vp), ('dst', c_vp), ('n', c_i64)]


class GatherCastParams(C.Structure):
    _fields_ = [('src', c_vp), ('lds', c_i64), ('in_row', c_vp), ('row_scale', c_vp), ('dst', c_vp),
                ('rows', c_i32), ('D', c_i32)]


class GeluParams(C.Structure):
    _fields_ = [('z', c_vp), ('dh', c_vp), ('out', c_vp), ('n', c_i64)]


class AttnFwdParams(C.Structure):
    _fields_ = [('qkv', c_vp), ('ctx', c_vp), ('lse', c_vp), ('probs', c_vp),
                ('Bp', c_i32), ('N', c_i32), ('H', c_i32), ('hd', c_i32), ('scale', c_f32), ('impl', c_i32)]


class AttnBwdParams(C.Structure):
    _fields_ = [('qkv', c_vp), ('ctx', c_vp), ('dctx', c_vp), ('lse', c_vp), ('dqkv', c_vp),
                ('Bp', c_i32), ('N', c_i32), ('H', c_i32), ('hd', c_i32), ('scale', c_f32), ('impl', c_i32)]


class Im2colParams(C.Structure):
    _fields_ = [('x', c_vp), ('cols', c_vp), ('B', c_i32), ('T', c_i32), ('C', c_i32), ('H', c_i32),
                ('W', c_i32), ('tube', c_i32), ('ph', c_i32), ('pw', c_i32)]


class HogParams(C.Structure):
    _fields_ = [('frames', c_vp), ('lut', c_vp), ('feat', c_vp), ('bins', c_vp),
                ('F', c_i32), ('H', c_i32), ('W', c_i32)]


EXPORTS = ['vt_version', 'vt_last_error', 'vt_sm_count', 'vt_set_reserved_sms', 'vt_launch_count', 'vt_gemm', 'vt_layernorm_fwd', 'vt_ln_bwd_blocks',
           'vt_layernorm_bwd', 'vt_reduce_rows', 'vt_colsum_chunks', 'vt_colsum_bf16', 'vt_cast_f32_bf16',
           'vt_gather_cast_bf16', 'vt_gelu_fwd_bf16', 'vt_gelu_bwd_bf16', 'vt_attn_fwd', 'vt_attn_bwd', 'vt_debug_buffer', 'vt_im2col_bf16', 'vt_col2im_f32', 'vt_hog']

_dll = None


def load_library() -> C.CDLL:
    """Load libvt_b200.so or fail loudly (no fallback path exists)."""
    global _dll
    if _dll is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f'{LIB_PATH} not found: build the sm_100a kernels first '
                f'(python -m videotransformer_pytorch_b200.build, or __graft_entry__.build()). '
                f'There is no CPU / library fallback for the hot path.')
        _dll = C.CDLL(LIB_PATH)
        for name in EXPORTS:
            getattr(_dll, name).restype = C.c_int
        if _dll.vt_version() != 1:
            raise RuntimeError('libvt_b200.so ABI version mismatch')
    return _dll


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _check(rc: int, what: str):
    if rc != 0:
        buf = C.create_string_buffer(512)
        load_library().vt_last_error(buf, 512)
        raise RuntimeError(f'{what} failed (code {rc}): {buf.value.decode(errors="replace")}')


def _req(t: torch.Tensor, dtype, name: str):
    if not t.is_cuda:
        raise RuntimeError(f'{name}: expected a CUDA tensor (the hot path has no CPU fallback)')
    if t.dtype != dtype:
        raise RuntimeError(f'{name}: expected dtype {dtype}, got {t.dtype}')
    return t


def _rows2d(t: torch.Tensor, name: str):
    if t.dim() != 2 or t.stride(1) != 1:
        raise RuntimeError(f'{name}: expected a 2-D row-contiguous tensor, got shape {tuple(t.shape)} stride {t.stride()}')
    return t


class CudaKernels:
    """Tensor-level wrappers; every method enqueues on torch's current CUDA stream."""

    name = 'cuda'

    def __init__(self):
        self._ws = {}
        self._counters = {}

    # -- scratch ------------------------------------------------------------------------------
    def workspace(self, device, nbytes: int) -> torch.Tensor:
        key = (device.index, torch.cuda.current_stream(device).cuda_stream)
        ws = self._ws.get(key)
        if ws is None or ws.numel() * 4 < nbytes:
            ws = torch.empty(max(nbytes, 64 << 20) // 4, dtype=torch.float32, device=device)
            self._ws[key] = ws
        return ws

    # -- GEMM ---------------------------------------------------------------------------------
    def gemm(self, a, b, M, N, Kdim, *, a_mn=False, b_mn=False, epi='bf16', bias=None, out=None, out2=None,
             aux=None, out_row=None, aux_row=None, row_scale=None, out_rows=None, split_ok=False,
             force_splits=0, force_bn=0, force_cluster=0, debug=None):
        lib = load_library()
        _rows2d(_req(a, torch.bfloat16, 'gemm.a'), 'gemm.a')
        _rows2d(_req(b, torch.bfloat16, 'gemm.b'), 'gemm.b')
        exp_a = (Kdim, M) if a_mn else (M, Kdim)
        exp_b = (Kdim, N) if b_mn else (N, Kdim)
        if tuple(a.shape) != exp_a or tuple(b.shape) != exp_b:
            raise RuntimeError(f'gemm: operand shapes {tuple(a.shape)} {tuple(b.shape)} != expected {exp_a} {exp_b}')
        odt = torch.float32 if epi == 'f32' else torch.bfloat16
        if out is None:
            out = torch.empty((out_rows if out_rows is not None else M, N), dtype=odt, device=a.device)
        _rows2d(_req(out, odt, 'gemm.out'), 'gemm.out')
        if epi == 'gelu' and out2 is None:
            out2 = torch.empty_like(out)
        p = GemmParams()
        p.a, p.b = a.data_ptr(), b.data_ptr()
        p.lda, p.ldb = a.stride(0), b.stride(0)
        p.M, p.N, p.K = M, N, Kdim
        p.a_mn_major, p.b_mn_major = int(a_mn), int(b_mn)
        p.epilogue = EPI[epi]
        p.bias = _ptr(None if bias is None else _req(bias, torch.float32, 'gemm.bias'))
        p.out, p.ldo = out.data_ptr(), out.stride(0)
        if out2 is not None:
            p.out2, p.ldo2 = out2.data_ptr(), out2.stride(0)
        if aux is not None:
            _rows2d(_req(aux, torch.float32 if epi == 'f32' else torch.bfloat16, 'gemm.aux'), 'gemm.aux')
            p.aux, p.ldaux = aux.data_ptr(), aux.stride(0)
        for nm, t in (('out_row', out_row), ('aux_row', aux_row)):
            if t is not None:
                _req(t, torch.int32, 'gemm.' + nm)
                if t.numel() != M:
                    raise RuntimeError(f'gemm.{nm}: expected {M} entries')
                setattr(p, nm, t.data_ptr())
        if row_scale is not None:
            _req(row_scale, torch.float32, 'gemm.row_scale')
            if row_scale.numel() != M:
                raise RuntimeError(f'gemm.row_scale: expected {M} entries, got {row_scale.numel()}')
            p.row_scale = row_scale.data_ptr()
        ws = None
        if split_ok and epi == 'f32':
            ws = self.workspace(a.device, 16 * M * N * 4)
            p.workspace, p.workspace_bytes = ws.data_ptr(), ws.numel() * 4
        p.force_splits, p.force_bn, p.force_cluster = force_splits, force_bn, force_cluster
        p.debug = _ptr(debug)
        _check(lib.vt_gemm(C.byref(p), _stream()), 'vt_gemm')
        return (out, out2) if epi == 'gelu' else out

    # -- LayerNorm ----------------------------------------------------------------------------
    def ln_fwd(self, x2d, gamma, beta, eps, in_row=None, rows=None, out_fp32=False):
        lib = load_library()
        _rows2d(_req(x2d, torch.float32, 'ln_fwd.x'), 'ln_fwd.x')
        rows = x2d.shape[0] if rows is None else rows
        D = x2d.shape[1]
        y = torch.empty((rows, D), dtype=torch.float32 if out_fp32 else torch.bfloat16, device=x2d.device)
        mean = torch.empty(rows, dtype=torch.float32, device=x2d.device)
        rstd = torch.empty_like(mean)
        p = LnFwdParams()
        p.x, p.ldx = x2d.data_ptr(), x2d.stride(0)
        p.in_row = _ptr(None if in_row is None else _req(in_row, torch.int32, 'ln_fwd.in_row'))
        p.gamma, p.beta = _req(gamma, torch.float32, 'gamma').data_ptr(), _req(beta, torch.float32, 'beta').data_ptr()
        p.y, p.mean, p.rstd = y.data_ptr(), mean.data_ptr(), rstd.data_ptr()
        p.rows, p.D, p.eps, p.y_fp32 = rows, D, eps, int(out_fp32)
        _check(lib.vt_layernorm_fwd(C.byref(p), _stream()), 'vt_layernorm_fwd')
        return y, mean, rstd

    def ln_bwd(self, dy, x2d, mean, rstd, gamma, in_row=None, out_row=None, dres=None, dx=None, n_aux=0):
        """-> (dx [x2d.shape] or given, dx_aux [n_aux, D] or None, dgamma, dbeta)"""
        lib = load_library()
        rows, D = dy.shape
        dev = dy.device
        if dx is None:
            dx = torch.empty((x2d.shape[0], D), dtype=torch.float32, device=dev)
        dx_aux = torch.empty((n_aux, D), dtype=torch.float32, device=dev) if n_aux else None
        blocks = lib.vt_ln_bwd_blocks(rows)
        partials = torch.empty((blocks, 2, D), dtype=torch.float32, device=dev)
        p = LnBwdParams()
        p.dy, p.dy_fp32 = dy.data_ptr(), int(dy.dtype == torch.float32)
        p.x, p.ldx = x2d.data_ptr(), x2d.stride(0)
        p.in_row = _ptr(in_row)
        p.mean, p.rstd, p.gamma = mean.data_ptr(), rstd.data_ptr(), gamma.data_ptr()
        p.dres = _ptr(dres)
        p.dx, p.lddx = dx.data_ptr(), dx.stride(0)
        p.dx_aux = _ptr(dx_aux)
        p.out_row = _ptr(out_row)
        p.partials = partials.data_ptr()
        p.rows, p.D = rows, D
        _check(lib.vt_layernorm_bwd(C.byref(p), _stream()), 'vt_layernorm_bwd')
        gb = torch.empty((2, D), dtype=torch.float32, device=dev)
        r = ReduceParams()
        r.inp, r.out, r.stride, r.S, r.n, r.accumulate, r.scale = partials.data_ptr(), gb.data_ptr(), 2 * D, blocks, 2 * D, 0, 1.0
        _check(lib.vt_reduce_rows(C.byref(r), _stream()), 'vt_reduce_rows')
        return dx, dx_aux, gb[0], gb[1]

    # -- reductions / casts -------------------------------------------------------------------
    def colsum(self, x):
        lib = load_library()
        _rows2d(_req(x, torch.bfloat16, 'colsum.x'), 'colsum.x')
        M, N = x.shape
        out = torch.empty(N, dtype=torch.float32, device=x.device)
        ws = torch.empty((lib.vt_colsum_chunks(M), N), dtype=torch.float32, device=x.device)
        key = (x.device.index, torch.cuda.current_stream(x.device).cuda_stream)
        cnt = self._counters.get(key)
        if cnt is None:
            cnt = self._counters[key] = torch.zeros(1024, dtype=torch.int32, device=x.device)
        p = ColsumParams()
        p.inp, p.ld, p.M, p.N, p.out, p.workspace = x.data_ptr(), x.stride(0), M, N, out.data_ptr(), ws.data_ptr()
        p.counters = cnt.data_ptr() if N <= 64 * 1024 else None
        _check(lib.vt_colsum_bf16(C.byref(p), _stream()), 'vt_colsum_bf16')
        return out

    def cast_bf16(self, x):
        lib = load_library()
        x = _req(x, torch.float32, 'cast.x').contiguous()
        out = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
        p = CastParams()
        p.src, p.dst, p.n = x.data_ptr(), out.data_ptr(), x.numel()
        _check(lib.vt_cast_f32_bf16(C.byref(p), _stream()), 'vt_cast_f32_bf16')
        return out

    def gather_cast(self, src2d, in_row=None, row_scale=None, rows=None):
        lib = load_library()
        _rows2d(_req(src2d, torch.float32, 'gather_cast.src'), 'gather_cast.src')
        rows = src2d.shape[0] if rows is None else rows
        D = src2d.shape[1]
        out = torch.empty((rows, D), dtype=torch.bfloat16, device=src2d.device)
        p = GatherCastParams()
        p.src, p.lds = src2d.data_ptr(), src2d.stride(0)
        p.in_row, p.row_scale = _ptr(in_row), _ptr(row_scale)
        p.dst, p.rows, p.D = out.data_ptr(), rows, D
        _check(lib.vt_gather_cast_bf16(C.byref(p), _stream()), 'vt_gather_cast_bf16')
        return out

    def gelu(self, z):
        lib = load_library()
        z = _req(z, torch.bfloat16, 'gelu.z')
        if not z.is_contiguous():
            raise RuntimeError('gelu: z must be contiguous')
        out = torch.empty_like(z)
        p = GeluParams()
        p.z, p.dh, p.out, p.n = z.data_ptr(), None, out.data_ptr(), z.numel()
        _check(lib.vt_gelu_fwd_bf16(C.byref(p), _stream()), 'vt_gelu_fwd_bf16')
        return out

    def dgelu(self, dh, z):
        lib = load_library()
        for t, n in ((dh, 'dh'), (z, 'z')):
            _req(t, torch.bfloat16, 'dgelu.' + n)
            if not t.is_contiguous():
                raise RuntimeError(f'dgelu: {n} must be contiguous')
        out = torch.empty_like(z)
        p = GeluParams()
        p.z, p.dh, p.out, p.n = z.data_ptr(), dh.data_ptr(), out.data_ptr(), z.numel()
        _check(lib.vt_gelu_bwd_bf16(C.byref(p), _stream()), 'vt_gelu_bwd_bf16')
        return out

    # -- attention ----------------------------------------------------------------------------
    def attn_fwd(self, qkv, Bp, N, H, hd, scale, want_probs=False, impl=0):
        lib = load_library()
        _req(qkv, torch.bfloat16, 'attn.qkv')
        if not qkv.is_contiguous() or qkv.numel() != Bp * N * 3 * H * hd:
            raise RuntimeError('attn_fwd: qkv must be contiguous [Bp, N, 3, H, hd]')
        ctx = torch.empty((Bp * N, H * hd), dtype=torch.bfloat16, device=qkv.device)
        lse = torch.empty((Bp, H, N), dtype=torch.float32, device=qkv.device)
        probs = torch.empty((Bp, H, N, N), dtype=torch.float32, device=qkv.device) if want_probs else None
        p = AttnFwdParams()
        p.qkv, p.ctx, p.lse, p.probs = qkv.data_ptr(), ctx.data_ptr(), lse.data_ptr(), _ptr(probs)
        p.Bp, p.N, p.H, p.hd, p.scale, p.impl = Bp, N, H, hd, scale, impl
        _check(lib.vt_attn_fwd(C.byref(p), _stream()), 'vt_attn_fwd')
        return ctx, lse, probs

    def attn_bwd(self, qkv, ctx, dctx, lse, Bp, N, H, hd, scale, impl=0):
        lib = load_library()
        for t, n in ((qkv, 'qkv'), (ctx, 'ctx'), (dctx, 'dctx')):
            _req(t, torch.bfloat16, 'attn_bwd.' + n)
            if not t.is_contiguous():
                raise RuntimeError(f'attn_bwd: {n} must be contiguous')
        dqkv = torch.empty_like(qkv)
        p = AttnBwdParams()
        p.qkv, p.ctx, p.dctx, p.lse, p.dqkv = qkv.data_ptr(), ctx.data_ptr(), dctx.data_ptr(), lse.data_ptr(), dqkv.data_ptr()
        p.Bp, p.N, p.H, p.hd, p.scale, p.impl = Bp, N, H, hd, scale, impl
        _check(lib.vt_attn_bwd(C.byref(p), _stream()), 'vt_attn_bwd')
        return dqkv

    # -- patch embedding operand -------------------------------------------------------------
    def im2col(self, x, tube, ph, pw):
        lib = load_library()
        x = _req(x, torch.float32, 'im2col.x').contiguous()
        B, T, Cc, H, W = x.shape
        rows = B * (T // tube) * (H // ph) * (W // pw)
        cols = torch.empty((rows, Cc * tube * ph * pw), dtype=torch.bfloat16, device=x.device)
        p = Im2colParams()
        p.x, p.cols = x.data_ptr(), cols.data_ptr()
        p.B, p.T, p.C, p.H, p.W, p.tube, p.ph, p.pw = B, T, Cc, H, W, tube, ph, pw
        _check(lib.vt_im2col_bf16(C.byref(p), _stream()), 'vt_im2col_bf16')
        return cols

    def col2im(self, cols, shape, tube, ph, pw):
        lib = load_library()
        cols = _req(cols, torch.float32, 'col2im.cols').contiguous()
        B, T, Cc, H, W = shape
        dx = torch.empty(shape, dtype=torch.float32, device=cols.device)
        p = Im2colParams()
        p.x, p.cols = cols.data_ptr(), dx.data_ptr()   # same POD layout: (src, dst, dims)
        p.B, p.T, p.C, p.H, p.W, p.tube, p.ph, p.pw = B, T, Cc, H, W, tube, ph, pw
        _check(lib.vt_col2im_f32(C.byref(p), _stream()), 'vt_col2im_f32')
        return dx

    # -- HOG ------------------------------------------------------------------------------------
    def hog(self, frames, lut, want_bins=False):
        lib = load_library()
        frames = _req(frames, torch.uint8, 'hog.frames').contiguous()
        F, H, W, c3 = frames.shape
        if c3 != 3:
            raise RuntimeError('hog: frames must be [F, H, W, 3] uint8')
        feat = torch.empty((F, H // 16, W // 16, 108), dtype=torch.float32, device=frames.device)
        bins = torch.empty((F, 3, H, W), dtype=torch.uint8, device=frames.device) if want_bins else None
        p = HogParams()
        p.frames, p.lut, p.feat, p.bins = frames.data_ptr(), _req(lut, torch.uint8, 'hog.lut').data_ptr(), feat.data_ptr(), _ptr(bins)
        p.F, p.H, p.W = F, H, W
        _check(lib.vt_hog(C.byref(p), _stream()), 'vt_hog')
        return feat, bins


def set_reserved_sms(n: int) -> None:
    """Keep n SMs free of persistent GEMM CTAs (for overlapped NCCL kernels); see vt_set_reserved_sms."""
    _check(load_library().vt_set_reserved_sms(int(n)), 'vt_set_reserved_sms')


def launch_count() -> int:
    """Kernels launched by libvt_b200.so in this process so far."""
    return int(load_library().vt_launch_count())


K = CudaKernels()
