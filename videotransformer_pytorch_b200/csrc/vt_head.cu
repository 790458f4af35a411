// Small fp32 kernels around the transformer body:
//   * skinny linear layer (classification head: 8 x 768 -> 400) forward / backward — warp-per-output GEMV, fp32 throughout
//   * softmax cross-entropy (hard labels or soft targets) forward + gradient in one launch
//   * attention probabilities softmax(q k^T * scale) for long sequences (get_last_selfattention of the joint variants)
//   * uint8 clip -> normalised bf16 patch operand with Mixup / CutMix of the flipped batch folded in
// All are latency / bandwidth bound warp-primitive kernels (no tensor cores: M <= 64 rows or one-off visualisation work).
#include "vt_common.cuh"

namespace vt {

// ------------------------------------------------------------------------------------------------
// y[m, n] = sum_k x[m, k] * W[n, k] + b[n]      (fp32; one warp per output column n, 8 rows of x per pass)
// ------------------------------------------------------------------------------------------------
constexpr int LS_ROWS = 8;

__global__ void __launch_bounds__(256)
linear_small_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                        float* __restrict__ y, int M, int N, int K) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.x * 8 + warp;
  const int m0 = blockIdx.y * LS_ROWS;
  if (n >= N) return;
  float acc[LS_ROWS];
#pragma unroll
  for (int r = 0; r < LS_ROWS; ++r) acc[r] = 0.f;
  const float* wr = w + (long long)n * K;
  for (int k = lane * 4; k < K; k += 128) {
    const float4 wv = __ldg(reinterpret_cast<const float4*>(wr + k));
#pragma unroll
    for (int r = 0; r < LS_ROWS; ++r) {
      if (m0 + r < M) {
        const float4 xv = __ldg(reinterpret_cast<const float4*>(x + (long long)(m0 + r) * K + k));
        acc[r] = fmaf(xv.x, wv.x, fmaf(xv.y, wv.y, fmaf(xv.z, wv.z, fmaf(xv.w, wv.w, acc[r]))));
      }
    }
  }
#pragma unroll
  for (int r = 0; r < LS_ROWS; ++r) acc[r] = warp_sum(acc[r]);
  if (lane == 0) {
    const float bias = b ? b[n] : 0.f;
#pragma unroll
    for (int r = 0; r < LS_ROWS; ++r)
      if (m0 + r < M) y[(long long)(m0 + r) * N + n] = acc[r] + bias;
  }
}

// dW[n, k] = sum_m dy[m, n] x[m, k] ;  db[n] = sum_m dy[m, n]          (one warp per n)
__global__ void __launch_bounds__(256)
linear_small_wgrad_kernel(const float* __restrict__ dy, const float* __restrict__ x, float* __restrict__ dw,
                          float* __restrict__ db, int M, int N, int K) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.x * 8 + warp;
  if (n >= N) return;
  for (int k = lane * 4; k < K; k += 128) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int m = 0; m < M; ++m) {
      const float g = __ldg(dy + (long long)m * N + n);
      const float4 xv = __ldg(reinterpret_cast<const float4*>(x + (long long)m * K + k));
      acc.x = fmaf(g, xv.x, acc.x); acc.y = fmaf(g, xv.y, acc.y); acc.z = fmaf(g, xv.z, acc.z); acc.w = fmaf(g, xv.w, acc.w);
    }
    *reinterpret_cast<float4*>(dw + (long long)n * K + k) = acc;
  }
  if (db && lane == 0) {
    float s = 0.f;
    for (int m = 0; m < M; ++m) s += dy[(long long)m * N + n];
    db[n] = s;
  }
}

// dx[m, k] = sum_n dy[m, n] W[n, k]: block = 128 columns x 8 rows, the 8 warps split n and merge through shared memory
__global__ void __launch_bounds__(256)
linear_small_dgrad_kernel(const float* __restrict__ dy, const float* __restrict__ w, float* __restrict__ dx, int M, int N,
                          int K) {
  __shared__ float4 red[8][LS_ROWS][32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k = blockIdx.x * 128 + lane * 4;
  const int m0 = blockIdx.y * LS_ROWS;
  float4 acc[LS_ROWS];
#pragma unroll
  for (int r = 0; r < LS_ROWS; ++r) acc[r] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (k < K) {
    for (int n = warp; n < N; n += 8) {
      const float4 wv = __ldg(reinterpret_cast<const float4*>(w + (long long)n * K + k));
#pragma unroll
      for (int r = 0; r < LS_ROWS; ++r) {
        if (m0 + r < M) {
          const float g = __ldg(dy + (long long)(m0 + r) * N + n);
          acc[r].x = fmaf(g, wv.x, acc[r].x); acc[r].y = fmaf(g, wv.y, acc[r].y);
          acc[r].z = fmaf(g, wv.z, acc[r].z); acc[r].w = fmaf(g, wv.w, acc[r].w);
        }
      }
    }
  }
#pragma unroll
  for (int r = 0; r < LS_ROWS; ++r) red[warp][r][lane] = acc[r];
  __syncthreads();
  if (k < K && warp < LS_ROWS && m0 + warp < M) {          // warp r finalises row r
    float4 s = red[0][warp][lane];
#pragma unroll
    for (int q = 1; q < 8; ++q) {
      const float4 t = red[q][warp][lane];
      s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
    }
    *reinterpret_cast<float4*>(dx + (long long)(m0 + warp) * K + k) = s;
  }
}

// ------------------------------------------------------------------------------------------------
// Softmax cross-entropy, mean over the M rows (nn.CrossEntropyLoss; timm SoftTargetCrossEntropy with `soft`):
//   loss = 1/M sum_m ( lse(z_m) * sum_c t_mc - sum_c t_mc z_mc ),   dz_mc = (softmax(z_m)_c * sum_c' t_mc' - t_mc) / M
// One CTA, rows in sequence (M is the per-GPU batch); deterministic.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
softmax_ce_kernel(const float* __restrict__ z, const int64_t* __restrict__ labels, const float* __restrict__ soft,
                  float* __restrict__ loss, float* __restrict__ row_loss, float* __restrict__ dz, int M, int N) {
  __shared__ float sh[8];
  __shared__ float total;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  auto block_reduce = [&](float v, bool is_max) {
    v = is_max ? warp_max(v) : warp_sum(v);
    __syncthreads();
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    float r = sh[0];
#pragma unroll
    for (int i = 1; i < 8; ++i) r = is_max ? fmaxf(r, sh[i]) : r + sh[i];
    return r;
  };
  if (threadIdx.x == 0) total = 0.f;
  const float inv_m = 1.0f / (float)M;
  for (int m = 0; m < M; ++m) {
    const float* zr = z + (long long)m * N;
    float mx = -INFINITY;
    for (int c = threadIdx.x; c < N; c += 256) mx = fmaxf(mx, zr[c]);
    mx = block_reduce(mx, true);
    float se = 0.f, tz = 0.f, ts = 0.f;
    const long long lab = labels ? (long long)labels[m] : -1;
    for (int c = threadIdx.x; c < N; c += 256) {
      const float v = zr[c];
      se += expf(v - mx);
      const float t = soft ? soft[(long long)m * N + c] : (c == lab ? 1.f : 0.f);
      tz = fmaf(t, v, tz);
      ts += t;
    }
    se = block_reduce(se, false);
    tz = block_reduce(tz, false);
    ts = block_reduce(ts, false);
    const float lse = mx + logf(se);
    const float inv_se = 1.0f / se;
    for (int c = threadIdx.x; c < N; c += 256) {
      const float t = soft ? soft[(long long)m * N + c] : (c == lab ? 1.f : 0.f);
      dz[(long long)m * N + c] = (expf(zr[c] - mx) * inv_se * ts - t) * inv_m;
    }
    if (threadIdx.x == 0) {
      const float l = lse * ts - tz;
      if (row_loss) row_loss[m] = l;
      total += l;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) loss[0] = total * inv_m;
}

// out[i] = in[i] * s[0]   (chain rule through the scalar loss; s is a device scalar so the step stays graph-capturable)
__global__ void scale_by_scalar_kernel(const float* __restrict__ in, const float* __restrict__ s, float* __restrict__ out,
                                       long long n) {
  const float f = s[0];
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = in[i] * f;
}

// ------------------------------------------------------------------------------------------------
// Attention probabilities for any sequence length (head dim 64): probs[bp, h, i, :] = softmax_j(q_i . k_j * scale)
// CTA = PR_ROWS query rows of one (batch', head): raw scores of those rows live in shared memory (rows x N fp32),
// keys stream through in tiles of 64; then a row softmax in place and coalesced fp32 stores.
// ------------------------------------------------------------------------------------------------
constexpr int PR_ROWS = 8;
constexpr int PR_KT = 64;

__global__ void __launch_bounds__(256)
attn_probs_kernel(const __nv_bfloat16* __restrict__ qkv, float* __restrict__ probs, int N, int H, float scale) {
  extern __shared__ float psm[];
  float* S = psm;                                   // [PR_ROWS][N]
  float* Q = S + (size_t)PR_ROWS * N;               // [PR_ROWS][64]
  float* Kt = Q + PR_ROWS * 64;                     // [64 keys][65]
  const int bh = blockIdx.y, bp = bh / H, h = bh - bp * H;
  const int i0 = blockIdx.x * PR_ROWS;
  const long long rs = 3LL * H * 64;
  const __nv_bfloat16* base = qkv + (long long)bp * N * rs + h * 64;
  for (int idx = threadIdx.x; idx < PR_ROWS * 64; idx += 256) {
    const int r = idx >> 6, d = idx & 63;
    Q[idx] = (i0 + r < N) ? __bfloat162float(base[(long long)(i0 + r) * rs + d]) * scale : 0.f;
  }
  for (int j0 = 0; j0 < N; j0 += PR_KT) {
    __syncthreads();
    for (int idx = threadIdx.x; idx < PR_KT * 8; idx += 256) {         // 64 keys x 8 vectors of 8 bf16
      const int j = idx >> 3, c = idx & 7;
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if (j0 + j < N) v = *reinterpret_cast<const uint4*>(base + (long long)(j0 + j) * rs + (long long)H * 64 + c * 8);
      float* d = Kt + j * 65 + c * 8;
      const float2 a = unpack_bf16x2(v.x), b2 = unpack_bf16x2(v.y), c2 = unpack_bf16x2(v.z), e2 = unpack_bf16x2(v.w);
      d[0] = a.x; d[1] = a.y; d[2] = b2.x; d[3] = b2.y; d[4] = c2.x; d[5] = c2.y; d[6] = e2.x; d[7] = e2.y;
    }
    __syncthreads();
    // 256 threads = 8 rows x 32 lanes; lane handles keys lane and lane + 32 of the tile
    const int r = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float s0 = 0.f, s1 = 0.f;
#pragma unroll 16
    for (int d = 0; d < 64; ++d) {
      const float q = Q[r * 64 + d];
      s0 = fmaf(q, Kt[lane * 65 + d], s0);
      s1 = fmaf(q, Kt[(lane + 32) * 65 + d], s1);
    }
    if (j0 + lane < N) S[(size_t)r * N + j0 + lane] = s0;
    if (j0 + lane + 32 < N) S[(size_t)r * N + j0 + lane + 32] = s1;
  }
  __syncthreads();
  const int r = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (i0 + r < N) {
    float* row = S + (size_t)r * N;
    float mx = -INFINITY;
    for (int j = lane; j < N; j += 32) mx = fmaxf(mx, row[j]);
    mx = warp_max(mx);
    float se = 0.f;
    for (int j = lane; j < N; j += 32) { const float e = expf(row[j] - mx); row[j] = e; se += e; }
    se = warp_sum(se);
    const float inv = 1.0f / se;
    float* out = probs + ((long long)bh * N + (i0 + r)) * N;
    for (int j = lane; j < N; j += 32) out[j] = row[j] * inv;
  }
}

// ------------------------------------------------------------------------------------------------
// uint8 clip [B, T, H, W, C] -> bf16 patch operand, ToTensor + Normalize fused, with the batch-level Mixup / CutMix of
// reference mixup.py:102-114 folded in: sample b is mixed with sample B-1-b (x.flip(0)).
//   plan = {mode, lam, yl, yh, xl, xh} as floats in device memory (mode 0 none, 1 mixup, 2 cutmix), so a captured graph
//   picks up each step's draw.
// ------------------------------------------------------------------------------------------------
__global__ void im2col_u8_mix_kernel(const uint8_t* __restrict__ x, const float* __restrict__ scale, const float* __restrict__ shift,
                                     const float* __restrict__ plan, __nv_bfloat16* __restrict__ cols, int B, int T, int C,
                                     int H, int W, int tube, int ph, int pw, long long total8) {
  const int Kc = C * tube * ph * pw;
  const int Hp = H / ph, Wp = W / pw, Tp = T / tube;
  const int mode = (int)plan[0];
  const float lam = plan[1];
  const int yl = (int)plan[2], yh = (int)plan[3], xl = (int)plan[4], xh = (int)plan[5];
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total8; idx += (long long)gridDim.x * blockDim.x) {
    const long long e = idx * 8;
    const long long row = e / Kc;
    int k = (int)(e - row * Kc);
    const int j = k % pw; k /= pw;
    const int i = k % ph; k /= ph;
    const int dt = k % tube; const int c = k / tube;
    long long rr = row;
    const int wp = (int)(rr % Wp); rr /= Wp;
    const int hp = (int)(rr % Hp); rr /= Hp;
    const int tp = (int)(rr % Tp); const int b = (int)(rr / Tp);
    const int yy = hp * ph + i, xx0 = wp * pw + j;
    const long long off = ((((long long)(tp * tube + dt)) * H + yy) * W + xx0) * C + c;
    const long long clip = (long long)T * H * W * C;
    const uint8_t* src = x + (long long)b * clip + off;
    const uint8_t* oth = x + (long long)(B - 1 - b) * clip + off;
    const float sc = scale[c], sh = shift[c];
    float v[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float a = fmaf((float)src[(long long)q * C], sc, sh);
      if (mode == 1) {
        const float o = fmaf((float)oth[(long long)q * C], sc, sh);
        v[q] = a * lam + o * (1.0f - lam);                           // x.mul_(lam).add_(x.flip(0).mul_(1 - lam))
      } else if (mode == 2) {
        const bool inside = yy >= yl && yy < yh && (xx0 + q) >= xl && (xx0 + q) < xh;
        v[q] = inside ? fmaf((float)oth[(long long)q * C], sc, sh) : a;   // x[:, :, yl:yh, xl:xh] = x.flip(0)[...]
      } else {
        v[q] = a;
      }
    }
    uint4 o;
    o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]);
    o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
    *reinterpret_cast<uint4*>(cols + e) = o;
  }
}

static int grid_1d(long long n, int threads) {
  long long b = (n + threads - 1) / threads;
  const long long cap = (long long)sm_count() * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace vt

using namespace vt;

extern "C" int vt_linear_small_fwd(const vt_linear_small_params* p, void* stream) {
  VT_REQUIRE(p && p->x && p->w && p->y, "vt_linear_small_fwd: null pointer");
  VT_REQUIRE(p->M > 0 && p->N > 0 && p->K > 0 && p->K % 4 == 0, "vt_linear_small_fwd: bad shape M=%d N=%d K=%d (K %% 4 == 0)", p->M, p->N, p->K);
  VT_REQUIRE(p->M <= 4096, "vt_linear_small_fwd: M=%d is not skinny (use vt_gemm)", p->M);
  dim3 grid((p->N + 7) / 8, (p->M + LS_ROWS - 1) / LS_ROWS);
  linear_small_fwd_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(p->x, p->w, p->b, p->y, p->M, p->N, p->K);
  return check_launch("linear_small_fwd_kernel");
}

extern "C" int vt_linear_small_bwd(const vt_linear_small_bwd_params* p, void* stream) {
  VT_REQUIRE(p && p->dy && p->x && p->w, "vt_linear_small_bwd: null pointer");
  VT_REQUIRE(p->M > 0 && p->N > 0 && p->K > 0 && p->K % 4 == 0, "vt_linear_small_bwd: bad shape");
  VT_REQUIRE(p->M <= 4096, "vt_linear_small_bwd: M=%d is not skinny", p->M);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (p->dw) {
    linear_small_wgrad_kernel<<<(p->N + 7) / 8, 256, 0, st>>>(p->dy, p->x, p->dw, p->db, p->M, p->N, p->K);
    int rc = check_launch("linear_small_wgrad_kernel");
    if (rc) return rc;
  }
  if (p->dx) {
    dim3 grid((p->K + 127) / 128, (p->M + LS_ROWS - 1) / LS_ROWS);
    linear_small_dgrad_kernel<<<grid, 256, 0, st>>>(p->dy, p->w, p->dx, p->M, p->N, p->K);
    return check_launch("linear_small_dgrad_kernel");
  }
  return 0;
}

extern "C" int vt_softmax_ce(const vt_softmax_ce_params* p, void* stream) {
  VT_REQUIRE(p && p->logits && p->loss && p->dlogits, "vt_softmax_ce: null pointer");
  VT_REQUIRE((p->labels != nullptr) != (p->soft_targets != nullptr), "vt_softmax_ce: give labels or soft_targets (exactly one)");
  VT_REQUIRE(p->M > 0 && p->M <= 4096 && p->N > 0, "vt_softmax_ce: bad shape M=%d N=%d", p->M, p->N);
  softmax_ce_kernel<<<1, 256, 0, static_cast<cudaStream_t>(stream)>>>(p->logits, p->labels, p->soft_targets, p->loss,
                                                                       p->row_loss, p->dlogits, p->M, p->N);
  return check_launch("softmax_ce_kernel");
}

extern "C" int vt_scale_by_scalar(const vt_scale_params* p, void* stream) {
  VT_REQUIRE(p && p->in && p->scalar && p->out && p->n > 0, "vt_scale_by_scalar: bad params");
  scale_by_scalar_kernel<<<grid_1d(p->n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(p->in, p->scalar, p->out, p->n);
  return check_launch("scale_by_scalar_kernel");
}

extern "C" int vt_attn_probs(const vt_attn_probs_params* p, void* stream) {
  VT_REQUIRE(p && p->qkv && p->probs, "vt_attn_probs: null pointer");
  VT_REQUIRE(p->hd == 64, "vt_attn_probs: head dim %d unsupported (64 only)", p->hd);
  VT_REQUIRE(p->Bp > 0 && p->H > 0 && p->N > 0, "vt_attn_probs: bad shape");
  const size_t smem = ((size_t)PR_ROWS * p->N + PR_ROWS * 64 + PR_KT * 65) * sizeof(float);
  VT_REQUIRE(smem <= 200 * 1024, "vt_attn_probs: N=%d too long (%zu bytes of shared memory)", p->N, smem);
  static size_t max_set = 48 * 1024;
  if (smem > max_set) {
    cudaError_t e = cudaFuncSetAttribute(attn_probs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    VT_REQUIRE(e == cudaSuccess, "vt_attn_probs: smem attribute: %s", cudaGetErrorString(e));
    max_set = 200 * 1024;
  }
  dim3 grid((p->N + PR_ROWS - 1) / PR_ROWS, p->Bp * p->H);
  attn_probs_kernel<<<grid, 256, smem, static_cast<cudaStream_t>(stream)>>>(static_cast<const __nv_bfloat16*>(p->qkv), p->probs,
                                                                             p->N, p->H, p->scale);
  return check_launch("attn_probs_kernel");
}

extern "C" int vt_im2col_u8_mix_bf16(const vt_im2col_u8_mix_params* p, void* stream) {
  VT_REQUIRE(p && p->x && p->scale && p->shift && p->cols && p->plan, "vt_im2col_u8_mix_bf16: null pointer");
  VT_REQUIRE(p->pw % 8 == 0 && p->W % p->pw == 0 && p->H % p->ph == 0 && p->T % p->tube == 0, "vt_im2col_u8_mix_bf16: unsupported geometry");
  const long long total8 = (long long)p->B * p->T * p->C * p->H * p->W / 8;
  im2col_u8_mix_kernel<<<grid_1d(total8, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      p->x, p->scale, p->shift, p->plan, static_cast<__nv_bfloat16*>(p->cols), p->B, p->T, p->C, p->H, p->W, p->tube, p->ph,
      p->pw, total8);
  return check_launch("im2col_u8_mix_kernel");
}

// ------------------------------------------------------------------------------------------------
// cls rows of the divided space-time blocks: dst[b, :] = src[b, :] + scale * sum_t extra[b, t, :]
// (extra == NULL: plain row copy).  One launch for what autograd spells as mean / sum + add + strided copy
// (transformer.py:282-283 cls passthrough of the temporal block, :371-377 mean over the per-frame cls replicas).
// ------------------------------------------------------------------------------------------------
namespace vt {
__global__ void cls_rows_kernel(const float* __restrict__ src, long long src_stride, const float* __restrict__ extra,
                                long long extra_bs, int T, float scale, float* __restrict__ dst, long long dst_stride, int B, int D4) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < B * D4; i += gridDim.x * blockDim.x) {
    const int b = i / D4, j = i - b * D4;
    float4 v = reinterpret_cast<const float4*>(src + (long long)b * src_stride)[j];
    if (extra) {
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int t = 0; t < T; ++t) {
        const float4 e = reinterpret_cast<const float4*>(extra + (long long)b * extra_bs + (long long)t * D4 * 4)[j];
        a.x += e.x; a.y += e.y; a.z += e.z; a.w += e.w;
      }
      v.x = fmaf(scale, a.x, v.x); v.y = fmaf(scale, a.y, v.y); v.z = fmaf(scale, a.z, v.z); v.w = fmaf(scale, a.w, v.w);
    }
    reinterpret_cast<float4*>(dst + (long long)b * dst_stride)[j] = v;
  }
}
}  // namespace vt

extern "C" int vt_cls_rows(const vt_cls_rows_params* p, void* stream) {
  using namespace vt;
  VT_REQUIRE(p && p->src && p->dst && p->B > 0 && p->D > 0, "vt_cls_rows: bad params");
  VT_REQUIRE(p->D % 4 == 0 && p->src_stride % 4 == 0 && p->dst_stride % 4 == 0 && (!p->extra || (p->extra_bs % 4 == 0 && p->T > 0)),
             "vt_cls_rows: D and strides must be multiples of 4");
  VT_REQUIRE(((reinterpret_cast<uintptr_t>(p->src) | reinterpret_cast<uintptr_t>(p->dst) | reinterpret_cast<uintptr_t>(p->extra)) & 15) == 0,
             "vt_cls_rows: pointers must be 16-byte aligned");
  const int n = p->B * (p->D / 4);
  int blocks = (n + 255) / 256;
  if (blocks > 1024) blocks = 1024;
  cls_rows_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(p->src, p->src_stride, p->extra, p->extra_bs, p->T, p->scale,
                                                                        p->dst, p->dst_stride, p->B, p->D / 4);
  return check_launch("cls_rows_kernel");
}

