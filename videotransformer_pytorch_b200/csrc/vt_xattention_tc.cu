// Pooling attention (MViT: separate strided Q / K / V, Nq != Nk, head dim 96) on tcgen05 tensor cores.
//
// Head dim 96 = 1.5 swizzle atoms.  Every operand tile is staged as 128 rows x 128 columns (two 128-byte-swizzled
// K-major blocks of 64 columns, two TMA boxes); contractions over the head dim issue 6 of the 8 K-steps, and MMAs whose
// N dimension is the head dim run with N = 128 and leave 32 unread garbage columns in TMEM.  No partial atoms anywhere.
//
//   forward : CTA = 128 queries of one (b, h); K/V tiles of 128 keys stream through a 2-stage TMA ring.  Two softmax
//             warpgroups take alternate key tiles (own S buffer, own P buffer, own P.V scratch in TMEM):
//             S = Q K^T -> row max / exp2 / row sum, P (bf16) to swizzled smem -> PV = P V (fresh accumulator) ->
//             running O (registers) = O * corr + PV.  The two partial (m, l, O) states are merged at the end.
//   dQ      : CTA = 128 queries; per key tile S = Q K^T, dP = dO V^T (TMEM) -> dS = P (dP - delta) scale (bf16, smem)
//             -> dQ += dS K (TMEM accumulator).  Also writes delta = rowsum(dO * O) for the dK/dV kernel.
//   dK/dV   : CTA = 128 keys x a chunk of query tiles; per query tile S, dP -> P, dS (smem) -> dK += dS^T Q,
//             dV += P^T dO (P/dS read transposed through MN-major descriptors); fp32 atomics merge the chunks.
#include "vt_common.cuh"
#include "vt_umma.cuh"

namespace vt {

int make_tmap_bf16_2d(CUtensorMap* map, const void* base, long long rows, long long cols, long long ld, int box_rows);

constexpr int XT_BLK = 128 * 128;      // one swizzled block: 128 rows x 64 bf16 = 16 KiB
constexpr int XT_TILE = 2 * XT_BLK;    // 128 rows x 128 (padded) columns
constexpr float XT_LOG2E = 1.4426950408889634f;
constexpr float XT_LN2 = 0.6931471805599453f;
constexpr int XT_THREADS = 320;        // warps 0-3 / 4-7: two warpgroups; warp 8: MMA issuer; warp 9: TMA producer
constexpr int XT_QTILES_PER_CHUNK = 16;

// addressing of one operand: tile rows start at b*rb + h*rh + n0, columns at h*ch
struct XtOp {
  int rb, rh, ch;
};

__device__ __forceinline__ void xt_st_sw128(uint8_t* block, int r, int g, uint4 v) {
  *reinterpret_cast<uint4*>(block + r * 128 + ((g ^ (r & 7)) << 4)) = v;
}
__device__ __forceinline__ void xt_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
// one 128 x 128 tile = two boxes of 64 columns
__device__ __forceinline__ void xt_load_tile(uint8_t* dst, const CUtensorMap* map, uint64_t* bar, int col, int row) {
  tma_load_2d(dst, map, bar, col, row);
  tma_load_2d(dst + XT_BLK, map, bar, col + 64, row);
}
// D[128 x 128] (+)= A[128 x HD] B[128 x HD]^T, both K-major tiles (contraction over the head dim: HD/16 K-steps)
template <int HD>
__device__ __forceinline__ void xt_mma_hd(uint32_t d_tmem, uint32_t a_addr, uint32_t b_addr, uint32_t idesc) {
#pragma unroll
  for (int k = 0; k < HD / 16; ++k)
    umma_bf16_ss(d_tmem, sdesc_kmajor(a_addr + (k >> 2) * XT_BLK + (k & 3) * 32),
                 sdesc_kmajor(b_addr + (k >> 2) * XT_BLK + (k & 3) * 32), idesc, k > 0);
}
__device__ __forceinline__ uint4 xt_pack8(const float* e) {
  uint4 o;
  o.x = pack_bf16x2(e[0], e[1]);
  o.y = pack_bf16x2(e[2], e[3]);
  o.z = pack_bf16x2(e[4], e[5]);
  o.w = pack_bf16x2(e[6], e[7]);
  return o;
}
template <int HD>
__device__ __forceinline__ void xt_store_row(__nv_bfloat16* dst, const float* v, float mul) {
  uint4* d = reinterpret_cast<uint4*>(dst);
#pragma unroll
  for (int g = 0; g < HD / 8; ++g) {
    uint4 o;
    o.x = pack_bf16x2(v[g * 8 + 0] * mul, v[g * 8 + 1] * mul);
    o.y = pack_bf16x2(v[g * 8 + 2] * mul, v[g * 8 + 3] * mul);
    o.z = pack_bf16x2(v[g * 8 + 4] * mul, v[g * 8 + 5] * mul);
    o.w = pack_bf16x2(v[g * 8 + 6] * mul, v[g * 8 + 7] * mul);
    d[g] = o;
  }
}

// ================================================================================================
// forward
// ================================================================================================
struct XtFwd {
  __nv_bfloat16* o;
  float* lse;
  long long o_bs, o_hs, o_rs;
  XtOp q, k, v;
  int H, Nq, Nk, nkt;
  float scale;
};

template <int XT_HD>
__global__ void __launch_bounds__(XT_THREADS, 1)
xattn_tc_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmV, const XtFwd p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + XT_TILE;           // [2]
  uint8_t* sV = sK + 2 * XT_TILE;       // [2]
  uint8_t* sP = sV + 2 * XT_TILE;       // [2]; reused as the merge stash at the end
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * XT_TILE);
  uint64_t* bar_q = bars + 0;
  uint64_t* bar_kv_full = bars + 1;     // [2]
  uint64_t* bar_kv_free = bars + 3;     // [2]
  uint64_t* bar_s = bars + 5;           // [2]
  uint64_t* bar_p = bars + 7;           // [2], 128 arrivals
  uint64_t* bar_pv = bars + 9;          // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 11);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = blockIdx.x, bh = blockIdx.y, b = bh / p.H, h = bh - b * p.H;

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(bar_q, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bar_kv_full[s], 1);
      mbar_init(&bar_kv_free[s], 1);
      mbar_init(&bar_s[s], 1);
      mbar_init(&bar_p[s], 128);
      mbar_init(&bar_pv[s], 1);
    }
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // TMEM columns: S[s] at s*128, PV[s] at 256 + s*128

  if (warp == 9) {
    if (lane == 0) {
      mbar_arrive_expect_tx(bar_q, XT_TILE);
      xt_load_tile(sQ, &tmQ, bar_q, h * p.q.ch, b * p.q.rb + h * p.q.rh + qt * 128);
      for (int j = 0; j < p.nkt; ++j) {
        const int s = j & 1;
        if (j >= 2) mbar_wait(&bar_kv_free[s], ((j >> 1) - 1) & 1);
        mbar_arrive_expect_tx(&bar_kv_full[s], 2 * XT_TILE);
        xt_load_tile(sK + s * XT_TILE, &tmK, &bar_kv_full[s], h * p.k.ch, b * p.k.rb + h * p.k.rh + j * 128);
        xt_load_tile(sV + s * XT_TILE, &tmV, &bar_kv_full[s], h * p.v.ch, b * p.v.rb + h * p.v.rh + j * 128);
      }
    }
  } else if (warp == 8) {
    if (lane == 0) {
      const uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0);
      const uint32_t idesc_o = make_idesc_bf16(128, 128, 0, 1);
      const uint32_t qa = smem_u32(sQ);
      mbar_wait(bar_q, 0);
      tc_fence_after();
      auto issue_qk = [&](int j) {
        const int s = j & 1;
        mbar_wait(&bar_kv_full[s], (j >> 1) & 1);
        tc_fence_after();
        xt_mma_hd<XT_HD>(tmem_base + s * 128, qa, smem_u32(sK + s * XT_TILE), idesc_s);
        umma_commit(&bar_s[s]);
      };
      issue_qk(0);
      if (p.nkt > 1) issue_qk(1);
      for (int j = 0; j < p.nkt; ++j) {
        const int s = j & 1;
        mbar_wait(&bar_p[s], (j >> 1) & 1);
        tc_fence_after();
        const uint32_t pa = smem_u32(sP + s * XT_TILE), va = smem_u32(sV + s * XT_TILE);
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)
          umma_bf16_ss(tmem_base + 256 + s * 128, sdesc_kmajor(pa + (ks >> 2) * XT_BLK + (ks & 3) * 32),
                       sdesc_mnmajor(va + ks * 2048, XT_BLK), idesc_o, ks > 0);
        umma_commit(&bar_pv[s]);
        umma_commit(&bar_kv_free[s]);
        // S(j+1) goes out after P.V(j): its K tile (stage (j+1)&1) was released by P.V(j-1), so the loads of tiles
        // j+1 and j+2 are both in flight while this thread waits
        if (j >= 1 && j + 1 < p.nkt) issue_qk(j + 1);
      }
    }
  } else {
    const int wg = warp >> 2;
    const int r = (warp & 3) * 32 + lane;                      // query row in the tile == TMEM lane
    const uint32_t tlane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    const float sl2 = p.scale * XT_LOG2E;
    float m = -INFINITY, l = 0.f;
    float acc[XT_HD];
#pragma unroll
    for (int d = 0; d < XT_HD; ++d) acc[d] = 0.f;
    uint8_t* Pt = sP + wg * XT_TILE;
    for (int j = wg; j < p.nkt; j += 2) {
      const uint32_t ph = (uint32_t)((j >> 1) & 1);
      const int nvalid = min(128, p.Nk - j * 128);
      mbar_wait(&bar_s[wg], ph);
      tc_fence_after();
      float mx = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        tmem_ld32(tlane + wg * 128 + c * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int jj = 0; jj < 32; ++jj)
          if (c * 32 + jj < nvalid) mx = fmaxf(mx, __uint_as_float(v[jj]));
      }
      const float mn = fmaxf(m, mx * sl2);
      const float corr = fast_exp2(m - mn);
      float rowsum = 0.f;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        tmem_ld32(tlane + wg * 128 + c * 32, v);
        tmem_ld_wait();
        float e[32];
#pragma unroll
        for (int jj = 0; jj < 32; ++jj) {
          e[jj] = (c * 32 + jj < nvalid) ? fast_exp2(fmaf(__uint_as_float(v[jj]), sl2, -mn)) : 0.f;
          rowsum += e[jj];
        }
        uint8_t* blk = Pt + (c >> 1) * XT_BLK;
#pragma unroll
        for (int g = 0; g < 4; ++g) xt_st_sw128(blk, r, (c & 1) * 4 + g, xt_pack8(e + g * 8));
      }
      l = l * corr + rowsum;
      m = mn;
      tc_fence_before();
      fence_proxy_async();
      mbar_arrive(&bar_p[wg]);
#pragma unroll
      for (int d = 0; d < XT_HD; ++d) acc[d] *= corr;
      mbar_wait(&bar_pv[wg], ph);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < XT_HD / 32; ++c) {
        uint32_t v[32];
        tmem_ld32(tlane + 256 + wg * 128 + c * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int jj = 0; jj < 32; ++jj) acc[c * 32 + jj] += __uint_as_float(v[jj]);
      }
    }
    // merge the two warpgroups' partial softmax states (all P.V MMAs have completed: each group waited for its last)
    tc_fence_before();
    xt_bar_sync(1, 256);
    constexpr int PITCH = XT_HD + 1;                           // odd pitch: conflict-free row-per-thread access
    float* stash = reinterpret_cast<float*>(sP);               // [128][PITCH] O rows, then m[128], l[128]
    float* stash_m = stash + 128 * PITCH;
    float* stash_l = stash_m + 128;
    if (wg == 1) {
#pragma unroll
      for (int d = 0; d < XT_HD; ++d) stash[r * PITCH + d] = acc[d];
      stash_m[r] = m;
      stash_l[r] = l;
    }
    xt_bar_sync(1, 256);
    if (wg == 0) {
      const float m1 = stash_m[r], l1 = stash_l[r];
      const float mm = fmaxf(m, m1);
      const float f0 = fast_exp2(m - mm), f1 = fast_exp2(m1 - mm);   // m1 = -inf when the other group had no tile
      const float lt = l * f0 + l1 * f1;
      const int q = qt * 128 + r;
      if (q < p.Nq) {
#pragma unroll
        for (int d = 0; d < XT_HD; ++d) acc[d] = acc[d] * f0 + stash[r * PITCH + d] * f1;
        xt_store_row<XT_HD>(p.o + (long long)b * p.o_bs + (long long)h * p.o_hs + (long long)q * p.o_rs, acc, 1.0f / lt);
        p.lse[(long long)bh * p.Nq + q] = (mm + log2f(lt)) * XT_LN2;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// ================================================================================================
// dQ (+ delta)
// ================================================================================================
struct XtDq {
  const __nv_bfloat16* o;
  const __nv_bfloat16* dout;
  const float* lse;
  float* delta;
  __nv_bfloat16* dq;
  long long o_bs, o_hs, o_rs, dq_bs, dq_hs, dq_rs;
  XtOp q, k, v, d;      // d: dout as a TMA operand
  int H, Nq, Nk, nkt;
  float scale;
};

template <int XT_HD>
__global__ void __launch_bounds__(XT_THREADS, 1)
xattn_tc_dq_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmD, const XtDq p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sDO = sQ + XT_TILE;
  uint8_t* sK = sDO + XT_TILE;          // [2]
  uint8_t* sV = sK + 2 * XT_TILE;       // [2]
  uint8_t* sDS = sV + 2 * XT_TILE;
  float* lse_s = reinterpret_cast<float*>(sDS + XT_TILE);   // [128]
  float* del_s = lse_s + 128;                               // [128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(del_s + 128);
  uint64_t* bar_qdo = bars + 0;
  uint64_t* bar_kv_full = bars + 1;     // [2]
  uint64_t* bar_kv_free = bars + 3;     // [2]
  uint64_t* bar_sdp = bars + 5;
  uint64_t* bar_ds = bars + 6;          // 256 arrivals
  uint64_t* bar_dq = bars + 7;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  constexpr uint32_t COL_S = 0, COL_DP = 128, COL_DQ = 256;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = blockIdx.x, bh = blockIdx.y, b = bh / p.H, h = bh - b * p.H;

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmD);
    mbar_init(bar_qdo, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bar_kv_full[s], 1);
      mbar_init(&bar_kv_free[s], 1);
    }
    mbar_init(bar_sdp, 1);
    mbar_init(bar_ds, 256);
    mbar_init(bar_dq, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 9) {
    if (lane == 0) {
      mbar_arrive_expect_tx(bar_qdo, 2 * XT_TILE);
      xt_load_tile(sQ, &tmQ, bar_qdo, h * p.q.ch, b * p.q.rb + h * p.q.rh + qt * 128);
      xt_load_tile(sDO, &tmD, bar_qdo, h * p.d.ch, b * p.d.rb + h * p.d.rh + qt * 128);
      for (int j = 0; j < p.nkt; ++j) {
        const int s = j & 1;
        if (j >= 2) mbar_wait(&bar_kv_free[s], ((j >> 1) - 1) & 1);
        mbar_arrive_expect_tx(&bar_kv_full[s], 2 * XT_TILE);
        xt_load_tile(sK + s * XT_TILE, &tmK, &bar_kv_full[s], h * p.k.ch, b * p.k.rb + h * p.k.rh + j * 128);
        xt_load_tile(sV + s * XT_TILE, &tmV, &bar_kv_full[s], h * p.v.ch, b * p.v.rb + h * p.v.rh + j * 128);
      }
    }
  } else if (warp == 8) {
    if (lane == 0) {
      const uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0);
      const uint32_t idesc_q = make_idesc_bf16(128, 128, 0, 1);
      const uint32_t qa = smem_u32(sQ), doa = smem_u32(sDO), dsa = smem_u32(sDS);
      mbar_wait(bar_qdo, 0);
      tc_fence_after();
      for (int j = 0; j < p.nkt; ++j) {
        const int s = j & 1;
        const uint32_t ka = smem_u32(sK + s * XT_TILE), va = smem_u32(sV + s * XT_TILE);
        mbar_wait(&bar_kv_full[s], (j >> 1) & 1);
        tc_fence_after();
        xt_mma_hd<XT_HD>(tmem_base + COL_S, qa, ka, idesc_s);
        xt_mma_hd<XT_HD>(tmem_base + COL_DP, doa, va, idesc_s);
        umma_commit(bar_sdp);
        mbar_wait(bar_ds, j & 1);
        tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)
          umma_bf16_ss(tmem_base + COL_DQ, sdesc_kmajor(dsa + (ks >> 2) * XT_BLK + (ks & 3) * 32),
                       sdesc_mnmajor(ka + ks * 2048, XT_BLK), idesc_q, (j > 0 || ks > 0) ? 1u : 0u);
        umma_commit(&bar_kv_free[s]);
      }
      umma_commit(bar_dq);
    }
  } else {
    const int half = warp >> 2;
    const int r = (warp & 3) * 32 + lane;
    const uint32_t tlane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    const int q = qt * 128 + r;
    const bool qok = q < p.Nq;
    if (half == 0) {     // delta_i = dO_i . O_i, log-sum-exp in the log2 domain (+inf for rows past Nq => P = 0)
      float dl = 0.f, ls = INFINITY;
      if (qok) {
        const long long off = (long long)b * p.o_bs + (long long)h * p.o_hs + (long long)q * p.o_rs;
        const uint4* o4 = reinterpret_cast<const uint4*>(p.o + off);
        const uint4* g4 = reinterpret_cast<const uint4*>(p.dout + off);
#pragma unroll
        for (int i = 0; i < XT_HD / 8; ++i) {
          const uint4 a = o4[i], g = g4[i];
          const float2 a0 = unpack_bf16x2(a.x), a1 = unpack_bf16x2(a.y), a2 = unpack_bf16x2(a.z), a3 = unpack_bf16x2(a.w);
          const float2 g0 = unpack_bf16x2(g.x), g1 = unpack_bf16x2(g.y), g2 = unpack_bf16x2(g.z), g3 = unpack_bf16x2(g.w);
          dl += a0.x * g0.x + a0.y * g0.y + a1.x * g1.x + a1.y * g1.y + a2.x * g2.x + a2.y * g2.y + a3.x * g3.x + a3.y * g3.y;
        }
        ls = p.lse[(long long)bh * p.Nq + q] * XT_LOG2E;
        p.delta[(long long)bh * p.Nq + q] = dl;
      }
      del_s[r] = dl;
      lse_s[r] = ls;
    }
    xt_bar_sync(1, 256);
    const float lq = lse_s[r], dq_ = del_s[r];
    const float sl2 = p.scale * XT_LOG2E;
    for (int j = 0; j < p.nkt; ++j) {
      mbar_wait(bar_sdp, j & 1);
      tc_fence_after();
#pragma unroll 1
      for (int cc = 0; cc < 2; ++cc) {
        const int c0 = half * 64 + cc * 32;
        uint32_t sv[32], dv[32];
        tmem_ld32(tlane + COL_S + c0, sv);
        tmem_ld32(tlane + COL_DP + c0, dv);
        tmem_ld_wait();
        float ds[32];
#pragma unroll
        for (int jj = 0; jj < 32; ++jj) {
          const bool ok = j * 128 + c0 + jj < p.Nk;
          const float pj = ok ? fast_exp2(fmaf(__uint_as_float(sv[jj]), sl2, -lq)) : 0.f;
          ds[jj] = pj * (__uint_as_float(dv[jj]) - dq_) * p.scale;
        }
        uint8_t* blk = sDS + half * XT_BLK;
#pragma unroll
        for (int g = 0; g < 4; ++g) xt_st_sw128(blk, r, cc * 4 + g, xt_pack8(ds + g * 8));
      }
      tc_fence_before();
      fence_proxy_async();
      mbar_arrive(bar_ds);
    }
    mbar_wait(bar_dq, 0);
    tc_fence_after();
    if (half == 0) {
      float out[XT_HD];
#pragma unroll
      for (int c = 0; c < XT_HD / 32; ++c) {
        uint32_t v[32];
        tmem_ld32(tlane + COL_DQ + c * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int jj = 0; jj < 32; ++jj) out[c * 32 + jj] = __uint_as_float(v[jj]);
      }
      if (qok)
        xt_store_row<XT_HD>(p.dq + (long long)b * p.dq_bs + (long long)h * p.dq_hs + (long long)q * p.dq_rs, out, 1.0f);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// ================================================================================================
// dK / dV
// ================================================================================================
struct XtDkv {
  const float* lse;
  const float* delta;
  float* dk;
  float* dv;
  XtOp q, k, v, d;
  int H, Nq, Nk, nqt;
  float scale;
};

template <int XT_HD>
__global__ void __launch_bounds__(XT_THREADS, 1)
xattn_tc_dkv_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmD, const XtDkv p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sK = smem;
  uint8_t* sV = sK + XT_TILE;
  uint8_t* sQ = sV + XT_TILE;
  uint8_t* sDO = sQ + XT_TILE;
  uint8_t* sP = sDO + XT_TILE;
  uint8_t* sDS = sP + XT_TILE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sDS + XT_TILE);
  uint64_t* bar_kv = bars + 0;
  uint64_t* bar_qdo_full = bars + 1;
  uint64_t* bar_qdo_free = bars + 2;
  uint64_t* bar_sdp = bars + 3;
  uint64_t* bar_pds = bars + 4;         // 256 arrivals
  uint64_t* bar_done = bars + 5;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6);
  constexpr uint32_t COL_S = 0, COL_DP = 128, COL_DK = 256, COL_DV = 384;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kt = blockIdx.x, bh = blockIdx.z, b = bh / p.H, h = bh - b * p.H;
  const int qt0 = blockIdx.y * XT_QTILES_PER_CHUNK;
  const int n_it = min(XT_QTILES_PER_CHUNK, p.nqt - qt0);

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmD);
    mbar_init(bar_kv, 1);
    mbar_init(bar_qdo_full, 1);
    mbar_init(bar_qdo_free, 1);
    mbar_init(bar_sdp, 1);
    mbar_init(bar_pds, 256);
    mbar_init(bar_done, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 9) {
    if (lane == 0) {
      mbar_arrive_expect_tx(bar_kv, 2 * XT_TILE);
      xt_load_tile(sK, &tmK, bar_kv, h * p.k.ch, b * p.k.rb + h * p.k.rh + kt * 128);
      xt_load_tile(sV, &tmV, bar_kv, h * p.v.ch, b * p.v.rb + h * p.v.rh + kt * 128);
      for (int it = 0; it < n_it; ++it) {
        if (it > 0) mbar_wait(bar_qdo_free, (it - 1) & 1);
        mbar_arrive_expect_tx(bar_qdo_full, 2 * XT_TILE);
        xt_load_tile(sQ, &tmQ, bar_qdo_full, h * p.q.ch, b * p.q.rb + h * p.q.rh + (qt0 + it) * 128);
        xt_load_tile(sDO, &tmD, bar_qdo_full, h * p.d.ch, b * p.d.rb + h * p.d.rh + (qt0 + it) * 128);
      }
    }
  } else if (warp == 8) {
    if (lane == 0) {
      const uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0);
      const uint32_t idesc_t = make_idesc_bf16(128, 128, 1, 1);
      const uint32_t qa = smem_u32(sQ), doa = smem_u32(sDO), ka = smem_u32(sK), va = smem_u32(sV);
      const uint32_t pa = smem_u32(sP), dsa = smem_u32(sDS);
      mbar_wait(bar_kv, 0);
      for (int it = 0; it < n_it; ++it) {
        mbar_wait(bar_qdo_full, it & 1);
        tc_fence_after();
        xt_mma_hd<XT_HD>(tmem_base + COL_S, qa, ka, idesc_s);       // S [query x key]
        xt_mma_hd<XT_HD>(tmem_base + COL_DP, doa, va, idesc_s);     // dP
        umma_commit(bar_sdp);
        mbar_wait(bar_pds, it & 1);
        tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)     // dK[key x hd] += dS^T Q : contraction over the 128 queries of the tile
          umma_bf16_ss(tmem_base + COL_DK, sdesc_mnmajor(dsa + ks * 2048, XT_BLK), sdesc_mnmajor(qa + ks * 2048, XT_BLK),
                       idesc_t, (it > 0 || ks > 0) ? 1u : 0u);
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)     // dV += P^T dO
          umma_bf16_ss(tmem_base + COL_DV, sdesc_mnmajor(pa + ks * 2048, XT_BLK), sdesc_mnmajor(doa + ks * 2048, XT_BLK),
                       idesc_t, (it > 0 || ks > 0) ? 1u : 0u);
        umma_commit(bar_qdo_free);
      }
      umma_commit(bar_done);
    }
  } else {
    const int half = warp >> 2;
    const int r = (warp & 3) * 32 + lane;
    const uint32_t tlane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    const float sl2 = p.scale * XT_LOG2E;
    for (int it = 0; it < n_it; ++it) {
      const int q = (qt0 + it) * 128 + r;
      const bool qok = q < p.Nq;
      const float lq = qok ? p.lse[(long long)bh * p.Nq + q] * XT_LOG2E : INFINITY;
      const float dl = qok ? p.delta[(long long)bh * p.Nq + q] : 0.f;
      mbar_wait(bar_sdp, it & 1);
      tc_fence_after();
#pragma unroll 1
      for (int cc = 0; cc < 2; ++cc) {
        const int c0 = half * 64 + cc * 32;
        uint32_t sv[32], dv[32];
        tmem_ld32(tlane + COL_S + c0, sv);
        tmem_ld32(tlane + COL_DP + c0, dv);
        tmem_ld_wait();
        float pv[32], ds[32];
#pragma unroll
        for (int jj = 0; jj < 32; ++jj) {
          const bool ok = kt * 128 + c0 + jj < p.Nk;
          const float pj = ok ? fast_exp2(fmaf(__uint_as_float(sv[jj]), sl2, -lq)) : 0.f;
          pv[jj] = pj;
          ds[jj] = pj * (__uint_as_float(dv[jj]) - dl) * p.scale;
        }
        uint8_t* pb = sP + half * XT_BLK;
        uint8_t* db = sDS + half * XT_BLK;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          xt_st_sw128(pb, r, cc * 4 + g, xt_pack8(pv + g * 8));
          xt_st_sw128(db, r, cc * 4 + g, xt_pack8(ds + g * 8));
        }
      }
      tc_fence_before();
      fence_proxy_async();
      mbar_arrive(bar_pds);
    }
    mbar_wait(bar_done, 0);
    tc_fence_after();
    const int key = kt * 128 + r;                       // accumulator row == key
    float* dst = (half == 0 ? p.dk : p.dv) + ((long long)bh * p.Nk + key) * XT_HD;
    const uint32_t col = half == 0 ? COL_DK : COL_DV;
#pragma unroll 1
    for (int c = 0; c < XT_HD / 32; ++c) {
      uint32_t v[32];
      tmem_ld32(tlane + col + c * 32, v);
      tmem_ld_wait();
      if (key < p.Nk) {
#pragma unroll
        for (int jj = 0; jj < 32; ++jj) atomicAdd(dst + c * 32 + jj, __uint_as_float(v[jj]));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// ================================================================================================
// host side
// ================================================================================================
// Classify an operand addressed as (b, h, n, c) -> b*bs + h*hs + n*rs + c into a 2-D TMA view.
//   token-major  (hs == hd):           rows = B*N, row = b*N + n, col = h*hd + c, ld = rs   (needs bs == N*rs)
//   head-major   (rs == hd, contiguous): rows = B*H*N, row = (b*H + h)*N + n, col = c, ld = hd
static bool xt_classify(long long bs, long long hs, long long rs, int B, int H, int N, int hd, XtOp* op, long long* rows,
                        long long* cols, long long* ld) {
  if (hs == hd && bs == (long long)N * rs && rs >= (long long)H * hd) {
    *op = XtOp{N, 0, hd};
    *rows = (long long)B * N;
    *cols = (long long)H * hd;
    *ld = rs;
    return true;
  }
  if (rs == hd && hs == (long long)N * hd && bs == (long long)H * N * hd) {
    *op = XtOp{H * N, N, 0};
    *rows = (long long)B * H * N;
    *cols = hd;
    *ld = hd;
    return true;
  }
  return false;
}

static int xt_map(CUtensorMap* m, const void* base, long long bs, long long hs, long long rs, int B, int H, int N, int hd, XtOp* op) {
  long long rows, cols, ld;
  if (!xt_classify(bs, hs, rs, B, H, N, hd, op, &rows, &cols, &ld)) return -1;
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (ld * 2) % 16 != 0) return -1;
  return make_tmap_bf16_2d(m, base, rows, cols, ld, 128);
}

bool xattn_tc_supported(const void* q, long long q_bs, long long q_hs, long long q_rs, const void* k, long long k_bs,
                        long long k_hs, long long k_rs, const void* v, long long v_bs, long long v_hs, long long v_rs, int B,
                        int H, int Nq, int Nk, int hd) {
  if (hd != 96 && hd != 64) return false;
  XtOp op;
  long long a, b2, c;
  auto ok = [&](const void* p, long long bs, long long hs, long long rs, int N) {
    return (reinterpret_cast<uintptr_t>(p) & 15) == 0 && xt_classify(bs, hs, rs, B, H, N, hd, &op, &a, &b2, &c) && (c * 2) % 16 == 0;
  };
  return ok(q, q_bs, q_hs, q_rs, Nq) && ok(k, k_bs, k_hs, k_rs, Nk) && ok(v, v_bs, v_hs, v_rs, Nk);
}

constexpr int XT_SMEM_FWD = 7 * XT_TILE + 256 + 1024;                // Q + 2K + 2V + 2P, barriers, alignment slack
constexpr int XT_SMEM_DQ = 7 * XT_TILE + 2 * 128 * 4 + 256 + 1024;   // Q + dO + 2K + 2V + dS, lse/delta
constexpr int XT_SMEM_DKV = 6 * XT_TILE + 256 + 1024;                // K + V + Q + dO + P + dS

template <typename Kern>
static int xt_set_smem(Kern kern, int bytes, bool* done, const char* what) {
  if (!*done) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    VT_REQUIRE(e == cudaSuccess, "%s: smem attribute (%d bytes): %s", what, bytes, cudaGetErrorString(e));
    *done = true;
  }
  return 0;
}

int xattn_tc_fwd_launch(const vt_xattn_fwd_params* q, cudaStream_t st) {
  static_assert(XT_SMEM_FWD <= 232448 && XT_SMEM_DQ <= 232448 && XT_SMEM_DKV <= 232448, "shared memory budget");
  XtFwd p;
  CUtensorMap tmQ, tmK, tmV;
  VT_REQUIRE(xt_map(&tmQ, q->q, q->q_bs, q->q_hs, q->q_rs, q->B, q->H, q->Nq, q->hd, &p.q) == 0, "xattn_tc_fwd: unsupported q layout");
  VT_REQUIRE(xt_map(&tmK, q->k, q->k_bs, q->k_hs, q->k_rs, q->B, q->H, q->Nk, q->hd, &p.k) == 0, "xattn_tc_fwd: unsupported k layout");
  VT_REQUIRE(xt_map(&tmV, q->v, q->v_bs, q->v_hs, q->v_rs, q->B, q->H, q->Nk, q->hd, &p.v) == 0, "xattn_tc_fwd: unsupported v layout");
  VT_REQUIRE((q->o_rs * 2) % 16 == 0 && (q->o_hs * 2) % 16 == 0 && (q->o_bs * 2) % 16 == 0 && ((uintptr_t)q->o & 15) == 0,
             "xattn_tc_fwd: output rows must be 16-byte aligned");
  p.o = static_cast<__nv_bfloat16*>(q->o);
  p.lse = q->lse;
  p.o_bs = q->o_bs; p.o_hs = q->o_hs; p.o_rs = q->o_rs;
  p.H = q->H; p.Nq = q->Nq; p.Nk = q->Nk; p.nkt = (q->Nk + 127) / 128; p.scale = q->scale;
  VT_REQUIRE(q->hd == 96 || q->hd == 64, "xattn_tc_fwd: head dim %d unsupported (64 or 96)", q->hd);
  static bool attr96 = false, attr64 = false;
  int rc = q->hd == 96 ? xt_set_smem(xattn_tc_fwd_kernel<96>, XT_SMEM_FWD, &attr96, "xattn_tc_fwd")
                       : xt_set_smem(xattn_tc_fwd_kernel<64>, XT_SMEM_FWD, &attr64, "xattn_tc_fwd");
  if (rc) return rc;
  dim3 grid((q->Nq + 127) / 128, q->B * q->H);
  if (q->hd == 96) xattn_tc_fwd_kernel<96><<<grid, XT_THREADS, XT_SMEM_FWD, st>>>(tmQ, tmK, tmV, p);
  else xattn_tc_fwd_kernel<64><<<grid, XT_THREADS, XT_SMEM_FWD, st>>>(tmQ, tmK, tmV, p);
  return check_launch("xattn_tc_fwd_kernel");
}

int xattn_tc_bwd_launch(const vt_xattn_bwd_params* q, cudaStream_t st) {
  XtDq a;
  XtDkv c;
  CUtensorMap tmQ, tmK, tmV, tmD;
  VT_REQUIRE(xt_map(&tmQ, q->q, q->q_bs, q->q_hs, q->q_rs, q->B, q->H, q->Nq, q->hd, &a.q) == 0, "xattn_tc_bwd: unsupported q layout");
  VT_REQUIRE(xt_map(&tmK, q->k, q->k_bs, q->k_hs, q->k_rs, q->B, q->H, q->Nk, q->hd, &a.k) == 0, "xattn_tc_bwd: unsupported k layout");
  VT_REQUIRE(xt_map(&tmV, q->v, q->v_bs, q->v_hs, q->v_rs, q->B, q->H, q->Nk, q->hd, &a.v) == 0, "xattn_tc_bwd: unsupported v layout");
  VT_REQUIRE(xt_map(&tmD, q->dout, q->o_bs, q->o_hs, q->o_rs, q->B, q->H, q->Nq, q->hd, &a.d) == 0, "xattn_tc_bwd: unsupported dout layout");
  VT_REQUIRE((q->o_rs * 2) % 16 == 0 && (q->o_hs * 2) % 16 == 0 && (q->o_bs * 2) % 16 == 0 && ((uintptr_t)q->o & 15) == 0 &&
                 ((uintptr_t)q->dout & 15) == 0, "xattn_tc_bwd: o / dout rows must be 16-byte aligned");
  VT_REQUIRE((q->dq_rs * 2) % 16 == 0 && (q->dq_hs * 2) % 16 == 0 && (q->dq_bs * 2) % 16 == 0 && ((uintptr_t)q->dq & 15) == 0,
             "xattn_tc_bwd: dq rows must be 16-byte aligned");
  const size_t kv_bytes = (size_t)q->B * q->H * q->Nk * q->hd * sizeof(float);
  cudaError_t e = cudaMemsetAsync(q->dk, 0, kv_bytes, st);
  VT_REQUIRE(e == cudaSuccess, "xattn_tc_bwd: memset dk: %s", cudaGetErrorString(e));
  e = cudaMemsetAsync(q->dv, 0, kv_bytes, st);
  VT_REQUIRE(e == cudaSuccess, "xattn_tc_bwd: memset dv: %s", cudaGetErrorString(e));
  a.o = static_cast<const __nv_bfloat16*>(q->o);
  a.dout = static_cast<const __nv_bfloat16*>(q->dout);
  a.lse = q->lse;
  a.delta = q->delta;
  a.dq = static_cast<__nv_bfloat16*>(q->dq);
  a.o_bs = q->o_bs; a.o_hs = q->o_hs; a.o_rs = q->o_rs;
  a.dq_bs = q->dq_bs; a.dq_hs = q->dq_hs; a.dq_rs = q->dq_rs;
  a.H = q->H; a.Nq = q->Nq; a.Nk = q->Nk; a.nkt = (q->Nk + 127) / 128; a.scale = q->scale;
  VT_REQUIRE(q->hd == 96 || q->hd == 64, "xattn_tc_bwd: head dim %d unsupported (64 or 96)", q->hd);
  static bool attr_dq96 = false, attr_dkv96 = false, attr_dq64 = false, attr_dkv64 = false;
  int rc = q->hd == 96 ? xt_set_smem(xattn_tc_dq_kernel<96>, XT_SMEM_DQ, &attr_dq96, "xattn_tc_dq")
                       : xt_set_smem(xattn_tc_dq_kernel<64>, XT_SMEM_DQ, &attr_dq64, "xattn_tc_dq");
  if (rc) return rc;
  rc = q->hd == 96 ? xt_set_smem(xattn_tc_dkv_kernel<96>, XT_SMEM_DKV, &attr_dkv96, "xattn_tc_dkv")
                   : xt_set_smem(xattn_tc_dkv_kernel<64>, XT_SMEM_DKV, &attr_dkv64, "xattn_tc_dkv");
  if (rc) return rc;
  dim3 gq((q->Nq + 127) / 128, q->B * q->H);
  if (q->hd == 96) xattn_tc_dq_kernel<96><<<gq, XT_THREADS, XT_SMEM_DQ, st>>>(tmQ, tmK, tmV, tmD, a);
  else xattn_tc_dq_kernel<64><<<gq, XT_THREADS, XT_SMEM_DQ, st>>>(tmQ, tmK, tmV, tmD, a);
  rc = check_launch("xattn_tc_dq_kernel");
  if (rc) return rc;
  c.lse = q->lse;
  c.delta = q->delta;
  c.dk = q->dk;
  c.dv = q->dv;
  c.q = a.q; c.k = a.k; c.v = a.v; c.d = a.d;
  c.H = q->H; c.Nq = q->Nq; c.Nk = q->Nk; c.nqt = (q->Nq + 127) / 128; c.scale = q->scale;
  const int chunks = (c.nqt + XT_QTILES_PER_CHUNK - 1) / XT_QTILES_PER_CHUNK;
  VT_REQUIRE(chunks <= 65535 && q->B * q->H <= 65535, "xattn_tc_bwd: grid too large");
  dim3 gk(a.nkt, chunks, q->B * q->H);
  if (q->hd == 96) xattn_tc_dkv_kernel<96><<<gk, XT_THREADS, XT_SMEM_DKV, st>>>(tmQ, tmK, tmV, tmD, c);
  else xattn_tc_dkv_kernel<64><<<gk, XT_THREADS, XT_SMEM_DKV, st>>>(tmQ, tmK, tmV, tmD, c);
  return check_launch("xattn_tc_dkv_kernel");
}

}  // namespace vt
