// Flash-style softmax attention on tcgen05 tensor cores for the spatial pass (sequence 1+P = 197, hd = 64).
// One CTA per (frame, head); the whole K/V of that problem is resident in shared memory, S / dP / the
// gradient accumulators live in TMEM, softmax is single-pass (no online rescale needed).
//
//   forward : S_t = Q_t K^T (tcgen05, N = keys padded to 16)  -> 2 softmax warpgroups (thread = query row):
//             row max, exp2, row sum, P (bf16) into 128B-swizzled smem -> O_t = P_t V (V read MN-major)
//             -> O / l, lse.
//   backward: for each key tile kt (128 keys) x query tile qt (128 queries):
//             S = Q K^T, dP = dO V^T (TMEM) -> P = exp(S*scale - lse), dS = P (dP - delta) scale  (bf16, smem)
//             dK_kt += dS^T Q_qt, dV_kt += P^T dO_qt (P/dS read MN-major, i.e. transposed for free),
//             dQ_qt += dS K_kt.   TMEM: S 128 | dP 128 | dK 64 | dV 64 | dQ0 64 | dQ1 64 = 512 columns.
// Operands arrive by TMA straight out of the packed qkv / dctx activations (no head split copies).
#include "vt_common.cuh"
#include "vt_umma.cuh"

namespace vt {

int make_tmap_bf16_2d(CUtensorMap* map, const void* base, long long rows, long long cols, long long ld, int box_rows);

// diagnostics: per-CTA clock64 stamps [cta][32], set through vt_debug_buffer(); NULL in production
__device__ long long* g_attn_dbg = nullptr;
#define VT_STAMP(i) do { if (dbg) dbg[(i)] = clock64(); } while (0)

constexpr int TC_HD = 64;
constexpr int TILE_BYTES = 128 * 128;  // 128 rows x 64 bf16 (one swizzled K-major block) = 16 KiB
constexpr float LOG2E = 1.4426950408889634f;

// 16-byte store of 8 bf16 into a K-major SW128 block: row r, 8-column group g (0..7)
__device__ __forceinline__ void st_sw128(uint8_t* block, int r, int g, uint4 v) {
  *reinterpret_cast<uint4*>(block + r * 128 + ((g ^ (r & 7)) << 4)) = v;
}

// ================================================================================================
// forward
// ================================================================================================
struct AttnTcFwd {
  __nv_bfloat16* ctx;
  float* lse;
  int N, H, NK, tiles;
  float scale;
};

constexpr int FWD_THREADS = 288;

__global__ void __launch_bounds__(FWD_THREADS, 1)
attn_tc_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV, const AttnTcFwd p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int kv_bytes = ((p.NK * 128 + 1023) / 1024) * 1024;
  uint8_t* sQ = smem;                       // 2 tiles
  uint8_t* sK = sQ + 2 * TILE_BYTES;
  uint8_t* sV = sK + kv_bytes;
  uint8_t* sP = sV + kv_bytes;              // 2 x (4 blocks)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * 4 * TILE_BYTES);
  uint64_t* bar_q = bars + 0;
  uint64_t* bar_k = bars + 1;
  uint64_t* bar_v = bars + 2;
  uint64_t* bar_s = bars + 3;   // [2]
  uint64_t* bar_p = bars + 5;   // [2]
  uint64_t* bar_o = bars + 7;   // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bh = blockIdx.x, bp = bh / p.H, h = bh - bp * p.H;
  const int row0 = bp * p.N;
  long long* dbg = (g_attn_dbg && (lane == 0) && (warp == 8 || warp == 0)) ? g_attn_dbg + (long long)blockIdx.x * 32 + (warp == 8 ? 0 : 16) : nullptr;
  VT_STAMP(0);

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmKV);
    mbar_init(bar_q, 1); mbar_init(bar_k, 1); mbar_init(bar_v, 1);
    for (int t = 0; t < 2; ++t) { mbar_init(&bar_s[t], 1); mbar_init(&bar_p[t], 128); mbar_init(&bar_o[t], 1); }
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 8) {
    if (lane == 0) {
      mbar_arrive_expect_tx(bar_q, p.tiles * TILE_BYTES);
      for (int t = 0; t < p.tiles; ++t) tma_load_2d(sQ + t * TILE_BYTES, &tmQ, bar_q, h * TC_HD, row0 + t * 128);
      mbar_arrive_expect_tx(bar_k, p.NK * 128);
      tma_load_2d(sK, &tmKV, bar_k, (p.H + h) * TC_HD, row0);
      mbar_arrive_expect_tx(bar_v, p.NK * 128);
      tma_load_2d(sV, &tmKV, bar_v, (2 * p.H + h) * TC_HD, row0);

      VT_STAMP(1);
      mbar_wait(bar_q, 0);
      mbar_wait(bar_k, 0);
      tc_fence_after();
      VT_STAMP(2);
      const uint32_t idesc_s = make_idesc_bf16(128, (uint32_t)p.NK, 0, 0);
      for (int t = 0; t < p.tiles; ++t) {
        const uint32_t qa = smem_u32(sQ + t * TILE_BYTES), ka = smem_u32(sK);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16_ss(tmem_base + t * 256, sdesc_kmajor(qa + k * 32), sdesc_kmajor(ka + k * 32), idesc_s, k > 0);
        umma_commit(&bar_s[t]);
      }
      mbar_wait(bar_v, 0);
      VT_STAMP(3);
      const uint32_t idesc_o = make_idesc_bf16(128, 64, 0, 1);
      const int ksteps = p.NK / 16;
      for (int t = 0; t < p.tiles; ++t) {
        mbar_wait(&bar_p[t], 0);
        tc_fence_after();
        VT_STAMP(4 + t);
        const uint32_t pa = smem_u32(sP + t * 4 * TILE_BYTES), va = smem_u32(sV);
        for (int s = 0; s < ksteps; ++s)
          umma_bf16_ss(tmem_base + t * 256, sdesc_kmajor(pa + (s >> 2) * TILE_BYTES + (s & 3) * 32),
                       sdesc_mnmajor(va + s * 2048, 8192), idesc_o, s > 0);
        umma_commit(&bar_o[t]);
      }
    }
  } else {
    const int t = warp >> 2;                 // warpgroup = query tile
    if (t < p.tiles) {
      const int r = (warp & 3) * 32 + lane;  // row in tile == TMEM lane
      const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(t * 256);
      const float sl2 = p.scale * LOG2E;
      const int nchunks = (p.NK + 31) / 32;
      mbar_wait(&bar_s[t], 0);
      tc_fence_after();
      VT_STAMP(1);
      float mx = -INFINITY;
      for (int c = 0; c < nchunks; ++c) {
        uint32_t v[32];
        tmem_ld32(taddr + c * 32, v);
        tmem_ld_wait();
        if (c * 32 + 32 <= p.N) {
#pragma unroll
          for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(v[j]));
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (c * 32 + j < p.N) mx = fmaxf(mx, __uint_as_float(v[j]));
        }
      }
      VT_STAMP(2);
      const float mb = mx * sl2;
      float l = 0.f;
      uint8_t* Pt = sP + t * 4 * TILE_BYTES;
      for (int c = 0; c < nchunks; ++c) {
        uint32_t v[32];
        tmem_ld32(taddr + c * 32, v);
        tmem_ld_wait();
        float e[32];
        if (c * 32 + 32 <= p.N) {      // full chunk: no key masking
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            e[j] = fast_exp2(fmaf(__uint_as_float(v[j]), sl2, -mb));
            l += e[j];
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            e[j] = (c * 32 + j < p.N) ? fast_exp2(fmaf(__uint_as_float(v[j]), sl2, -mb)) : 0.f;
            l += e[j];
          }
        }
        uint8_t* blk = Pt + (c >> 1) * TILE_BYTES;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 o;
          o.x = pack_bf16x2(e[g * 8 + 0], e[g * 8 + 1]);
          o.y = pack_bf16x2(e[g * 8 + 2], e[g * 8 + 3]);
          o.z = pack_bf16x2(e[g * 8 + 4], e[g * 8 + 5]);
          o.w = pack_bf16x2(e[g * 8 + 6], e[g * 8 + 7]);
          st_sw128(blk, r, (c & 1) * 4 + g, o);
        }
      }
      tc_fence_before();
      fence_proxy_async();
      mbar_arrive(&bar_p[t]);
      VT_STAMP(3);
      mbar_wait(&bar_o[t], 0);
      tc_fence_after();
      VT_STAMP(4);
      const int q = t * 128 + r;
      const float inv = 1.0f / l;
      uint32_t o0[32], o1[32];
      tmem_ld32(taddr, o0);
      tmem_ld32(taddr + 32, o1);
      tmem_ld_wait();
      if (q < p.N) {
        uint4* dst = reinterpret_cast<uint4*>(p.ctx + ((long long)(row0 + q) * p.H + h) * TC_HD);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 o;
          o.x = pack_bf16x2(__uint_as_float(o0[g * 8 + 0]) * inv, __uint_as_float(o0[g * 8 + 1]) * inv);
          o.y = pack_bf16x2(__uint_as_float(o0[g * 8 + 2]) * inv, __uint_as_float(o0[g * 8 + 3]) * inv);
          o.z = pack_bf16x2(__uint_as_float(o0[g * 8 + 4]) * inv, __uint_as_float(o0[g * 8 + 5]) * inv);
          o.w = pack_bf16x2(__uint_as_float(o0[g * 8 + 6]) * inv, __uint_as_float(o0[g * 8 + 7]) * inv);
          dst[g] = o;
        }
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 o;
          o.x = pack_bf16x2(__uint_as_float(o1[g * 8 + 0]) * inv, __uint_as_float(o1[g * 8 + 1]) * inv);
          o.y = pack_bf16x2(__uint_as_float(o1[g * 8 + 2]) * inv, __uint_as_float(o1[g * 8 + 3]) * inv);
          o.z = pack_bf16x2(__uint_as_float(o1[g * 8 + 4]) * inv, __uint_as_float(o1[g * 8 + 5]) * inv);
          o.w = pack_bf16x2(__uint_as_float(o1[g * 8 + 6]) * inv, __uint_as_float(o1[g * 8 + 7]) * inv);
          dst[4 + g] = o;
        }
        p.lse[(long long)bh * p.N + q] = mx * p.scale + __logf(l);
      }
      VT_STAMP(5);
    }
  }
  tc_fence_before();
  __syncthreads();
  VT_STAMP(15);
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// ================================================================================================
// backward
// ================================================================================================
struct AttnTcBwd {
  const __nv_bfloat16* ctx;
  const __nv_bfloat16* dctx;
  const float* lse;
  __nv_bfloat16* dqkv;
  int N, H, nq, nk, NK0, NK1;
  float scale;
};

constexpr int BWD_THREADS = 288;
constexpr uint32_t COL_S = 0, COL_DP = 128, COL_DK = 256, COL_DV = 320, COL_DQ = 384;

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

__device__ __forceinline__ void store_row64_bf16(__nv_bfloat16* dst, const uint32_t (&a)[32], const uint32_t (&b)[32]) {
  uint4* d = reinterpret_cast<uint4*>(dst);
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    uint4 o;
    o.x = pack_bf16x2(__uint_as_float(a[g * 8 + 0]), __uint_as_float(a[g * 8 + 1]));
    o.y = pack_bf16x2(__uint_as_float(a[g * 8 + 2]), __uint_as_float(a[g * 8 + 3]));
    o.z = pack_bf16x2(__uint_as_float(a[g * 8 + 4]), __uint_as_float(a[g * 8 + 5]));
    o.w = pack_bf16x2(__uint_as_float(a[g * 8 + 6]), __uint_as_float(a[g * 8 + 7]));
    d[g] = o;
  }
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    uint4 o;
    o.x = pack_bf16x2(__uint_as_float(b[g * 8 + 0]), __uint_as_float(b[g * 8 + 1]));
    o.y = pack_bf16x2(__uint_as_float(b[g * 8 + 2]), __uint_as_float(b[g * 8 + 3]));
    o.z = pack_bf16x2(__uint_as_float(b[g * 8 + 4]), __uint_as_float(b[g * 8 + 5]));
    o.w = pack_bf16x2(__uint_as_float(b[g * 8 + 6]), __uint_as_float(b[g * 8 + 7]));
    d[4 + g] = o;
  }
}

__global__ void __launch_bounds__(BWD_THREADS, 1)
attn_tc_bwd_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO, const AttnTcBwd p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                        // 2 query tiles
  uint8_t* sDO = sQ + 2 * TILE_BYTES;        // 2 query tiles
  uint8_t* sK = sDO + 2 * TILE_BYTES;        // 2 key tiles
  uint8_t* sV = sK + 2 * TILE_BYTES;         // 2 key tiles
  uint8_t* sP = sV + 2 * TILE_BYTES;         // 2 blocks of 64 keys
  uint8_t* sDS = sP + 2 * TILE_BYTES;        // 2 blocks
  float* lse_s = reinterpret_cast<float*>(sDS + 2 * TILE_BYTES);  // [256]
  float* del_s = lse_s + 256;                                     // [256]
  uint64_t* bars = reinterpret_cast<uint64_t*>(del_s + 256);
  uint64_t* bar_load = bars + 0;
  uint64_t* bar_sdp = bars + 1;       // S,dP in TMEM
  uint64_t* bar_pds = bars + 2;       // P,dS in smem (256 arrivals)
  uint64_t* bar_dkv_full = bars + 3;
  uint64_t* bar_dkv_free = bars + 4;  // 256 arrivals
  uint64_t* bar_dq = bars + 5;
  uint64_t* bar_load2 = bars + 6;     // second query tile / second key tile
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 7);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bh = blockIdx.x, bp = bh / p.H, h = bh - bp * p.H;
  const int row0 = bp * p.N;
  const int n_it = p.nq * p.nk;
  long long* dbg = (g_attn_dbg && (lane == 0) && (warp == 8 || warp == 0)) ? g_attn_dbg + (long long)blockIdx.x * 32 + (warp == 8 ? 0 : 16) : nullptr;
  VT_STAMP(0);

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tmQKV);
    tma_prefetch_desc(&tmDO);
    mbar_init(bar_load, 1); mbar_init(bar_sdp, 1); mbar_init(bar_pds, 256);
    mbar_init(bar_dkv_full, 1); mbar_init(bar_dkv_free, 256); mbar_init(bar_dq, 1); mbar_init(bar_load2, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 8) {
    if (lane == 0) {
      // first (query, key) tile pair on its own barrier so iteration 0 starts while the rest is in flight
      mbar_arrive_expect_tx(bar_load, 4 * TILE_BYTES);
      tma_load_2d(sQ, &tmQKV, bar_load, h * TC_HD, row0);
      tma_load_2d(sK, &tmQKV, bar_load, (p.H + h) * TC_HD, row0);
      tma_load_2d(sDO, &tmDO, bar_load, h * TC_HD, row0);
      tma_load_2d(sV, &tmQKV, bar_load, (2 * p.H + h) * TC_HD, row0);
      if (p.nq > 1) {
        mbar_arrive_expect_tx(bar_load2, 4 * TILE_BYTES);
        tma_load_2d(sQ + TILE_BYTES, &tmQKV, bar_load2, h * TC_HD, row0 + 128);
        tma_load_2d(sDO + TILE_BYTES, &tmDO, bar_load2, h * TC_HD, row0 + 128);
        tma_load_2d(sK + TILE_BYTES, &tmQKV, bar_load2, (p.H + h) * TC_HD, row0 + 128);
        tma_load_2d(sV + TILE_BYTES, &tmQKV, bar_load2, (2 * p.H + h) * TC_HD, row0 + 128);
      }
      VT_STAMP(1);
      mbar_wait(bar_load, 0);
      tc_fence_after();
      VT_STAMP(2);
      const uint32_t idesc_t = make_idesc_bf16(128, 64, 1, 1);   // dK, dV: A, B MN-major
      const uint32_t idesc_q = make_idesc_bf16(128, 64, 0, 1);   // dQ: A K-major, B MN-major
      const uint32_t pa = smem_u32(sP), dsa = smem_u32(sDS);
      for (int it = 0; it < n_it; ++it) {
        const int kt = it / p.nq, qt = it - kt * p.nq;
        const int NKt = kt == 0 ? p.NK0 : p.NK1;
        if (it == 1) { mbar_wait(bar_load2, 0); tc_fence_after(); }
        const uint32_t idesc_s = make_idesc_bf16(128, (uint32_t)NKt, 0, 0);
        const uint32_t qa = smem_u32(sQ + qt * TILE_BYTES), doa = smem_u32(sDO + qt * TILE_BYTES);
        const uint32_t ka = smem_u32(sK + kt * TILE_BYTES), va = smem_u32(sV + kt * TILE_BYTES);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16_ss(tmem_base + COL_S, sdesc_kmajor(qa + k * 32), sdesc_kmajor(ka + k * 32), idesc_s, k > 0);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16_ss(tmem_base + COL_DP, sdesc_kmajor(doa + k * 32), sdesc_kmajor(va + k * 32), idesc_s, k > 0);
        umma_commit(bar_sdp);
        VT_STAMP(3 + 2 * it);
        mbar_wait(bar_pds, it & 1);
        tc_fence_after();
        VT_STAMP(4 + 2 * it);
        if (qt == 0 && kt > 0) {
          mbar_wait(bar_dkv_free, (kt - 1) & 1);
          tc_fence_after();
        }
        // dK_kt (+)= dS^T Q_qt ; dV_kt (+)= P^T dO_qt     (K = 128 queries)
#pragma unroll
        for (int s = 0; s < 8; ++s)
          umma_bf16_ss(tmem_base + COL_DK, sdesc_mnmajor(dsa + s * 2048, TILE_BYTES), sdesc_mnmajor(qa + s * 2048, 8192),
                       idesc_t, (qt > 0 || s > 0) ? 1u : 0u);
#pragma unroll
        for (int s = 0; s < 8; ++s)
          umma_bf16_ss(tmem_base + COL_DV, sdesc_mnmajor(pa + s * 2048, TILE_BYTES), sdesc_mnmajor(doa + s * 2048, 8192),
                       idesc_t, (qt > 0 || s > 0) ? 1u : 0u);
        // dQ_qt (+)= dS K_kt                                (K = NKt keys)
        for (int s = 0; s < NKt / 16; ++s)
          umma_bf16_ss(tmem_base + COL_DQ + qt * 64, sdesc_kmajor(dsa + (s >> 2) * TILE_BYTES + (s & 3) * 32),
                       sdesc_mnmajor(ka + s * 2048, 8192), idesc_q, (kt > 0 || s > 0) ? 1u : 0u);
        if (qt == p.nq - 1) umma_commit(bar_dkv_full);
      }
      umma_commit(bar_dq);
      VT_STAMP(11);
    }
  } else {
    const int quad = warp & 3, half = warp >> 2;
    const int r = quad * 32 + lane;
    const uint32_t tlane = tmem_base + ((uint32_t)(quad * 32) << 16);
    {  // delta_i = dO_i . O_i and lse for query row half*128 + r
      const int q = half * 128 + r;
      float d = 0.f, ls = 0.f;
      if (q < p.N) {
        const uint4* o4 = reinterpret_cast<const uint4*>(p.ctx + ((long long)(row0 + q) * p.H + h) * TC_HD);
        const uint4* g4 = reinterpret_cast<const uint4*>(p.dctx + ((long long)(row0 + q) * p.H + h) * TC_HD);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const uint4 a = o4[i], b = g4[i];
          const float2 a0 = unpack_bf16x2(a.x), a1 = unpack_bf16x2(a.y), a2 = unpack_bf16x2(a.z), a3 = unpack_bf16x2(a.w);
          const float2 b0 = unpack_bf16x2(b.x), b1 = unpack_bf16x2(b.y), b2 = unpack_bf16x2(b.z), b3 = unpack_bf16x2(b.w);
          d += a0.x * b0.x + a0.y * b0.y + a1.x * b1.x + a1.y * b1.y + a2.x * b2.x + a2.y * b2.y + a3.x * b3.x + a3.y * b3.y;
        }
        ls = p.lse[(long long)bh * p.N + q];
      }
      del_s[q] = d;
      lse_s[q] = ls * LOG2E;
    }
    named_bar_sync(1, 256);
    VT_STAMP(1);
    const float sl2 = p.scale * LOG2E;
    for (int it = 0; it < n_it; ++it) {
      const int kt = it / p.nq, qt = it - kt * p.nq;
      const int NKt = kt == 0 ? p.NK0 : p.NK1;
      mbar_wait(bar_sdp, it & 1);
      tc_fence_after();
      VT_STAMP(2 + 2 * it);
      const int q = qt * 128 + r;
      const float lq = lse_s[q], dq = del_s[q];
      const bool qok = q < p.N;
#pragma unroll 1
      for (int cc = 0; cc < 2; ++cc) {
        const int c0 = half * 64 + cc * 32;       // first key column (within the tile) of this chunk
        if (c0 >= NKt) break;                      // warp-uniform
        uint32_t sv[32], dv[32];
        tmem_ld32(tlane + COL_S + c0, sv);
        tmem_ld32(tlane + COL_DP + c0, dv);
        tmem_ld_wait();
        float pv[32], ds[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const bool ok = qok && (kt * 128 + c0 + j < p.N);
          const float pj = ok ? fast_exp2(fmaf(__uint_as_float(sv[j]), sl2, -lq)) : 0.f;
          pv[j] = pj;
          ds[j] = pj * (__uint_as_float(dv[j]) - dq) * p.scale;
        }
        uint8_t* pb = sP + half * TILE_BYTES;
        uint8_t* db = sDS + half * TILE_BYTES;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 o;
          o.x = pack_bf16x2(pv[g * 8 + 0], pv[g * 8 + 1]); o.y = pack_bf16x2(pv[g * 8 + 2], pv[g * 8 + 3]);
          o.z = pack_bf16x2(pv[g * 8 + 4], pv[g * 8 + 5]); o.w = pack_bf16x2(pv[g * 8 + 6], pv[g * 8 + 7]);
          st_sw128(pb, r, cc * 4 + g, o);
          o.x = pack_bf16x2(ds[g * 8 + 0], ds[g * 8 + 1]); o.y = pack_bf16x2(ds[g * 8 + 2], ds[g * 8 + 3]);
          o.z = pack_bf16x2(ds[g * 8 + 4], ds[g * 8 + 5]); o.w = pack_bf16x2(ds[g * 8 + 6], ds[g * 8 + 7]);
          st_sw128(db, r, cc * 4 + g, o);
        }
      }
      tc_fence_before();
      fence_proxy_async();
      mbar_arrive(bar_pds);
      VT_STAMP(3 + 2 * it);
      if (qt == p.nq - 1) {
        mbar_wait(bar_dkv_full, kt & 1);
        tc_fence_after();
        uint32_t a[32], b[32];
        const uint32_t col = half == 0 ? COL_DK : COL_DV;
        tmem_ld32(tlane + col, a);
        tmem_ld32(tlane + col + 32, b);
        tmem_ld_wait();
        const int key = kt * 128 + r;
        if (key < p.N)
          store_row64_bf16(p.dqkv + ((long long)(row0 + key) * 3 + 1 + half) * p.H * TC_HD + h * TC_HD, a, b);
        tc_fence_before();
        mbar_arrive(bar_dkv_free);
      }
    }
    mbar_wait(bar_dq, 0);
    tc_fence_after();
    VT_STAMP(12);
    if (half < p.nq) {
      uint32_t a[32], b[32];
      tmem_ld32(tlane + COL_DQ + half * 64, a);
      tmem_ld32(tlane + COL_DQ + half * 64 + 32, b);
      tmem_ld_wait();
      const int q = half * 128 + r;
      if (q < p.N) store_row64_bf16(p.dqkv + ((long long)(row0 + q) * 3) * p.H * TC_HD + h * TC_HD, a, b);
    }
    VT_STAMP(13);
  }
  tc_fence_before();
  __syncthreads();
  VT_STAMP(15);
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------------
// host launchers (called from vt_attn_fwd / vt_attn_bwd dispatch in vt_attention.cu)
// ------------------------------------------------------------------------------------------------
static int round16(int v) { return (v + 15) & ~15; }

int attn_tc_fwd_launch(const vt_attn_fwd_params* q, cudaStream_t st) {
  const int N = q->N, H = q->H;
  AttnTcFwd p;
  p.ctx = static_cast<__nv_bfloat16*>(q->ctx);
  p.lse = q->lse;
  p.N = N; p.H = H; p.NK = round16(N); p.tiles = (N + 127) / 128; p.scale = q->scale;
  const long long rows = (long long)q->Bp * N, ld = 3LL * H * TC_HD;
  CUtensorMap tmQ, tmKV;
  int rc = make_tmap_bf16_2d(&tmQ, q->qkv, rows, ld, ld, 128);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tmKV, q->qkv, rows, ld, ld, p.NK);
  if (rc) return rc;
  const int kv_bytes = ((p.NK * 128 + 1023) / 1024) * 1024;
  const int smem = 2 * TILE_BYTES + 2 * kv_bytes + 8 * TILE_BYTES + 256 + 1024;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(attn_tc_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
    VT_REQUIRE(e == cudaSuccess, "attn_tc_fwd: smem attribute: %s", cudaGetErrorString(e));
    attr = true;
  }
  VT_REQUIRE(smem <= 232448, "attn_tc_fwd: smem %d too large", smem);
  attn_tc_fwd_kernel<<<q->Bp * H, FWD_THREADS, smem, st>>>(tmQ, tmKV, p);
  return check_launch("attn_tc_fwd_kernel");
}

int attn_tc_bwd_launch(const vt_attn_bwd_params* q, cudaStream_t st) {
  const int N = q->N, H = q->H;
  AttnTcBwd p;
  p.ctx = static_cast<const __nv_bfloat16*>(q->ctx);
  p.dctx = static_cast<const __nv_bfloat16*>(q->dctx);
  p.lse = q->lse;
  p.dqkv = static_cast<__nv_bfloat16*>(q->dqkv);
  p.N = N; p.H = H; p.scale = q->scale;
  p.nq = (N + 127) / 128;
  p.nk = p.nq;
  p.NK0 = N >= 128 ? 128 : round16(N);
  p.NK1 = N > 128 ? round16(N - 128) : 16;
  const long long rows = (long long)q->Bp * N;
  CUtensorMap tmQKV, tmDO;
  int rc = make_tmap_bf16_2d(&tmQKV, q->qkv, rows, 3LL * H * TC_HD, 3LL * H * TC_HD, 128);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tmDO, q->dctx, rows, (long long)H * TC_HD, (long long)H * TC_HD, 128);
  if (rc) return rc;
  const int smem = 12 * TILE_BYTES + 2 * 256 * 4 + 256 + 1024;   // 6 operand/P/dS tile pairs + lse/delta + barriers
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(attn_tc_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
    VT_REQUIRE(e == cudaSuccess, "attn_tc_bwd: smem attribute: %s", cudaGetErrorString(e));
    attr = true;
  }
  attn_tc_bwd_kernel<<<q->Bp * H, BWD_THREADS, smem, st>>>(tmQKV, tmDO, p);
  return check_launch("attn_tc_bwd_kernel");
}

}  // namespace vt

extern "C" int vt_debug_buffer(void* ptr) {
  long long* p = static_cast<long long*>(ptr);
  cudaError_t e = cudaMemcpyToSymbol(vt::g_attn_dbg, &p, sizeof(p));
  if (e != cudaSuccess) { vt::set_error("vt_debug_buffer: %s", cudaGetErrorString(e)); return 1; }
  return 0;
}
