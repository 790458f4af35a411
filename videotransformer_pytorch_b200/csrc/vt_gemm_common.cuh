// Shared pieces of the tcgen05 GEMM kernels (vt_gemm.cu: one CTA per tile; vt_gemm2.cu: CTA pairs, cta_group::2):
// launch-time descriptor of the problem and the fused epilogue that drains one 128 x BN accumulator tile.
#pragma once
#include "vt_common.cuh"
#include "vt_umma.cuh"

namespace vt {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int GEMM_THREADS = 384;        // 4 control warps + 8 epilogue warps
constexpr int CHUNK_BYTES = 64 * BK * 2;  // one 64-wide MN chunk of an MN-major tile (8 KiB)
constexpr int EPI_PITCH = 33;             // words per staged row (32 + 1 pad: conflict-free row writes and column-group reads)
constexpr int EPI_WARPS = 8;

struct GemmDev {
  int M, N, K;
  int num_m, num_n, splits, kblocks;
  int num_mp;   // macro row-tiles: ceil(num_m / cluster size)
  int a_mn, b_mn, epi;
  const float* bias;
  const float* bias2;      // VT_EPI_F32 with aux only: second bias, added after the row scale (unscaled)
  void* out;
  void* out2;
  const void* aux;
  long long ldo, ldo2, ldaux;
  const int* out_row;
  const int* aux_row;
  const float* row_scale;
  long long split_stride;  // elements between split partials (EPI_F32 only)
  long long* dbg;          // optional diagnostics: per-CTA clock64 stamps [cta][16] (NULL in production)
  int tma_store;           // 1: plain row-major output written by TMA bulk stores (epilogue_tile_tma); 2: split-K partials
                           //    reduce-added into the (pre-zeroed) output by TMA; 3: fp32 residual epilogue, residual
                           //    boxes TMA-loaded into the staging slot, result TMA-stored (epilogue_tile_tma_res)
  // affine row map of the residual epilogue: GEMM row m -> outer = m / map_period, inner = m % map_period;
  // inner < map_skip: "special" row (replicated cls token), else tensor coordinates
  // t = outer % map_tcount, p = inner - map_skip, b = outer / map_tcount -> box start (col, p * map_tcount + t, b) in the
  // (col, row in sample, sample) out / aux maps
  int map_period, map_skip, map_tcount;
  int map_rank;            // 3: tensor maps are (col, row in sample, sample), box {32, 32 * map_tcount, 1} with element stride map_tcount
  float* special_out;      // special rows go to special_out + outer * special_ld (plain per-thread stores), or are dropped
  long long special_ld;
  // narrow tail units: the last (partial) macro row of tiles is cut into units of tail_bn columns so that its few valid
  // rows do not cost a whole extra round of full tiles
  int tail_bn;             // 0 = none
  int full_units;          // units [0, full_units) are regular tiles x splits; [full_units, full_units + tail_units) are tail units
  int tail_units;
  int tail_mp;             // macro row index of the tail units
};

// One unit of work of the persistent loop.
struct GemmUnit {
  int mp;        // macro row (x cluster size + cluster rank = m block)
  int n0;        // first output column
  int bn;        // width of the unit (BN, or tail_bn)
  int split, kb0, kb1;
};
template <int BN>
__device__ __forceinline__ GemmUnit decode_unit(const GemmDev& p, int unit) {
  GemmUnit u;
  if (unit < p.full_units) {
    const int tile = unit / p.splits;
    u.split = unit - tile * p.splits;
    u.mp = tile / p.num_n;                     // n fastest: the A row-block stays hot in L2
    u.n0 = (tile - u.mp * p.num_n) * BN;
    u.bn = BN;
    u.kb0 = (int)(((long long)p.kblocks * u.split) / p.splits);
    u.kb1 = (int)(((long long)p.kblocks * (u.split + 1)) / p.splits);
  } else {
    u.split = 0;
    u.mp = p.tail_mp;
    u.n0 = (unit - p.full_units) * p.tail_bn;
    u.bn = p.tail_bn;
    u.kb0 = 0;
    u.kb1 = p.kblocks;
  }
  return u;
}

// Drain accumulator tile (m_blk, n_blk) of this CTA: TMEM columns [t_base, t_base + BN) of lane quadrant q.
// Called by the 8 epilogue warps; waits on `tfull` (parity `ph`) after issuing the first operand prefetch.
template <int BN>
__device__ __forceinline__ void epilogue_tile(const GemmDev& p, float* stg, uint32_t t_base, int m_blk, int n_base, int bn, int split,
                                              int q, int half, int lane, uint64_t* tfull, uint32_t ph) {
  const int rsub = lane >> 3, cg = lane & 7;
  const bool f32_aux = p.epi == VT_EPI_F32 && p.aux != nullptr;
  const bool z_aux = p.epi == VT_EPI_DGELU;
  // per-row metadata of the 8 rows this lane serves in the transposed phase
  float rs[8];
  int orow[8], arow[8];
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int row = m_blk * BM + q * 32 + it * 4 + rsub;
    const bool ok = row < p.M;
    rs[it] = (ok && p.row_scale) ? p.row_scale[row] : 1.0f;
    orow[it] = ok ? (p.out_row ? p.out_row[row] : row) : -1;
    arow[it] = !ok ? -1 : (f32_aux ? (p.aux_row ? p.aux_row[row] : row) : (z_aux ? row : -1));
  }
  uint4 pre[8];
#pragma unroll
  for (int it = 0; it < 8; ++it) pre[it] = make_uint4(0u, 0u, 0u, 0u);   // no epilogue operand => adds 0
  auto prefetch = [&](int c) {
    const int n = n_base + c * 32 + cg * 4;
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      pre[it] = make_uint4(0u, 0u, 0u, 0u);
      if (arow[it] >= 0 && orow[it] >= 0 && n < p.N) {
        if (f32_aux) {
          pre[it] = *reinterpret_cast<const uint4*>(static_cast<const float*>(p.aux) + (long long)arow[it] * p.ldaux + n);
        } else if (z_aux) {
          const uint2 z = *reinterpret_cast<const uint2*>(static_cast<const __nv_bfloat16*>(p.aux) + (long long)arow[it] * p.ldaux + n);
          pre[it].x = z.x; pre[it].y = z.y;
        }
      }
    }
  };
  if (f32_aux || z_aux) prefetch(half);
  mbar_wait(tfull, ph);
  tc_fence_after();
#pragma unroll 1
  for (int c = half; c < bn / 32; c += 2) {
    uint32_t r[32];
    tmem_ld32(t_base + c * 32, r);
    tmem_ld_wait();
    const int n = n_base + c * 32 + cg * 4;
    if (n_base + c * 32 >= p.N) break;   // warp-uniform
#pragma unroll
    for (int j = 0; j < 32; ++j) stg[lane * EPI_PITCH + j] = __uint_as_float(r[j]);
    uint4 cur[8];
#pragma unroll
    for (int it = 0; it < 8; ++it) cur[it] = pre[it];
    if ((f32_aux || z_aux) && c + 2 < bn / 32) prefetch(c + 2);
    __syncwarp();
    float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p.bias && n < p.N) bias4 = __ldg(reinterpret_cast<const float4*>(p.bias + n));
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const float* sp = stg + (it * 4 + rsub) * EPI_PITCH + cg * 4;
      float4 v = make_float4(sp[0], sp[1], sp[2], sp[3]);
      if (orow[it] < 0 || n >= p.N) continue;
      v.x += bias4.x; v.y += bias4.y; v.z += bias4.z; v.w += bias4.w;
      const long long o = (long long)orow[it];
      if (p.epi == VT_EPI_BF16) {
        const float s = rs[it];
        uint2 w;
        w.x = pack_bf16x2(s * v.x, s * v.y);
        w.y = pack_bf16x2(s * v.z, s * v.w);
        *reinterpret_cast<uint2*>(static_cast<__nv_bfloat16*>(p.out) + o * p.ldo + n) = w;
      } else if (p.epi == VT_EPI_F32) {
        const float s = rs[it];
        v.x = fmaf(s, v.x, __uint_as_float(cur[it].x)); v.y = fmaf(s, v.y, __uint_as_float(cur[it].y));
        v.z = fmaf(s, v.z, __uint_as_float(cur[it].z)); v.w = fmaf(s, v.w, __uint_as_float(cur[it].w));
        if (p.bias2) {
          const float4 b2 = __ldg(reinterpret_cast<const float4*>(p.bias2 + n));
          v.x += b2.x; v.y += b2.y; v.z += b2.z; v.w += b2.w;
        }
        *reinterpret_cast<float4*>(static_cast<float*>(p.out) + (long long)split * p.split_stride + o * p.ldo + n) = v;
      } else if (p.epi == VT_EPI_GELU) {
        uint2 z, h;
        z.x = pack_bf16x2(v.x, v.y); z.y = pack_bf16x2(v.z, v.w);
        h.x = pack_bf16x2(gelu_fast(v.x), gelu_fast(v.y)); h.y = pack_bf16x2(gelu_fast(v.z), gelu_fast(v.w));
        *reinterpret_cast<uint2*>(static_cast<__nv_bfloat16*>(p.out) + o * p.ldo + n) = z;
        *reinterpret_cast<uint2*>(static_cast<__nv_bfloat16*>(p.out2) + o * p.ldo2 + n) = h;
      } else {  // VT_EPI_DGELU: cur = z (bf16 x4)
        const float2 z0 = unpack_bf16x2(cur[it].x), z1 = unpack_bf16x2(cur[it].y);
        uint2 w;
        w.x = pack_bf16x2(v.x * dgelu_fast(z0.x), v.y * dgelu_fast(z0.y));
        w.y = pack_bf16x2(v.z * dgelu_fast(z1.x), v.w * dgelu_fast(z1.y));
        *reinterpret_cast<uint2*>(static_cast<__nv_bfloat16*>(p.out) + o * p.ldo + n) = w;
      }
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------------
// TMA-store epilogue for plain row-major outputs (no row maps, no epilogue operand): bf16 or fp32.
// Stays in tcgen05.ld's natural layout (thread = accumulator row): bias / row-scale in registers, the 32x32 block
// is written once to a swizzled shared-memory box (conflict-free 16-byte stores) and leaves with ONE
// cp.async.bulk.tensor store per chunk — no transposed read-back, no per-thread global stores or address math.
// Each warp owns a 4 KiB staging slot (two 2 KiB boxes for bf16 => double buffered; one 4 KiB box for fp32).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, const void* smem_src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
// same box, but added into global memory (fp32 reduction performed by the L2): split-K partial tiles accumulate straight
// into the output, no workspace round trip and no second kernel
__device__ __forceinline__ void tma_reduce_add_3d(const CUtensorMap* map, const void* smem_src, int c0, int c1, int c2) {
  asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}

template <int BN>
__device__ __forceinline__ void epilogue_tile_tma(const GemmDev& p, const CUtensorMap* tmC, const CUtensorMap* tmC2, uint8_t* slot,
                                                  uint32_t t_base, int m_blk, int n_base, int bn, int split, int q, int half, int lane,
                                                  uint64_t* tfull, uint32_t ph) {
  const int row = m_blk * BM + q * 32 + lane;
  const float s = (row < p.M && p.row_scale) ? p.row_scale[row] : 1.0f;
  const bool f32 = p.epi == VT_EPI_F32;
  const bool gelu = p.epi == VT_EPI_GELU;      // two bf16 boxes per chunk: z = acc + bias (tmC) and h = gelu(z) (tmC2)
  mbar_wait(tfull, ph);
  tc_fence_after();
  int it = 0;
#pragma unroll 1
  for (int c = half; c < bn / 32; c += 2, ++it) {
    const int n0 = n_base + c * 32;
    float4 b[8];
#pragma unroll
    for (int g = 0; g < 8; ++g) b[g] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p.bias && n0 < p.N) {
#pragma unroll
      for (int g = 0; g < 8; ++g) b[g] = __ldg(reinterpret_cast<const float4*>(p.bias + n0) + g);   // warp-uniform address
    }
    uint32_t r[32];
    tmem_ld32(t_base + c * 32, r);
    tmem_ld_wait();
    if (n0 >= p.N) break;   // warp-uniform
    uint8_t* buf = slot + ((f32 || gelu || (it & 1) == 0) ? 0 : 2048);
    if (lane == 0) {        // the box about to be overwritten must have been read by its store
      if (f32 || gelu) bulk_wait_read<0>(); else bulk_wait_read<1>();
    }
    __syncwarp();
    if (gelu) {
      // h is computed from the bf16-rounded z, exactly what the stand-alone GELU kernel reads back (same results either way)
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        uint4 w, hh;
        uint32_t* wz = reinterpret_cast<uint32_t*>(&w);
        uint32_t* wh = reinterpret_cast<uint32_t*>(&hh);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 bb = b[2 * g + (j >> 1)];
          const float b0 = (j & 1) ? bb.z : bb.x, b1 = (j & 1) ? bb.w : bb.y;
          wz[j] = pack_bf16x2(__uint_as_float(r[8 * g + 2 * j]) + b0, __uint_as_float(r[8 * g + 2 * j + 1]) + b1);
          const float2 zr = unpack_bf16x2(wz[j]);
          wh[j] = pack_bf16x2(gelu_fast(zr.x), gelu_fast(zr.y));
        }
        const int off = lane * 64 + ((g ^ ((lane >> 1) & 3)) << 4);                          // SWIZZLE_64B
        *reinterpret_cast<uint4*>(buf + off) = w;
        *reinterpret_cast<uint4*>(buf + 2048 + off) = hh;
      }
    } else if (f32) {
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        float4 v;
        v.x = s * (__uint_as_float(r[4 * g + 0]) + b[g].x);
        v.y = s * (__uint_as_float(r[4 * g + 1]) + b[g].y);
        v.z = s * (__uint_as_float(r[4 * g + 2]) + b[g].z);
        v.w = s * (__uint_as_float(r[4 * g + 3]) + b[g].w);
        *reinterpret_cast<float4*>(buf + lane * 128 + ((g ^ (lane & 7)) << 4)) = v;       // SWIZZLE_128B
      }
    } else {
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        uint4 w;
        w.x = pack_bf16x2(s * (__uint_as_float(r[8 * g + 0]) + b[2 * g].x), s * (__uint_as_float(r[8 * g + 1]) + b[2 * g].y));
        w.y = pack_bf16x2(s * (__uint_as_float(r[8 * g + 2]) + b[2 * g].z), s * (__uint_as_float(r[8 * g + 3]) + b[2 * g].w));
        w.z = pack_bf16x2(s * (__uint_as_float(r[8 * g + 4]) + b[2 * g + 1].x), s * (__uint_as_float(r[8 * g + 5]) + b[2 * g + 1].y));
        w.w = pack_bf16x2(s * (__uint_as_float(r[8 * g + 6]) + b[2 * g + 1].z), s * (__uint_as_float(r[8 * g + 7]) + b[2 * g + 1].w));
        *reinterpret_cast<uint4*>(buf + lane * 64 + ((g ^ ((lane >> 1) & 3)) << 4)) = w;   // SWIZZLE_64B
      }
    }
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) {
      if (p.tma_store == 2) tma_reduce_add_3d(tmC, buf, n0, m_blk * BM + q * 32, 0);
      else tma_store_3d(tmC, buf, n0, m_blk * BM + q * 32, split);
      if (gelu) tma_store_3d(tmC2, buf + 2048, n0, m_blk * BM + q * 32, 0);
      bulk_commit();
    }
  }
  if (lane == 0) bulk_wait_read<0>();   // staging slot is reused by the next tile
  __syncwarp();
}

// ------------------------------------------------------------------------------------------------
// fp32 residual epilogue on TMA:   out[map(m), n] = s(m) * (acc + bias[n]) + aux[map(m), n]
// The residual box (32 rows x 32 fp32) is TMA-loaded into the warp's staging buffer — requested before the accumulator
// is even waited for, the next chunk's one chunk ahead — added to in place (thread = row, same swizzled 16-byte slots
// the load wrote), and leaves through a TMA store from the same buffer.  Rows reach memory through a 3-D tensor map
// (col, row in sample, sample); for the spatial regrouping the 32 rows of a box are T stream rows apart, which the map
// expresses as box height 32*T with element stride T (a rank-4 map (col, t, p, b) is not encodable: TMA strides must nest,
// and the sample stride (1 + P*T) * D is not a multiple of the patch stride T * D — such maps fault on the first store).
// So the temporal ('b (p t)') and spatial ('(b t) p' + replicated cls) regroupings cost nothing:
// a 32-row group that crosses a period boundary is served by two boxes (out-of-range rows are clipped by the TMA unit;
// for the loads the two segments land in the two buffers and every row reads the one its segment wrote).
// Per warp: 2 buffers of 4 KiB + 2 mbarriers.
// ------------------------------------------------------------------------------------------------
constexpr int RES_SLOT_BYTES = 8192;      // per epilogue warp: two 32 x 32 fp32 boxes

// ------------------------------------------------------------------------------------------------
// dGELU epilogue on TMA (autograd of nn.GELU fused into the FC2 data-gradient GEMM):
//   out_bf16[m, n] = acc[m, n] * gelu'(z_bf16[m, n])        plain rows, no maps
// Same mechanics as the residual epilogue with 2 KiB bf16 boxes (64B swizzle): the z box is TMA-loaded one chunk ahead into
// one of the warp's two buffers, multiplied into the accumulator in place and TMA-stored from the same buffer.
// ------------------------------------------------------------------------------------------------
template <int BN>
__device__ __forceinline__ void epilogue_tile_tma_dgelu(const GemmDev& p, const CUtensorMap* tmC, const CUtensorMap* tmZ,
                                                        uint8_t* slot, uint64_t* aux_bar, uint32_t (&aux_use)[2], uint32_t t_base,
                                                        int m_blk, int n_base, int bn, int q, int half, int lane, uint64_t* tfull,
                                                        uint32_t ph) {
  const int m0 = m_blk * BM + q * 32;
  const bool any = m0 < p.M;
  const int nchunks = (bn / 32 - half + 1) / 2;
  auto chunk_cols_ok = [&](int it) { return n_base + (half + 2 * it) * 32 < p.N; };
  auto request = [&](int it, int buf) {
    mbar_arrive_expect_tx(&aux_bar[buf], 2048);
    tma_load_3d(slot + buf * 2048, tmZ, &aux_bar[buf], n_base + (half + 2 * it) * 32, m0, 0);
  };
  if (lane == 0 && any) {
    if (nchunks > 0 && chunk_cols_ok(0)) request(0, 0);
    if (nchunks > 1 && chunk_cols_ok(1)) request(1, 1);
  }
  mbar_wait(tfull, ph);
  tc_fence_after();
  if (!any) return;
#pragma unroll 1
  for (int it = 0; it < nchunks; ++it) {
    const int c = half + 2 * it;
    const int n0 = n_base + c * 32;
    if (n0 >= p.N) break;   // warp-uniform
    uint32_t r[32];
    tmem_ld32(t_base + c * 32, r);
    tmem_ld_wait();
    const int cur = it & 1;
    if (cur == 0) { mbar_wait(&aux_bar[0], aux_use[0] & 1); ++aux_use[0]; }
    else { mbar_wait(&aux_bar[1], aux_use[1] & 1); ++aux_use[1]; }
    uint8_t* buf = slot + cur * 2048;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      uint4* cell = reinterpret_cast<uint4*>(buf + lane * 64 + ((g ^ ((lane >> 1) & 3)) << 4));     // SWIZZLE_64B
      const uint4 zz = *cell;
      const uint32_t zw[4] = {zz.x, zz.y, zz.z, zz.w};
      uint32_t ow[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 z2 = unpack_bf16x2(zw[j]);
        ow[j] = pack_bf16x2(__uint_as_float(r[8 * g + 2 * j]) * dgelu_fast(z2.x),
                            __uint_as_float(r[8 * g + 2 * j + 1]) * dgelu_fast(z2.y));
      }
      *cell = make_uint4(ow[0], ow[1], ow[2], ow[3]);
    }
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) {
      tma_store_3d(tmC, buf, n0, m0, 0);
      bulk_commit();
      if (it + 2 < nchunks && chunk_cols_ok(it + 2)) {
        bulk_wait_read<0>();
        request(it + 2, cur);
      }
    }
    __syncwarp();
  }
  if (lane == 0) bulk_wait_read<0>();
  __syncwarp();
}

// aux_use[j]: how often buffer j's mbarrier has completed so far (its wait parity); carried across tiles by the caller
template <int BN>
__device__ __forceinline__ void epilogue_tile_tma_res(const GemmDev& p, const CUtensorMap* tmC, const CUtensorMap* tmX,
                                                      uint8_t* slot, uint64_t* aux_bar, uint32_t (&aux_use)[2], uint32_t t_base,
                                                      int m_blk, int n_base, int bn, int q, int half, int lane, uint64_t* tfull,
                                                      uint32_t ph) {
  const int m0 = m_blk * BM + q * 32;
  const int row = m0 + lane;
  const float s = (row < p.M && p.row_scale) ? p.row_scale[row] : 1.0f;
  // segments of this 32-row group: rows [inner0, period) of period `outer0` and, when the group crosses the boundary, rows
  // [0, ...) of period outer0 + 1.  A segment is "live" when its box overlaps the tensor at all; boxes that would lie fully
  // outside (rows >= M, or a second segment consisting of the special row only) are never issued.
  const int outer0 = m0 / p.map_period, inner0 = m0 - outer0 * p.map_period;
  const int n_outer = p.M / p.map_period, pcount = p.map_period - p.map_skip;
  int seg_t[2], seg_b[2], seg_p[2];
  bool live[2];
#pragma unroll
  for (int sgm = 0; sgm < 2; ++sgm) {
    const int outer = outer0 + sgm;
    seg_b[sgm] = outer / p.map_tcount;
    seg_t[sgm] = outer - seg_b[sgm] * p.map_tcount;
    seg_p[sgm] = inner0 - sgm * p.map_period - p.map_skip;
    live[sgm] = outer < n_outer && seg_p[sgm] + 31 >= 0 && seg_p[sgm] < pcount && (sgm == 0 || inner0 + 32 > p.map_period);
  }
  const bool any = live[0] || live[1];
  const int only = live[0] ? 0 : 1;                    // the segment served by TMA
  // TMA moves the group's box only when it is ONE segment that starts inside the tensor (rows past the end are clipped by
  // the unit — proven on hardware).  Groups that straddle a period boundary, or start on a special row (negative start
  // coordinate), go row by row: thread = row, 128 contiguous bytes per chunk, addresses from the out_row / aux_row arrays.
  // (Boxes with negative start coordinates raised "illegal instruction" at the store on the B200, with rank-3 and rank-4 maps.)
  const bool direct = any && ((live[0] && live[1]) || seg_p[only] < 0);
  // this lane's row: which segment, special?
  const int my_seg = (inner0 + lane >= p.map_period) ? 1 : 0;
  const int my_inner = inner0 + lane - my_seg * p.map_period;
  const bool special = row < p.M && my_inner < p.map_skip;
  const int nchunks = (bn / 32 - half + 1) / 2;          // chunks c = half, half + 2, ... of this warpgroup
  auto chunk_cols_ok = [&](int it) { return n_base + (half + 2 * it) * 32 < p.N; };
  // row coordinate of a segment in the (col, row-in-sample, sample) map: rows of one frame sit map_tcount apart (the box
  // walks them through the map's element stride), so segment row p of frame t starts at p * map_tcount + t
  auto request = [&](int it, int buf) {        // lane 0: residual box of chunk `it` -> buffer buf
    const int n = n_base + (half + 2 * it) * 32;
    tma_load_3d(slot + buf * 4096, tmX, &aux_bar[buf], n, seg_p[only] * p.map_tcount + seg_t[only], seg_b[only]);
  };
  // chunk `it` lives in buffer it & 1 and chunk it + 1 is prefetched into the other buffer
  if (lane == 0 && any && !direct) {
    if (nchunks > 0 && chunk_cols_ok(0)) { mbar_arrive_expect_tx(&aux_bar[0], 4096); request(0, 0); }
    if (nchunks > 1 && chunk_cols_ok(1)) { mbar_arrive_expect_tx(&aux_bar[1], 4096); request(1, 1); }
  }
  long long o_off = -1, a_off = -1;             // direct mode: this lane's output / addend row (element offsets), -1 = none
  if (direct && row < p.M && !special) {
    const int orow = p.out_row ? p.out_row[row] : row;
    const int arow = p.aux_row ? p.aux_row[row] : row;
    if (orow >= 0) o_off = (long long)orow * p.ldo;
    if (arow >= 0) a_off = (long long)arow * p.ldaux;
  }
  mbar_wait(tfull, ph);
  tc_fence_after();
  if (!any) return;        // nothing of this group exists (rows >= M): the caller still releases the accumulator
#pragma unroll 1
  for (int it = 0; it < nchunks; ++it) {
    const int c = half + 2 * it;
    const int n0 = n_base + c * 32;
    if (n0 >= p.N) break;   // warp-uniform
    float4 b[8];
#pragma unroll
    for (int g = 0; g < 8; ++g) b[g] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p.bias) {
#pragma unroll
      for (int g = 0; g < 8; ++g) b[g] = __ldg(reinterpret_cast<const float4*>(p.bias + n0) + g);   // warp-uniform address
    }
    uint32_t r[32];
    tmem_ld32(t_base + c * 32, r);
    tmem_ld_wait();
    if (direct) {
      float4 v[8];
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        if (a_off >= 0) a = *reinterpret_cast<const float4*>(static_cast<const float*>(p.aux) + a_off + n0 + 4 * g);
        if (p.bias2) {
          const float4 b2 = __ldg(reinterpret_cast<const float4*>(p.bias2 + n0) + g);
          a.x += b2.x; a.y += b2.y; a.z += b2.z; a.w += b2.w;
        }
        v[g].x = fmaf(s, __uint_as_float(r[4 * g + 0]) + b[g].x, a.x);
        v[g].y = fmaf(s, __uint_as_float(r[4 * g + 1]) + b[g].y, a.y);
        v[g].z = fmaf(s, __uint_as_float(r[4 * g + 2]) + b[g].z, a.z);
        v[g].w = fmaf(s, __uint_as_float(r[4 * g + 3]) + b[g].w, a.w);
      }
      float* dst = nullptr;
      if (special) {
        if (p.special_out) dst = p.special_out + (long long)(outer0 + my_seg) * p.special_ld + n0;
      } else if (o_off >= 0) {
        dst = static_cast<float*>(p.out) + o_off + n0;
      }
      if (dst) {
#pragma unroll
        for (int g = 0; g < 8; ++g) *reinterpret_cast<float4*>(dst + 4 * g) = v[g];
      }
      continue;
    }
    const int cur = it & 1;
    // residual box of this chunk has landed
    if (cur == 0) { mbar_wait(&aux_bar[0], aux_use[0] & 1); ++aux_use[0]; }
    else { mbar_wait(&aux_bar[1], aux_use[1] & 1); ++aux_use[1]; }
    uint8_t* buf = slot + cur * 4096;
    float4 v[8];
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      float4 a = *reinterpret_cast<const float4*>(buf + lane * 128 + ((g ^ (lane & 7)) << 4));    // SWIZZLE_128B
      if (p.bias2) {
        const float4 b2 = __ldg(reinterpret_cast<const float4*>(p.bias2 + n0) + g);
        a.x += b2.x; a.y += b2.y; a.z += b2.z; a.w += b2.w;
      }
      v[g].x = fmaf(s, __uint_as_float(r[4 * g + 0]) + b[g].x, a.x);
      v[g].y = fmaf(s, __uint_as_float(r[4 * g + 1]) + b[g].y, a.y);
      v[g].z = fmaf(s, __uint_as_float(r[4 * g + 2]) + b[g].z, a.z);
      v[g].w = fmaf(s, __uint_as_float(r[4 * g + 3]) + b[g].w, a.w);
    }
    if (special && p.special_out) {
      // the next period's special row at the tail of the box: its box row was clipped to zeros by the load and will be
      // clipped again by the store; the value goes to the side rows
      float* dst = p.special_out + (long long)(outer0 + my_seg) * p.special_ld + n0;
#pragma unroll
      for (int g = 0; g < 8; ++g) *reinterpret_cast<float4*>(dst + 4 * g) = v[g];
    }
#pragma unroll
    for (int g = 0; g < 8; ++g) *reinterpret_cast<float4*>(buf + lane * 128 + ((g ^ (lane & 7)) << 4)) = v[g];
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) {
      tma_store_3d(tmC, buf, n0, seg_p[only] * p.map_tcount + seg_t[only], seg_b[only]);
      bulk_commit();
      // refill: the buffer just stored from must have been read out by its store first
      if (it + 2 < nchunks && chunk_cols_ok(it + 2)) {
        bulk_wait_read<0>();
        mbar_arrive_expect_tx(&aux_bar[cur], 4096);
        request(it + 2, cur);
      }
    }
    __syncwarp();
  }
  if (lane == 0) bulk_wait_read<0>();   // both buffers are reused by the next tile
  __syncwarp();
}

}  // namespace vt
