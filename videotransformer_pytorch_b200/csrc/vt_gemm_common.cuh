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
                           //    reduce-added into the (pre-zeroed) output by TMA
};

// Drain accumulator tile (m_blk, n_blk) of this CTA: TMEM columns [t_base, t_base + BN) of lane quadrant q.
// Called by the 8 epilogue warps; waits on `tfull` (parity `ph`) after issuing the first operand prefetch.
template <int BN>
__device__ __forceinline__ void epilogue_tile(const GemmDev& p, float* stg, uint32_t t_base, int m_blk, int n_blk, int split,
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
    const int n = n_blk * BN + c * 32 + cg * 4;
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
  for (int c = half; c < BN / 32; c += 2) {
    uint32_t r[32];
    tmem_ld32(t_base + c * 32, r);
    tmem_ld_wait();
    const int n = n_blk * BN + c * 32 + cg * 4;
    if (n_blk * BN + c * 32 >= p.N) break;   // warp-uniform
#pragma unroll
    for (int j = 0; j < 32; ++j) stg[lane * EPI_PITCH + j] = __uint_as_float(r[j]);
    uint4 cur[8];
#pragma unroll
    for (int it = 0; it < 8; ++it) cur[it] = pre[it];
    if ((f32_aux || z_aux) && c + 2 < BN / 32) prefetch(c + 2);
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
__device__ __forceinline__ void epilogue_tile_tma(const GemmDev& p, const CUtensorMap* tmC, uint8_t* slot, uint32_t t_base,
                                                  int m_blk, int n_blk, int split, int q, int half, int lane, uint64_t* tfull,
                                                  uint32_t ph) {
  const int row = m_blk * BM + q * 32 + lane;
  const float s = (row < p.M && p.row_scale) ? p.row_scale[row] : 1.0f;
  const bool f32 = p.epi == VT_EPI_F32;
  mbar_wait(tfull, ph);
  tc_fence_after();
  int it = 0;
#pragma unroll 1
  for (int c = half; c < BN / 32; c += 2, ++it) {
    const int n0 = n_blk * BN + c * 32;
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
    uint8_t* buf = slot + ((f32 || (it & 1) == 0) ? 0 : 2048);
    if (lane == 0) {        // the box about to be overwritten must have been read by its store
      if (f32) bulk_wait_read<0>(); else bulk_wait_read<1>();
    }
    __syncwarp();
    if (f32) {
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
      bulk_commit();
    }
  }
  if (lane == 0) bulk_wait_read<0>();   // staging slot is reused by the next tile
  __syncwarp();
}

}  // namespace vt
