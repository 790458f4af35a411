// CTA-pair GEMM: tcgen05.mma.cta_group::2 — two SMs of one cluster compute a 256 x BN tile together.
// Each CTA stages only its own 128 rows of A and HALF of the B tile (BN/2 rows); the pair's tensor cores
// read both halves, so a pipeline stage costs 16 KiB + BN/2*128 B per SM instead of 16 KiB + BN*128 B.
// With BN = 256 that is 32 KiB/stage -> 6 stages in flight per SM (vs 4 x 48 KiB for the single-CTA kernel),
// which is what the single-CTA kernel's mainloop was starved of (shared-memory bytes in flight, not L2
// bandwidth — see profiles/).  Same epilogue (vt_gemm_common.cuh), same K-/MN-major operand support.
//
// Pair protocol (leader = cluster rank 0):
//   * both producers issue their TMA loads with .cta_group::2, completing on the LEADER's "full" barrier
//     (peer bit of the barrier address cleared); the leader arms it with the bytes of both CTAs;
//   * only the leader's MMA thread issues tcgen05.mma.cta_group::2 (M = 256) and commits with
//     .multicast::cluster to the "empty" (stage free) and "tmem full" barriers of BOTH CTAs;
//   * each CTA's 8 epilogue warps drain their own TMEM (128 lanes x BN columns) and arrive on the leader's
//     "tmem empty" barrier (the peer through a mapa-translated remote arrive).
#include <stdlib.h>
#include <string.h>

#include "vt_gemm_common.cuh"

namespace vt {

int make_tmap_bf16_2d(CUtensorMap* map, const void* base, long long rows, long long cols, long long ld, int box_rows);
int launch_reduce_rows(const float* in, float* out, long long stride, int S, long long n, int accumulate, float scale,
                       cudaStream_t st);
int setup_out_map(const vt_gemm_params* q, GemmDev& d, CUtensorMap* tmC, bool in_place, CUtensorMap* tmC2);
int setup_res_maps(const vt_gemm_params* q, GemmDev& d, CUtensorMap* tmC, CUtensorMap* tmX);
struct Schedule { int full_units, tail_units, tail_bn, tail_mp; double makespan; };
Schedule plan_units(int M, int N, int bn, int rows_per_macro, int splits, int slots, int tail_bn_cand, int tail_mode);
bool splitk_in_place(const vt_gemm_params* q);
bool tail_units_enabled();
int splitk_zero(const vt_gemm_params* q, cudaStream_t st);

template <int BN, bool RES>
struct Gemm2Cfg {
  static constexpr int A_BYTES = BM * BK * 2;             // this CTA's 128 rows
  static constexpr int B_BYTES = (BN / 2) * BK * 2;       // this CTA's half of the B tile
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = RES ? ((BN == 256) ? 5 : 6) : ((BN == 256) ? 6 : 8);
  static constexpr int TMEM_COLS = (BN == 128) ? 256 : 512;
  static constexpr int STAGING_BYTES = RES ? EPI_WARPS * RES_SLOT_BYTES : EPI_WARPS * 32 * EPI_PITCH * 4;
  static constexpr int BAR_BYTES = 512;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + BAR_BYTES + STAGING_BYTES;
  static_assert(SMEM_BYTES <= 232448, "shared memory budget");
};

constexpr uint32_t PEER_BIT_MASK = 0xFEFFFFFFu;   // clears the CTA-in-pair bit of a shared::cluster address

__device__ __forceinline__ void tma_load_2d_2cta(void* smem_dst, const CUtensorMap* map, uint32_t leader_bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(leader_bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma2_bf16_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma2_commit_mc(uint64_t* bar, uint16_t cta_mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(cta_mask)
      : "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_alloc_2cta(uint32_t* smem_result) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "n"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_dealloc_2cta(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}
// arrive on the barrier at the same offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t rank) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}\n" ::"r"(smem_u32(bar)),
      "r"(rank)
      : "memory");
}

template <int BN, bool RES>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm2_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                     const __grid_constant__ CUtensorMap tmBt, const __grid_constant__ CUtensorMap tmC,
                     const __grid_constant__ CUtensorMap tmX, const GemmDev p) {
  using Cfg = Gemm2Cfg<BN, RES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* staging = smem + Cfg::STAGES * Cfg::STAGE_BYTES;                       // 1024-aligned epilogue staging
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(staging + Cfg::STAGING_BYTES);              // used in the leader
  uint64_t* empty_bar = full_bar + Cfg::STAGES;                                              // local
  uint64_t* tfull_bar = empty_bar + Cfg::STAGES;                                             // local
  uint64_t* tempty_bar = tfull_bar + 2;                                                      // used in the leader
  uint64_t* aux_bar = tempty_bar + 2;                      // [EPI_WARPS][2] residual-box barriers (RES only), local
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(aux_bar + 2 * EPI_WARPS);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank();
  const bool leader = crank == 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (p.tail_bn) tma_prefetch_desc(&tmBt);
    if (p.tma_store) tma_prefetch_desc(&tmC);
    if (RES || p.tma_store >= 4) tma_prefetch_desc(&tmX);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < Cfg::STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 2 * EPI_WARPS);   // epilogue warps of both CTAs
    }
    if (RES || p.tma_store == 5) {
      for (int i = 0; i < 2 * EPI_WARPS; ++i) mbar_init(&aux_bar[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_2cta<Cfg::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int total_units = p.full_units + p.tail_units;
  const int unit0 = blockIdx.x >> 1, unit_step = gridDim.x >> 1;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0, phase = 0;
      for (int unit = unit0; unit < total_units; unit += unit_step) {
        const GemmUnit u = decode_unit<BN>(p, unit);
        const int m_blk = u.mp * 2 + (int)crank;
        const bool narrow = u.bn != BN;          // tail unit: own B map with a (tail_bn / 2)-row box
        const int kb0 = u.kb0, kb1 = u.kb1;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sA = smem + stage * Cfg::STAGE_BYTES;
          uint8_t* sB = sA + Cfg::A_BYTES;
          const uint32_t lbar = smem_u32(&full_bar[stage]) & PEER_BIT_MASK;
          if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * (Cfg::A_BYTES + (u.bn / 2) * BK * 2));
          if (!p.a_mn) {
            tma_load_2d_2cta(sA, &tmA, lbar, kb * BK, m_blk * BM);
          } else {
#pragma unroll
            for (int c = 0; c < BM / 64; ++c) tma_load_2d_2cta(sA + c * CHUNK_BYTES, &tmA, lbar, m_blk * BM + c * 64, kb * BK);
          }
          const int nb0 = u.n0 + (int)crank * (u.bn / 2);   // first B row (= output column) of my half
          if (!p.b_mn) {
            tma_load_2d_2cta(sB, narrow ? &tmBt : &tmB, lbar, kb * BK, nb0);
          } else {
            for (int c = 0; c < u.bn / 128; ++c) tma_load_2d_2cta(sB + c * CHUNK_BYTES, &tmB, lbar, nb0 + c * 64, kb * BK);
          }
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (leader && lane == 0) {
      int stage = 0, phase = 0, acc = 0, acc_phase = 0;
      for (int unit = unit0; unit < total_units; unit += unit_step) {
        const GemmUnit u = decode_unit<BN>(p, unit);
        const uint32_t idesc = make_idesc_bf16(2 * BM, (uint32_t)u.bn, (uint32_t)p.a_mn, (uint32_t)p.b_mn);
        const int kb0 = u.kb0, kb1 = u.kb1;
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + stage * Cfg::STAGE_BYTES);
          const uint32_t b_addr = a_addr + Cfg::A_BYTES;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t adesc = p.a_mn ? sdesc_mnmajor(a_addr + k * 2048, CHUNK_BYTES) : sdesc_kmajor(a_addr + k * 32);
            const uint64_t bdesc = p.b_mn ? sdesc_mnmajor(b_addr + k * 2048, CHUNK_BYTES) : sdesc_kmajor(b_addr + k * 32);
            umma2_bf16_ss(d_tmem, adesc, bdesc, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          umma2_commit_mc(&empty_bar[stage], 3);    // stage free in both CTAs
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        }
        umma2_commit_mc(&tfull_bar[acc], 3);        // accumulators ready in both CTAs
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    const int q = warp & 3;
    const int half = (warp - 4) >> 2;
    float* stg = reinterpret_cast<float*>(staging) + (warp - 4) * (32 * EPI_PITCH);
    uint8_t* slot = staging + (warp - 4) * (RES ? RES_SLOT_BYTES : 4096);
    uint32_t aux_use[2] = {0u, 0u};
    int acc = 0, acc_phase = 0;
    for (int unit = unit0; unit < total_units; unit += unit_step) {
      const GemmUnit u = decode_unit<BN>(p, unit);
      const int m_blk = u.mp * 2 + (int)crank;
      const uint32_t t_base = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN);
      if constexpr (RES) {
        epilogue_tile_tma_res<BN>(p, &tmC, &tmX, slot, aux_bar + 2 * (warp - 4), aux_use, t_base, m_blk, u.n0, u.bn, q, half, lane,
                                  &tfull_bar[acc], (uint32_t)acc_phase);
      } else {
        if (p.tma_store == 5) epilogue_tile_tma_dgelu<BN>(p, &tmC, &tmX, slot, aux_bar + 2 * (warp - 4), aux_use, t_base, m_blk, u.n0, u.bn, q, half, lane, &tfull_bar[acc], (uint32_t)acc_phase);
        else if (p.tma_store) epilogue_tile_tma<BN>(p, &tmC, &tmX, slot, t_base, m_blk, u.n0, u.bn, u.split, q, half, lane, &tfull_bar[acc], (uint32_t)acc_phase);
        else epilogue_tile<BN>(p, stg, t_base, m_blk, u.n0, u.bn, u.split, q, half, lane, &tfull_bar[acc], (uint32_t)acc_phase);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (leader) mbar_arrive(&tempty_bar[acc]);
        else mbar_arrive_remote(&tempty_bar[acc], 0);
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_2cta<Cfg::TMEM_COLS>(tmem_base);
  }
}

template <int BN, bool RES>
static int launch_gemm2_t(const vt_gemm_params* q, GemmDev& d, cudaStream_t st) {
  using Cfg = Gemm2Cfg<BN, RES>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm2_tcgen05_kernel<BN, RES>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    VT_REQUIRE(e == cudaSuccess, "gemm2: cudaFuncSetAttribute(smem=%d) failed: %s", Cfg::SMEM_BYTES, cudaGetErrorString(e));
    attr_set = true;
  }
  CUtensorMap tmA, tmB, tmBt, tmX;
  memset(&tmBt, 0, sizeof(tmBt));
  memset(&tmX, 0, sizeof(tmX));
  int rc;
  if (!q->a_mn_major) rc = make_tmap_bf16_2d(&tmA, q->a, q->M, q->K, q->lda, BM);
  else rc = make_tmap_bf16_2d(&tmA, q->a, q->K, q->M, q->lda, BK);
  if (rc) return rc;
  if (!q->b_mn_major) rc = make_tmap_bf16_2d(&tmB, q->b, q->N, q->K, q->ldb, BN / 2);
  else rc = make_tmap_bf16_2d(&tmB, q->b, q->K, q->N, q->ldb, BK);
  if (rc) return rc;

  d.num_m = (q->M + BM - 1) / BM;
  d.num_mp = (d.num_m + 1) / 2;
  d.num_n = (q->N + BN - 1) / BN;
  d.kblocks = (q->K + BK - 1) / BK;
  const int pairs = persistent_sm_count() / 2;
  const long long tile_out = (long long)q->M * q->N;
  int splits = 1;
  if (q->epilogue == VT_EPI_F32 && q->workspace && !q->out_row && !q->aux && !q->row_scale && !q->bias) {
    splits = q->force_splits > 0 ? q->force_splits : (d.splits > 0 ? d.splits : 1);
    const long long max_by_ws = q->workspace_bytes / (tile_out * 4);
    if (splits > max_by_ws) splits = (int)max_by_ws;
    if (splits > d.kblocks) splits = d.kblocks;
    if (splits < 1) splits = 1;
  }
  d.splits = splits;
  void* final_out = d.out;
  const bool in_place = splits > 1 && splitk_in_place(q);
  d.split_stride = 0;
  if (in_place) {
    rc = q->out_zeroed ? 0 : splitk_zero(q, st);
    if (rc) return rc;
  } else if (splits > 1) {
    VT_REQUIRE(q->ldo == q->N, "vt_gemm: split-K requires ldo == N");
    d.out = q->workspace;
    d.ldo = q->N;
    d.split_stride = tile_out;
  }
  CUtensorMap tmC;
  if (RES) rc = setup_res_maps(q, d, &tmC, &tmX);
  else rc = setup_out_map(q, d, &tmC, in_place, &tmX);
  if (rc) return rc;
  const Schedule sch = plan_units(q->M, q->N, BN, 2 * BM, splits, pairs, 128, tail_units_enabled() ? q->force_tail : 1);
  d.full_units = sch.full_units; d.tail_units = sch.tail_units; d.tail_bn = sch.tail_bn; d.tail_mp = sch.tail_mp;
  if (d.tail_bn && !q->b_mn_major) {
    rc = make_tmap_bf16_2d(&tmBt, q->b, q->N, q->K, q->ldb, d.tail_bn / 2);
    if (rc) return rc;
  }
  const int units = d.full_units + d.tail_units;
  const int grid = 2 * (units < pairs ? units : pairs);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(GEMM_THREADS);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t le = cudaLaunchKernelEx(&cfg, gemm2_tcgen05_kernel<BN, RES>, tmA, tmB, tmBt, tmC, tmX, d);
  if (le != cudaSuccess) {
    set_error("gemm2_tcgen05_kernel: cudaLaunchKernelEx failed: %s", cudaGetErrorString(le));
    return 2;
  }
  rc = check_launch("gemm2_tcgen05_kernel");
  if (rc) return rc;
  if (splits > 1 && !in_place)
    return launch_reduce_rows(static_cast<const float*>(q->workspace), static_cast<float*>(final_out), tile_out, splits,
                              tile_out, 0, 1.0f, st);
  return 0;
}

// entry used by vt_gemm(): bn in {128, 256}
int launch_gemm2(const vt_gemm_params* q, GemmDev& d, int bn, bool res, cudaStream_t st) {
  if (res) {
    if (bn == 256) return launch_gemm2_t<256, true>(q, d, st);
    return launch_gemm2_t<128, true>(q, d, st);
  }
  if (bn == 256) return launch_gemm2_t<256, false>(q, d, st);
  return launch_gemm2_t<128, false>(q, d, st);
}

}  // namespace vt
