// CTA-pair h16 GEMM for sm_100a:  out[M,N] = epi(A[M,K] . W[N,K]^T + bias),  N % 256 == 0.
//
// Two CTAs of one cluster (the two SMs of a TPC) cooperate on a 256 x 256 output tile with
// tcgen05.mma.cta_group::2 (256 x 256 x 16): each CTA stages its own 128 rows of A and only HALF of
// the W tile (128 of the 256 rows), i.e. 32 KiB instead of 48 KiB per 64-wide K block.  That is what
// lifts the single-CTA kernel's ceiling: there, TMA writes (48 KiB) plus the tensor core's operand
// reads (48 KiB) per K block exceed what shared memory moves in the 512 clk the MMAs take.
//
// Roles per CTA (384 threads):
//   warp 0      TMA producer: A rows of this CTA + its half of W.  BOTH CTAs credit the bytes to the
//               LEADER's full barrier (cp.async.bulk.tensor ... .cta_group::2 with a mapa'd barrier
//               address), so no CTA ever forwards an arrive: a releasing remote mbarrier.arrive costs
//               the issuing thread 500-1500 clk and was the whole story of the first, slow version.
//   warp 1      leader CTA only: MMA issuer; tcgen05.commit multicasts "stage free" / "accumulator
//               full" to the barriers of both CTAs
//   warp 2      TMEM allocator (cta_group::2: one warp in each CTA)
//   warps 4-11  epilogue of this CTA's 128 rows: TMEM -> bias/GELU/residual | SwiGLU -> swizzled
//               staging (two tiles per warp) -> TMA store; accumulator columns go back to the leader
//               with a RELAXED remote arrive (nothing but TMEM reads has to be ordered)
#pragma once

#include "common.cuh"
#include "gemm.cuh"

namespace b2e {

constexpr int G2_BN = 256;
constexpr int G2_THREADS = 384;
constexpr int G2_EPI_WARPS = 8;

template <int STAGES>
struct Gemm2Cfg {
  static constexpr int A_BYTES = 128 * GEMM_BK * 2;            // this CTA's 128 rows of A
  static constexpr int B_BYTES = (G2_BN / 2) * GEMM_BK * 2;    // this CTA's half of the W tile
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;        // 32 KiB
  static constexpr int STAGING_OFFSET = STAGES * STAGE_BYTES;
  static constexpr int STAGING_PER_WARP = 2 * GEMM_STAGING_BYTES;   // one tile per 64-column chunk
  static constexpr int BAR_OFFSET = STAGING_OFFSET + G2_EPI_WARPS * STAGING_PER_WARP;
  static constexpr int SMEM_BYTES = BAR_OFFSET + 256;
  static constexpr int TMEM_COLS = 2 * G2_BN;
  static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KiB per-CTA shared memory limit");
};

// profiling aids (b2e_debug_set_clock_buffer / b2e_debug_set_pair_flags)
__device__ long long* g_gemm2_clock = nullptr;   // CTAs 0/1: clock64() timelines, [cta*2 + role][256]
__device__ int g_gemm2_flags = 0;                // 1 skip epilogue math+stores, 2 no MMAs, 4 no TMA loads

// TL: the profiling instantiation (clock64 timelines, experiment knobs); production kernels carry neither the
// volatile clock reads nor the flag tests.
template <int STAGES, int EPI, bool TL = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(G2_THREADS, 1)
gemm2_h16_pair_kernel(const __grid_constant__ CUtensorMap tm_a,    // [M,K], box 64 x 128
                       const __grid_constant__ CUtensorMap tm_b,    // [N,K], box 64 x 128
                       const __grid_constant__ CUtensorMap tm_out,  // [M,N] (SwiGLU: [M,N/2]), box 64 x 32
                       const float* __restrict__ bias, const h16* __restrict__ resid, int M, int N,
                       int K, const int* __restrict__ m_dev) {
  // m_dev (nullable): the row count lives on the device (packed token layout, pack.cuh); M is then only the
  // upper bound the grid and the tensor maps were sized for
  if (m_dev != nullptr) M = __ldg(m_dev);
  using Cfg = Gemm2Cfg<STAGES>;
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t smem_base = smem_u32(smem);
  if ((smem_base & 1023u) != 0) __trap();

  const uint32_t full_bar = smem_base + Cfg::BAR_OFFSET;   // used in the leader only
  const uint32_t empty_bar = full_bar + 8u * STAGES;
  const uint32_t tfull_bar = empty_bar + 8u * STAGES;
  const uint32_t tempty_bar = tfull_bar + 16u;             // used in the leader only
  volatile uint32_t* tmem_slot =
      reinterpret_cast<volatile uint32_t*>(smem + Cfg::BAR_OFFSET + 8 * (2 * STAGES + 4));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int flags = TL ? g_gemm2_flags : 0;   // experiment knobs, read once (compiled out of production kernels)
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_b);
    tma_prefetch_desc(&tm_out);
  }
  if (warp == 1 && elect_one()) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar + 8u * s, 1);   // leader: its producer's expect_tx(bytes of BOTH CTAs)
      mbar_init(empty_bar + 8u * s, 1);  // one multicast commit
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar + 8u * s, 1);
      mbar_init(tempty_bar + 8u * s, 2 * G2_EPI_WARPS);  // leader: epilogue warps of both CTAs
    }
    mbar_fence_init();
  }
  cluster_sync_all();  // barriers of both CTAs initialised before any remote arrive / multicast
  if (warp == 2) tmem_alloc_pair(smem_u32(const_cast<uint32_t*>(tmem_slot)), Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int m_pairs = (M + 255) / 256;
  const int n_tiles = N / G2_BN;
  const int total = m_pairs * n_tiles;
  const int kblocks = K / GEMM_BK;
  const int cluster_id = blockIdx.x >> 1;
  const int n_clusters = gridDim.x >> 1;
  long long* const clk = (TL && blockIdx.x < 2) ? g_gemm2_clock : nullptr;
  int clk_n = 0;
#define G2_STAMP(role)                                                                       \
  do {                                                                                       \
    if constexpr (TL) {                                                                      \
      if (clk != nullptr && clk_n < 256) clk[(blockIdx.x * 2 + (role)) * 256 + clk_n++] = clock64(); \
    }                                                                                        \
  } while (0)

  if (warp == 0) {
    if (elect_one()) {
      const uint32_t lead_full0 = mapa_shared(full_bar, 0);
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = cluster_id; tile < total; tile += n_clusters) {
        const int m_pair = tile / n_tiles, n_blk = tile % n_tiles;
        const int row_a = m_pair * 256 + static_cast<int>(rank) * 128;
        const int row_b = n_blk * G2_BN + static_cast<int>(rank) * (G2_BN / 2);
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(empty_bar + 8u * stage, phase ^ 1u);
          G2_STAMP(0);
          const uint32_t a_dst = smem_base + stage * Cfg::STAGE_BYTES;
          if (flags & 4) {   // experiment: no loads, just hand the (stale) stage over
            if (leader) mbar_arrive(full_bar + 8u * stage);
          } else {
            // the peer's complete_tx may reach the leader's barrier before the leader's expect_tx:
            // the phase still cannot complete before that (single) arrival
            if (leader) mbar_expect_tx(full_bar + 8u * stage, 2 * Cfg::STAGE_BYTES);
            const uint32_t fb = lead_full0 + 8u * stage;
            tma_load_2d_pair(a_dst, &tm_a, fb, kb * GEMM_BK, row_a);
            tma_load_2d_pair(a_dst + Cfg::A_BYTES, &tm_b, fb, kb * GEMM_BK, row_b);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (leader && elect_one()) {
      constexpr uint32_t idesc = make_idesc_h16(256, G2_BN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int local = 0;
      for (int tile = cluster_id; tile < total; tile += n_clusters, ++local) {
        const int as = local & 1;
        const uint32_t aphase = (local >> 1) & 1u;
        mbar_wait_cluster(tempty_bar + 8u * as, aphase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(as * G2_BN);
        for (int kb = 0; kb < kblocks; ++kb) {
          G2_STAMP(1);
          // bytes of both CTAs have landed (plain CTA-scope wait, as CUTLASS' cluster transaction
          // barrier does: the data is consumed by the async proxy, not by this thread)
          mbar_wait(full_bar + 8u * stage, phase);
          G2_STAMP(1);
          tc_fence_after();
          const uint32_t a_addr = smem_base + stage * Cfg::STAGE_BYTES;
          const uint64_t a_desc = make_smem_desc_sw128(a_addr, 16, 1024);
          const uint64_t b_desc = make_smem_desc_sw128(a_addr + Cfg::A_BYTES, 16, 1024);
          if (!(flags & 2)) {
#pragma unroll
            for (int k = 0; k < GEMM_BK / 16; ++k)
              tc_mma_f16_ss_pair(d_tmem, a_desc + 2u * k, b_desc + 2u * k, idesc,
                                 static_cast<uint32_t>((kb | k) != 0));
          }
          tc_commit_pair(empty_bar + 8u * stage, 3);  // frees the stage in both CTAs
          G2_STAMP(1);
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        tc_commit_pair(tfull_bar + 8u * as, 3);
      }
    }
  } else if (warp >= 4) {
    const int q = warp & 3;
    const int half = (warp - 4) >> 2;
    constexpr int COLS_PER_WARP = G2_BN / 2;
    uint8_t* staging = smem + Cfg::STAGING_OFFSET + (warp - 4) * Cfg::STAGING_PER_WARP;
    const uint32_t staging_addr = smem_base + Cfg::STAGING_OFFSET + (warp - 4) * Cfg::STAGING_PER_WARP;
    const uint32_t lead_tempty0 = mapa_shared(tempty_bar, 0);
    const bool stamp = (warp == 4 && lane == 0 && !leader);   // the peer's issuer slot is free
    int local = 0;
    for (int tile = cluster_id; tile < total; tile += n_clusters, ++local) {
      const int m_pair = tile / n_tiles, n_blk = tile % n_tiles;
      const int as = local & 1;
      const uint32_t aphase = (local >> 1) & 1u;
      const int row0 = m_pair * 256 + static_cast<int>(rank) * 128 + q * 32;
      const int row = row0 + lane;
      const bool row_ok = row < M;
      const int col0 = half * COLS_PER_WARP;
      const int gcol0 = n_blk * G2_BN + col0;
      const h16* resid_row =
          (EPI == EPI_BIAS_RESID) ? resid + static_cast<size_t>(row) * N + gcol0 : nullptr;

      if (stamp) G2_STAMP(1);
      if (flags & 16) {   // experiment: poll with back-off instead of try_wait's own spin
        while (!mbar_test(tfull_bar + 8u * as, aphase)) __nanosleep(256);
      } else {
        mbar_wait(tfull_bar + 8u * as, aphase);
      }
      if (stamp) G2_STAMP(1);
      tc_fence_after();
      const uint32_t t_base = tmem_base + (static_cast<uint32_t>(q * 32) << 16) +
                              static_cast<uint32_t>(as * G2_BN + col0);
      // Two staging tiles per warp, used alternately: before a tile is overwritten only the TMA store
      // issued TWO stores ago must have finished reading it (bulk groups retire in order).
      if constexpr (epi_is_glu(EPI)) {
        uint32_t g[2][32], u[2][32];
        tmem_ld32(t_base, g[0]);
        tmem_ld32(t_base + 32u, g[1]);
        tmem_ld32(t_base + 64u, u[0]);
        tmem_ld32(t_base + 96u, u[1]);
        if (lane == 0) tma_store_wait_read<1>();
        __syncwarp();
        tmem_ld_wait();
        if (!(flags & 1)) {
          const int sidx = local & 1;
          gemm_swiglu_chunk<EPI == EPI_GEGLU>(g, u, staging + sidx * GEMM_STAGING_BYTES, lane);
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0 && row0 < M) {
            tma_store_2d(&tm_out, staging_addr + sidx * GEMM_STAGING_BYTES,
                         n_blk * (G2_BN / 2) + half * 64, row0);
            tma_store_commit();
          }
        }
      } else {
#pragma unroll 1
        for (int c = 0; c < COLS_PER_WARP / 64; ++c) {
          uint32_t acc[2][32];
          tmem_ld32(t_base + static_cast<uint32_t>(c * 64), acc[0]);
          tmem_ld32(t_base + static_cast<uint32_t>(c * 64 + 32), acc[1]);
          if (lane == 0) tma_store_wait_read<1>();
          __syncwarp();
          tmem_ld_wait();
          if (flags & 1) continue;
          gemm_epilogue_chunk_gbias<EPI>(acc, bias != nullptr ? bias + gcol0 + c * 64 : nullptr,
                                         resid_row + c * 64, row_ok, staging + c * GEMM_STAGING_BYTES,
                                         lane);
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0 && row0 < M) {
            tma_store_2d(&tm_out, staging_addr + c * GEMM_STAGING_BYTES, gcol0 + c * 64, row0);
            tma_store_commit();
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        // the accumulator stage is recycled by the leader's MMA warp: tell ITS barrier
        if (leader) mbar_arrive(tempty_bar + 8u * as);
        else mbar_arrive_cluster_relaxed(lead_tempty0 + 8u * as);
      }
      if (stamp) G2_STAMP(1);
    }
    if (lane == 0) tma_store_wait_all();
  }
#undef G2_STAMP

  tc_fence_before();
  cluster_sync_all();  // neither CTA may free TMEM / exit while the pair still references it
  if (warp == 2) tmem_dealloc_pair(tmem_base, Cfg::TMEM_COLS);
}

}  // namespace b2e
