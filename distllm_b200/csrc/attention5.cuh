// Streaming bidirectional self-attention for head_dim 64, fourth generation: the same item / ring / MMA structure
// as attention3.cuh, with FOUR softmax warpgroups per CTA instead of two.
//
// Why (profiles/r02_ncu_att3_source.md, r02_notes.md section 2): with two softmax warps per SM sub-partition the
// chunk body needs ~800 issue slots and 768 MUFU clocks per pair of chunks, but every warp also spends ~650 clk
// per chunk in latency-bound phases (wait for S, tcgen05.ld, tcgen05.st + wait, arrive, loop) during which only
// ONE other warp can use the sub-partition: the measured period is ~1600 clk.  Four warps per sub-partition give
// the scheduler three other streams to fill those gaps.
//
// How: the two warpgroups of a query tile split every 64-key chunk by KEY columns -- warpgroup "half" h owns keys
// [32 h, 32 h + 32) of each chunk -- and run two fully independent online softmaxes (own running maximum, own row
// sum, own accumulator O_h = sum_j P_j[:, half h] V_j[half h, :]); the two partial results of a row are merged
// once per tile, in the epilogue, exactly like split-K flash decoding:
//     M = max(m_0, m_1),  O = (O_0 2^(m_0 - M) + O_1 2^(m_1 - M)) / (l_0 2^(m_0 - M) + l_1 2^(m_1 - M)).
// No per-chunk communication between the halves; a thread keeps 32 scores instead of 64 (96 registers).
//
//   warps 0-3 / 4-7     softmax of query tile A (slot 0), keys [0,32) / [32,64) of every chunk
//   warps 8-11 / 12-15  softmax of query tile B (slot 1), keys [0,32) / [32,64)
//   warp 16, 18         MMA issuers of slot 0 / 1 (one elected thread each)
//   warp 17             loader (TMA)
//
// TMEM per slot (256 columns): S/P buffer 0 [0,64), S/P buffer 1 [64,128), O_0 [128,192), O_1 [192,256).
// Half h reads S columns [32 h, 32 h + 32) and writes its 16-bit P over the first 16 of those columns, so that
// neither half ever touches the other's scores.
#pragma once

#include "attention3.cuh"

namespace b2e {

constexpr int AT5_THREADS = 608;       // 16 softmax warps + 3 single-thread roles
constexpr int AT5_THREADS_EPI = 736;   // ... + 4 epilogue warps (V bit 7)

// Shared-memory layout: as attention3.cuh; with the epilogue role the K/V ring gives up one stage for the
// (m, l) exchange between the softmax warpgroups and the epilogue warpgroup.
template <bool EPI>
struct At5Smem {
  static constexpr int NST = EPI ? 7 : AT3_NST;
  static constexpr int BIAS = AT3_SMEM_KV + NST * 2 * AT3_KVTILE;
  static constexpr int OST = (BIAS + NST * AT3_KC * 4 + 1023) / 1024 * 1024;
  static constexpr int XCHG = OST + 2 * AT3_QTILE;            // [2 slot][2 half][128] float2 (EPI only)
  static constexpr int BAR = XCHG + (EPI ? 4096 : 0);
  static constexpr int BYTES = BAR + 512;
  static_assert(OST % 1024 == 0 && BYTES <= 232448, "shared memory layout");
};
static_assert(At5Smem<false>::BAR == AT3_SMEM_BAR && At5Smem<false>::BYTES == AT3_SMEM_BYTES, "same as attention3");

// V: bit 0 plain chunks from plain_chunks[b]; bits 2-3 exponentials per four on the FMA pipe (0, 1, 2);
//    bit 4 bidirectional sliding window.  (Same meaning as attention3_d64_kernel's.)
//    bit 7 a fifth warpgroup (warps 19-22) takes the per-tile epilogue -- merge of the halves, normalisation,
//    staging, store -- off the softmax warpgroups, which hand it (m, l) through shared memory and move on.
//    bit 8 timeline stamps compiled in (profiling builds only).
template <int V>
__global__ void __launch_bounds__((V & 128) ? AT5_THREADS_EPI : AT5_THREADS, 1)
attention5_d64_kernel(const __grid_constant__ CUtensorMap tm_q,   // [T, 3H] h16, box 64 x 128
                      const __grid_constant__ CUtensorMap tm_kv,  // [T, 3H] h16, box 64 x 64
                      const float* __restrict__ bias,             // [B, S_pad]
                      const int* __restrict__ kv_chunks,          // [B]
                      const int* __restrict__ plain_chunks,       // [B] or nullptr
                      const __grid_constant__ CUtensorMap tm_ctx, // [T, H] h16, box 64 x 128 (full tiles)
                      int B, int S, int S_pad, int heads, float scale_log2e, int window,
                      const int* __restrict__ seq_cu, const int* __restrict__ seq_len,
                      h16* __restrict__ ctx_out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sb = smem_u32(smem);
  if ((sb & 1023u) != 0) __trap();
  const int warp = threadIdx.x >> 5;
  const int H = heads * AT3_D;
  const int nq = (S + 127) / 128;
  const int npairs = (nq + 1) / 2;
  const int n_items = B * heads * npairs;
  constexpr bool kEpiRole = (V & 128) != 0;
  using Lay = At5Smem<kEpiRole>;
  constexpr int kNst = Lay::NST;

  // barriers (8 B each) in the 512 bytes behind the staging tiles (and the exchange)
  const uint32_t bar0 = sb + Lay::BAR;
  const uint32_t kv_full = bar0;                       // [NST]
  const uint32_t kv_empty = kv_full + 8 * kNst;        // [NST]
  const uint32_t q_full = kv_empty + 8 * kNst;         // [2 buf][2 slot]
  const uint32_t q_empty = q_full + 32;                // [2][2]
  const uint32_t s_ready = q_empty + 32;               // [2 slot][2 sbuf]
  const uint32_t p_ready = s_ready + 32;               // [2 slot][2 sbuf][2 half]
  const uint32_t pv_done = p_ready + 64;               // [2][2][2]  P_j V_j of that half has completed
  const uint32_t o_ready = pv_done + 64;               // [2 slot]
  const uint32_t o_empty = o_ready + 16;               // [2 slot]
  const uint32_t stats_ready = o_empty + 16;           // [2 slot]  (m, l) of all 256 rows-halves are in the exchange
  const uint32_t xchg_empty = stats_ready + 16;        // [2 slot]  the epilogue role has read them
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + Lay::BAR + 448);
  // (m, l) of every row and half are exchanged once per tile through the slot's output staging tile (2 KiB of its
  // 16 KiB; the staging rows are written only after every thread has read the exchange).
  if (warp == 16) {
    if (elect_one()) {
      tma_prefetch_desc(&tm_q);
      tma_prefetch_desc(&tm_kv);
      tma_prefetch_desc(&tm_ctx);
      for (int i = 0; i < kNst; ++i) {
        mbar_init(kv_full + 8u * i, 1);
        mbar_init(kv_empty + 8u * i, 2);   // one arrival from each slot's MMA issuer
      }
      for (int i = 0; i < 4; ++i) {
        mbar_init(q_full + 8u * i, 1);
        mbar_init(q_empty + 8u * i, 1);
        mbar_init(s_ready + 8u * i, 1);
      }
      for (int i = 0; i < 8; ++i) {
        mbar_init(p_ready + 8u * i, 128);
        mbar_init(pv_done + 8u * i, 1);
      }
      for (int i = 0; i < 2; ++i) {
        mbar_init(o_ready + 8u * i, 1);
        mbar_init(o_empty + 8u * i, kEpiRole ? 128 : 256);
        mbar_init(stats_ready + 8u * i, 256);
        mbar_init(xchg_empty + 8u * i, 128);
      }
      mbar_fence_init();
    }
    __syncwarp();
    tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_slot)), 512);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  long long* const clk = (blockIdx.x == 0) ? g_att3_clock : nullptr;
  int clk_n = 0;
  // timeline stamps exist only in the instantiations with V bit 8 (tools/att3_timeline.py): four predicated stamp
  // sites per chunk were 2-3 % of the softmax warps' issue slots (ISETP / CS2R samples in r02_ncu_att5_source.md)
  constexpr bool kStamp = (V & 256) != 0;
#define AT5_STAMP(role, code)                                          \
  do {                                                                 \
    if constexpr (kStamp) {                                            \
      if (clk != nullptr && clk_n < 256) {                             \
        clk[(role) * 512 + clk_n] = clock64();                         \
        clk[(role) * 512 + 256 + clk_n] = (code);                      \
        ++clk_n;                                                       \
      }                                                                \
    }                                                                  \
  } while (0)

  // No setmaxnreg here: the whole kernel compiles to the 96 registers that 608 threads leave per thread, so there
  // is nothing to hand from the single-thread roles to the softmax warps (and a setmaxnreg.inc that the register
  // file cannot satisfy never returns).
  if (warp >= 16) {
    if (warp == 17) {
      if (elect_one()) {
        // ------------------------------------------------------------ loader (as attention3)
        uint32_t chunk_ctr = 0;
        uint32_t q_par = 0, q_any = 0;
        int it = 0;
        int item = blockIdx.x;
        At3Walk wk;
        wk.init(item, gridDim.x, npairs, heads);
        At3Item cur = at3_finish(at3_fetch(wk, kv_chunks, plain_chunks, n_items, S, seq_cu, seq_len), window);
        for (; item < n_items; item += gridDim.x, ++it) {
          wk.step(npairs, heads);
          const At3Raw nxt = at3_fetch(wk, kv_chunks, plain_chunks, n_items, S, seq_cu, seq_len);
          const int pr = cur.pr, h = cur.h, b = cur.b;
          const int row_base = cur.row0;
          const int buf = it & 1;
          for (int slot = 0; slot < 2; ++slot) {
            const int t = 2 * pr + slot;
            if (t >= cur.nq) break;
            const int idx = buf * 2 + slot;
            const uint32_t bit = 1u << idx;
            if (q_any & bit) mbar_wait(q_empty + 8u * idx, ((q_par >> idx) & 1u) ^ 1u);
            const uint32_t qb = q_full + 8u * idx;
            mbar_expect_tx(qb, AT3_QTILE);
            tma_load_2d(sb + AT3_SMEM_Q + idx * AT3_QTILE, &tm_q, qb, h * AT3_D, row_base + t * 128);
            q_par ^= bit;
            q_any |= bit;
          }
          const int n = cur.n;
          for (int j = 0; j < n; ++j, ++chunk_ctr) {
            const int st = chunk_ctr % kNst;
            const uint32_t use = chunk_ctr / kNst;
            if (use > 0) mbar_wait(kv_empty + 8u * st, (use - 1) & 1u);
            const uint32_t fb = kv_full + 8u * st;
            mbar_expect_tx(fb, 2 * AT3_KVTILE + AT3_KC * 4);
            const uint32_t dst = sb + AT3_SMEM_KV + st * 2 * AT3_KVTILE;
            const int jk = (cur.j0 + j) * AT3_KC;
            tma_load_2d(dst, &tm_kv, fb, H + h * AT3_D, row_base + jk);
            tma_load_2d(dst + AT3_KVTILE, &tm_kv, fb, 2 * H + h * AT3_D, row_base + jk);
            bulk_load_1d(sb + Lay::BIAS + st * AT3_KC * 4,
                         bias + static_cast<size_t>(b) * S_pad + jk, AT3_KC * 4, fb);
          }
          cur = at3_finish(nxt, window);
        }
      }
    } else if (warp == 16 || warp == 18) {
      if (elect_one()) {
        // ------------------------------------------------------------ MMA issuer of ONE slot
        const int slot = (warp == 16) ? 0 : 1;
        constexpr uint32_t idesc_s = make_idesc_h16(128, AT3_KC, 0, 0);
        constexpr uint32_t idesc_o = make_idesc_h16(128, AT3_D, 0, 1);  // B (= V) is MN-major
        const uint32_t t_slot = tmem_base + static_cast<uint32_t>(slot * 256);
        uint32_t chunk_base = 0;
        uint32_t q_cnt[2] = {0, 0};
        uint32_t p_par = 0;           // bit (sbuf * 2 + half): parity of the p_ready phase to wait for
        uint32_t tile_cnt = 0;
        int it = 0;
        int item = blockIdx.x;
        At3Walk wk;
        wk.init(item, gridDim.x, npairs, heads);
        At3Item cur = at3_finish(at3_fetch(wk, kv_chunks, plain_chunks, n_items, S, seq_cu, seq_len), window);
        for (; item < n_items; item += gridDim.x, ++it) {
          wk.step(npairs, heads);
          const At3Raw nxt = at3_fetch(wk, kv_chunks, plain_chunks, n_items, S, seq_cu, seq_len);
          const int n = cur.n;
          const int buf = it & 1;
          const bool active = 2 * cur.pr + slot < cur.nq;
          if (slot == 0) AT5_STAMP(2, 9000 + n);
          if (!active) {
            for (int j = 0; j < n; ++j) {
              const uint32_t c = chunk_base + j;
              mbar_wait(kv_full + 8u * (c % kNst), (c / kNst) & 1u);
              mbar_arrive(kv_empty + 8u * (c % kNst));
            }
          } else {
            const int qidx = buf * 2 + slot;
            const uint64_t q_desc = make_smem_desc_sw128(sb + AT3_SMEM_Q + qidx * AT3_QTILE, 16, 1024);
            auto issue_qk = [&](int j) {
              const uint32_t c = chunk_base + j;
              const int st = c % kNst;
              mbar_wait(kv_full + 8u * st, (c / kNst) & 1u);
              tc_fence_after();
              const uint64_t k_desc =
                  make_smem_desc_sw128(sb + AT3_SMEM_KV + st * 2 * AT3_KVTILE, 16, 1024);
              const uint32_t d = t_slot + static_cast<uint32_t>((j & 1) * 64);
#pragma unroll
              for (int k = 0; k < AT3_D / 16; ++k)
                tc_mma_f16_ss(d, q_desc + 2u * k, k_desc + 2u * k, idesc_s, static_cast<uint32_t>(k != 0));
              tc_commit(s_ready + 8u * (slot * 2 + (j & 1)));
              if (slot == 0) AT5_STAMP(2, j * 10 + 1);
              if (j + 1 == n) tc_commit(q_empty + 8u * qidx);
            };
            mbar_wait(q_full + 8u * qidx, q_cnt[buf] & 1u);
            ++q_cnt[buf];
            issue_qk(0);
            for (int j = 0; j < n; ++j) {
              if (j + 1 < n) issue_qk(j + 1);
              const int sbuf = j & 1;
              const uint32_t c = chunk_base + j;
              const int st = c % kNst;
              const uint32_t v_base = sb + AT3_SMEM_KV + st * 2 * AT3_KVTILE + AT3_KVTILE;
#pragma unroll
              for (int half = 0; half < 2; ++half) {
                const int pb = sbuf * 2 + half;
                mbar_wait(p_ready + 8u * (slot * 4 + pb), (p_par >> pb) & 1u);
                p_par ^= 1u << pb;
                // the previous tile's epilogue (o_empty) precedes this tile's first p_ready
                if (j == 0 && half == 0 && tile_cnt > 0) mbar_wait(o_empty + 8u * slot, (tile_cnt - 1) & 1u);
                tc_fence_after();
                // P of this half: 16 packed columns at the start of its own 32 score columns
                const uint32_t p = t_slot + static_cast<uint32_t>(sbuf * 64 + half * 32);
                const uint32_t o = t_slot + 128u + static_cast<uint32_t>(half * 64);
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                  const uint64_t v_desc = make_smem_desc_sw128(v_base + (half * 2 + k) * 16 * 128, 1024, 1024);
                  tc_mma_f16_ts(o, p + static_cast<uint32_t>(8 * k), v_desc, idesc_o,
                                static_cast<uint32_t>((j | k) != 0));
                }
                tc_commit(pv_done + 8u * (slot * 4 + pb));
              }
              tc_commit(kv_empty + 8u * st);   // this slot is done with the stage (K by Q K^T, V by both P V)
              if (slot == 0) AT5_STAMP(2, j * 10 + 2);
              if (j + 1 == n) {
                tc_commit(o_ready + 8u * slot);
                ++tile_cnt;
              }
            }
          }
          chunk_base += static_cast<uint32_t>(n);
          cur = at3_finish(nxt, window);
        }
      }
    }
    if (kEpiRole && warp >= 19) {
      // ---------------------------------------------------------------- epilogue warpgroup (V bit 7)
      // thread r owns row r of whichever tile is finishing: merges the two halves (m, l from the exchange, O_0 and
      // O_1 from TMEM), normalises, stages 128 bytes, and the tile leaves through TMA (or row by row when partial)
      const int r = (warp & 3) * 32 + (threadIdx.x & 31);
      const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
      uint32_t cnt[2] = {0, 0};   // tiles finished per slot (parity of stats_ready / o_ready)
      int item = blockIdx.x;
      At3Walk wk;
      wk.init(item, gridDim.x, npairs, heads);
      At3Item cur = at3_finish(at3_fetch(wk, kv_chunks, plain_chunks, n_items, S, seq_cu, seq_len), window);
      for (; item < n_items; item += gridDim.x) {
        wk.step(npairs, heads);
        const At3Raw nxt = at3_fetch(wk, kv_chunks, plain_chunks, n_items, S, seq_cu, seq_len);
        for (int slot = 0; slot < 2; ++slot) {
          const int t = 2 * cur.pr + slot;
          if (t >= cur.nq) break;
          const uint32_t t_o = tmem_base + lane_base + static_cast<uint32_t>(slot * 256) + 128u;
          uint8_t* ostage = smem + Lay::OST + slot * AT3_QTILE;
          const uint32_t ostage_addr = sb + Lay::OST + slot * AT3_QTILE;
          const float2* xchg = reinterpret_cast<const float2*>(smem + Lay::XCHG) + slot * 256;
          mbar_wait(stats_ready + 8u * slot, cnt[slot] & 1u);
          const float2 h0 = xchg[r], h1 = xchg[128 + r];   // (m, l) of half 0 / 1; m = -inf when l == 0
          const float m_all = fmaxf(h0.x, h1.x);
          const float w0 = (h0.y > 0.0f) ? fast_exp2(h0.x - m_all) : 0.0f;
          const float w1 = (h1.y > 0.0f) ? fast_exp2(h1.x - m_all) : 0.0f;
          const float inv_l = 1.0f / (h0.y * w0 + h1.y * w1);
          const float c0 = w0 * inv_l, c1 = w1 * inv_l;
          mbar_arrive(xchg_empty + 8u * slot);   // the softmax warpgroups may post the next tile's numbers
          if (r == 0) tma_store_wait_read<0>();  // the slot's previous store has read the staging tile
          mbar_wait(o_ready + 8u * slot, cnt[slot] & 1u);
          ++cnt[slot];
          tc_fence_after();
          asm volatile("bar.sync %0, 128;" ::"r"(2) : "memory");
#pragma unroll 1
          for (int cc = 0; cc < 2; ++cc) {
            uint32_t o0[32], o1[32];
            tmem_ld32(t_o + static_cast<uint32_t>(cc * 32), o0);
            tmem_ld32(t_o + static_cast<uint32_t>(64 + cc * 32), o1);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; i += 8) {
              float f[8];
#pragma unroll
              for (int e = 0; e < 8; ++e)
                f[e] = fmaf(__uint_as_float(o0[i + e]), c0, __uint_as_float(o1[i + e]) * c1);
              uint4 w;
              w.x = pack_h16x2(f[0], f[1]);
              w.y = pack_h16x2(f[2], f[3]);
              w.z = pack_h16x2(f[4], f[5]);
              w.w = pack_h16x2(f[6], f[7]);
              const int unit = cc * 4 + (i >> 3);
              *reinterpret_cast<uint4*>(ostage + r * 128 + ((unit ^ (r & 7)) << 4)) = w;
            }
          }
          tc_fence_before();
          mbar_arrive(o_empty + 8u * slot);   // O's TMEM columns may be overwritten by the next tile
          fence_proxy_async_smem();
          asm volatile("bar.sync %0, 128;" ::"r"(2) : "memory");
          const int valid = cur.len - t * 128;
          if (valid >= 128) {
            if (r == 0) {
              tma_store_2d(&tm_ctx, ostage_addr, cur.h * AT3_D, cur.row0 + t * 128);
              tma_store_commit();
            }
          } else if (r < valid) {
            h16* dst = ctx_out + static_cast<size_t>(cur.row0 + t * 128 + r) * H + cur.h * AT3_D;
#pragma unroll
            for (int u = 0; u < 8; ++u)
              *reinterpret_cast<uint4*>(dst + u * 8) =
                  *reinterpret_cast<const uint4*>(ostage + r * 128 + ((u ^ (r & 7)) << 4));
          }
        }
        cur = at3_finish(nxt, window);
      }
      if (r == 0) tma_store_wait_all();
    }
  } else {
    // -------------------------------------------------------------- softmax warpgroups
    const int slot = warp >> 3;
    const int half = (warp >> 2) & 1;
    const int r = threadIdx.x & 127;
    const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const uint32_t t_slot = tmem_base + lane_base + static_cast<uint32_t>(slot * 256);
    const uint32_t t_o = t_slot + 128u;   // O_0 at +0, O_1 at +64
    uint32_t chunk_base = 0;
    uint32_t s_par = 0;   // bit sbuf: parity of the s_ready[slot][sbuf] phase to wait for
    uint32_t o_cnt = 0;
    int item = blockIdx.x;
    At3Walk wk;
        wk.init(item, gridDim.x, npairs, heads);
        At3Item cur = at3_finish(at3_fetch(wk, kv_chunks, plain_chunks, n_items, S, seq_cu, seq_len), window);
    uint8_t* ostage = smem + Lay::OST + slot * AT3_QTILE;
    const uint32_t ostage_addr = sb + Lay::OST + slot * AT3_QTILE;
    // the (m, l) exchange of the slot: two float2 per row in the slot's staging tile (free until the merge)
    float2* xchg = reinterpret_cast<float2*>(ostage);
    const int stamp_role = slot;   // half 0 of each slot records the timeline
    for (; item < n_items; item += gridDim.x) {
      wk.step(npairs, heads);
          const At3Raw nxt = at3_fetch(wk, kv_chunks, plain_chunks, n_items, S, seq_cu, seq_len);
      const int pr = cur.pr, h = cur.h, n = cur.n;
      const int t = 2 * pr + slot;
      if (t < cur.nq) {
        constexpr bool kPlainCount = (V & 1) != 0;
        constexpr int kPoly = ((V >> 2) & 3) > 2 ? 2 : ((V >> 2) & 3);
        const int n_plain = cur.np;
        float m_used = 0.0f, l = 0.0f;
        uint32_t s[32];
        for (int j = 0; j < n; ++j) {
          const int sbuf = j & 1;
          const uint32_t c = chunk_base + j;
          const int st = c % kNst;
          if (r == 0 && half == 0) AT5_STAMP(stamp_role, j * 10 + 0);
          mbar_wait(s_ready + 8u * (slot * 2 + sbuf), (s_par >> sbuf) & 1u);
          s_par ^= 1u << sbuf;
          tc_fence_after();
          const uint32_t t_s = t_slot + static_cast<uint32_t>(sbuf * 64 + half * 32);
          tmem_ld32(t_s, s);
          tmem_ld_wait();
          if (r == 0 && half == 0) AT5_STAMP(stamp_role, j * 10 + 1);
          const float* bias_j =
              reinterpret_cast<const float*>(smem + Lay::BIAS + st * AT3_KC * 4) + half * 32;
          bool plain;
          if (kPlainCount) {
            plain = j < n_plain;
            if (!plain) mbar_wait(kv_full + 8u * st, (c / kNst) & 1u);  // complete: acquires the bias bytes
          } else {
            mbar_wait(kv_full + 8u * st, (c / kNst) & 1u);
            const float bz = bias_j[threadIdx.x & 31];
            plain = !__any_sync(0xffffffffu, bz != 0.0f);
          }
          uint32_t pk[16];
          if ((V & 16) != 0) {
            // sliding window: key (j0 + j) * 64 + 32 half + i is visible to query row q iff |q - key| <= window
            const int q_abs = t * 128 + r;
            const int ilo = q_abs - window - (cur.j0 + j) * AT3_KC - half * 32;
            const unsigned span = static_cast<unsigned>(2 * window);
            constexpr uint32_t kOut = 0xfcf0bdc2u;   // -1e37f
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (static_cast<unsigned>(i - ilo) > span) s[i] = kOut;
          }
          // ONE copy of each chunk body (the first chunk differs only in where its maximum comes from): the
          // instruction footprint matters, every item boundary walks code that was not executed for 8 chunks
          bool done = false;
          if (plain) {
            if (j == 0) m_used = scale_log2e * at3_smax_plain(s, -INFINITY);   // exact maximum: every p <= 1
            const float sum = at3_exp_pack_plain<kPoly>(s, scale_log2e, -m_used, pk);
            // every p <= row sum: a sum within 2^threshold proves that no score ran away
            const bool calm = (j == 0) || sum <= 256.0f;
            if (__all_sync(0xffffffffu, calm)) {
              l += sum;
              done = true;
            } else if (kPlainCount) {
              mbar_wait(kv_full + 8u * st, (c / kNst) & 1u);   // the general path reads the bias row
            }
          }
          if (!done) {
            if (j == 0) {
              // first chunk of this half: exact maximum first.  A half whose 32 keys all lie beyond the padded
              // sequence (bias -inf: only when S <= 32) starts from 0 and contributes l = 0, O = 0.
              float cmax = at3_max(s, bias_j, scale_log2e, -INFINITY);
              if (cmax == -INFINITY) cmax = 0.0f;
              m_used = cmax;
            }
            float xmax = -INFINITY;
            float sum = at3_exp_pack(s, bias_j, scale_log2e, m_used, pk, xmax);
            const bool need = (j > 0) && xmax > m_used + AT3_RESCALE_THRESHOLD;
            if (__any_sync(0xffffffffu, need)) {
              const float m_new = need ? xmax : m_used;
              const float sc = fast_exp2(m_used - m_new);  // 1 for rows that keep their maximum
              m_used = m_new;
              l *= sc;
              float dummy = -INFINITY;
              sum = at3_exp_pack(s, bias_j, scale_log2e, m_used, pk, dummy);
              // O_half = sum_{i<j} P_i V_i must be complete before it is rescaled (see attention3.cuh)
              const int pb = sbuf ^ 1;
              mbar_wait(pv_done + 8u * (slot * 4 + pb * 2 + half), ((s_par >> pb) & 1u) ^ 1u);
              tc_fence_after();
#pragma unroll 1
              for (int cc = 0; cc < 2; ++cc) {
                uint32_t o[32];
                const uint32_t t_oh = t_o + static_cast<uint32_t>(half * 64 + cc * 32);
                tmem_ld32(t_oh, o);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * sc);
                tmem_st32(t_oh, o);
              }
            }
            l += sum;
          }
          if (r == 0 && half == 0) AT5_STAMP(stamp_role, j * 10 + 2);
          tmem_st16(t_s, pk);   // 16-bit P over the first 16 of this half's own score columns
          tmem_st_wait();
          tc_fence_before();
          mbar_arrive(p_ready + 8u * (slot * 4 + sbuf * 2 + half));
          if (r == 0 && half == 0) AT5_STAMP(stamp_role, j * 10 + 3);
        }
        if constexpr (kEpiRole) {
          // ---- hand (m, l) to the epilogue warpgroup and move on (a half that saw no key with non-zero weight
          // takes no part in the maximum); the exchange is free once that warpgroup has read the previous tile's
          if (r == 0 && half == 0) AT5_STAMP(stamp_role, 900);
          if (o_cnt > 0) mbar_wait(xchg_empty + 8u * slot, (o_cnt - 1) & 1u);
          ++o_cnt;
          reinterpret_cast<float2*>(smem + Lay::XCHG)[slot * 256 + half * 128 + r] =
              make_float2(l > 0.0f ? m_used : -INFINITY, l);
          mbar_arrive(stats_ready + 8u * slot);
          if (r == 0 && half == 0) AT5_STAMP(stamp_role, 902);
        } else {
        // ---- epilogue: merge the two halves of every row, O -> h16 -> swizzled staging tile -> TMA store
        if (r == 0 && half == 0) AT5_STAMP(stamp_role, 900);
        if (r == 0 && half == 0) tma_store_wait_read<0>();   // the previous tile's store has read the staging
        asm volatile("bar.sync %0, 256;" ::"r"(2 + slot) : "memory");
        // a half that saw no key with non-zero weight takes no part in the maximum
        xchg[half * 128 + r] = make_float2(l > 0.0f ? m_used : -INFINITY, l);
        asm volatile("bar.sync %0, 256;" ::"r"(2 + slot) : "memory");
        const float2 other = xchg[(half ^ 1) * 128 + r];
        const float m_mine = l > 0.0f ? m_used : -INFINITY;
        const float m_all = fmaxf(m_mine, other.x);
        const float c_mine = (l > 0.0f) ? fast_exp2(m_mine - m_all) : 0.0f;
        const float c_other = (other.y > 0.0f) ? fast_exp2(other.x - m_all) : 0.0f;
        const float inv_l = 1.0f / (l * c_mine + other.y * c_other);
        const float c0 = (half == 0 ? c_mine : c_other) * inv_l;   // weight of O_0
        const float c1 = (half == 0 ? c_other : c_mine) * inv_l;   // weight of O_1
        mbar_wait(o_ready + 8u * slot, o_cnt & 1u);
        if (r == 0 && half == 0) AT5_STAMP(stamp_role, 901);
        ++o_cnt;
        tc_fence_after();
        // everybody has read the exchange before the staging tile is overwritten
        asm volatile("bar.sync %0, 256;" ::"r"(2 + slot) : "memory");
        {
          // this thread's 32 output columns [32 half, 32 half + 32) of row r
          uint32_t o0[32], o1[32];
          tmem_ld32(t_o + static_cast<uint32_t>(half * 32), o0);
          tmem_ld32(t_o + static_cast<uint32_t>(64 + half * 32), o1);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; i += 8) {
            float f[8];
#pragma unroll
            for (int e = 0; e < 8; ++e)
              f[e] = fmaf(__uint_as_float(o0[i + e]), c0, __uint_as_float(o1[i + e]) * c1);
            uint4 w;
            w.x = pack_h16x2(f[0], f[1]);
            w.y = pack_h16x2(f[2], f[3]);
            w.z = pack_h16x2(f[4], f[5]);
            w.w = pack_h16x2(f[6], f[7]);
            const int unit = half * 4 + (i >> 3);
            *reinterpret_cast<uint4*>(ostage + r * 128 + ((unit ^ (r & 7)) << 4)) = w;
          }
        }
        tc_fence_before();
        mbar_arrive(o_empty + 8u * slot);   // O's TMEM columns may be overwritten by the next tile
        fence_proxy_async_smem();
        asm volatile("bar.sync %0, 256;" ::"r"(2 + slot) : "memory");
        const int valid = cur.len - t * 128;   // rows of this tile that belong to the sequence
        if (valid >= 128) {
          if (r == 0 && half == 0) {
            tma_store_2d(&tm_ctx, ostage_addr, h * AT3_D, cur.row0 + t * 128);
            tma_store_commit();
            AT5_STAMP(stamp_role, 902);
          }
        } else if (r < valid) {
          // last, partial tile of the sequence: every thread stores the half of its row that it staged
          h16* dst = ctx_out + static_cast<size_t>(cur.row0 + t * 128 + r) * H + h * AT3_D + half * 32;
#pragma unroll
          for (int u = 0; u < 4; ++u)
            *reinterpret_cast<uint4*>(dst + u * 8) =
                *reinterpret_cast<const uint4*>(ostage + r * 128 + (((half * 4 + u) ^ (r & 7)) << 4));
        }
        }   // !kEpiRole
      }
      chunk_base += static_cast<uint32_t>(n);
      cur = at3_finish(nxt, window);
    }
    if (!kEpiRole && r == 0 && half == 0) tma_store_wait_all();
  }

#undef AT5_STAMP
  tc_fence_before();
  __syncthreads();
  if (warp == 16) tmem_dealloc(tmem_base, 512);
}

}  // namespace b2e
