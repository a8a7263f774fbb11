// Streaming CAUSAL self-attention for head_dim 128 with grouped-query K/V and an optional sliding
// window, on sm_100a (the Mistral-family attention; same machinery as attention3.cuh).
//
// Persistent CTAs (one per SM) walk work items (sequence b, query head h, pair of 128-row query
// tiles), heaviest pairs first.  A query tile t only visits the 64-key chunks that can hold a visible
// key: from (128 t - window + 1) / 64 up to the diagonal (2 t + 1), clipped to the chunks that hold
// an attended key.  The two tiles of a pair share one K/V ring; chunks that only the lower tile needs
// are released by the upper tile's slot without being touched.
//
//   warp 9      loader      Q tiles and {K_j, V_j, key-bias_j} ring stages (TMA)
//   warps 8,10  MMA issuers (one thread per query tile): S_j = Q K_j^T (SS, 128x64x16, 8 K-steps
//                           over two 64-wide slabs) and O += P_j V_j (TS: P from TMEM, V_j as MN-major
//                           smem operand, N = 128)
//   warps 0-3   softmax for query tile A (slot 0)   one row per thread, online softmax in the exp2
//   warps 4-7   softmax for query tile B (slot 1)   domain with lazy rescale; h16 P over S's columns
//
// TMEM: slot s at column 256*s: S/P buffer 0 [0,64), S/P buffer 1 [64,128), O [128,256).
// Masking: a key j is visible to query i iff j <= i, (window == 0 or i - j < window) and the key is
// attended (transformers/models/mistral/modeling_mistral.py:122-180, masking_utils sliding-window
// causal mask).  Invisible scores take the same finite "most negative" value as padded keys, so a row
// whose first visited chunk holds no visible key carries a harmless running maximum that the lazy
// rescale (factor exp2(-3e38) = 0) wipes as soon as a visible key shows up.  Rows that never see a key
// (queries inside left padding) produce finite garbage; HF's SDPA gives zeros there.  Such rows are
// never read: not as keys (masked), not by the poolers.
#pragma once

#include "attention3.cuh"

namespace b2e {

constexpr int AT4_D = 128;
constexpr int AT4_KC = 64;
constexpr int AT4_THREADS = 384;
constexpr int AT4_NST = 4;                              // K/V ring stages
constexpr int AT4_QSLAB = 128 * 64 * 2;                 // 16 KiB: 128 rows x 64 columns
constexpr int AT4_QTILE = 2 * AT4_QSLAB;                // 32 KiB
constexpr int AT4_KVSLAB = AT4_KC * 64 * 2;             // 8 KiB: 64 keys x 64 columns
constexpr int AT4_KVTILE = 2 * AT4_KVSLAB;              // 16 KiB (K or V of one chunk)
constexpr int AT4_SMEM_Q = 0;                           // [2 slots]
constexpr int AT4_SMEM_KV = AT4_SMEM_Q + 2 * AT4_QTILE;               // stage: K | V
constexpr int AT4_SMEM_OST = AT4_SMEM_KV + AT4_NST * 2 * AT4_KVTILE;  // [2 slots] 128 x 64 staging
constexpr int AT4_SMEM_BIAS = AT4_SMEM_OST + 2 * AT4_QSLAB;
constexpr int AT4_SMEM_BAR = AT4_SMEM_BIAS + AT4_NST * AT4_KC * 4;
constexpr int AT4_SMEM_BYTES = AT4_SMEM_BAR + 512;
static_assert(AT4_SMEM_KV % 1024 == 0 && AT4_SMEM_OST % 1024 == 0, "swizzled tiles: 1024-byte aligned");
static_assert(AT4_SMEM_BYTES <= 232448, "exceeds the 227 KiB per-CTA shared memory limit");

// x = scale*s + bias, or the masked value when the column is outside [klo, klo + span]
template <bool EDGE>
__device__ __forceinline__ float at4_x(uint32_t s, float scale, float b, int col, int klo,
                                       unsigned span) {
  const float x = fmaf(__uint_as_float(s), scale, b);
  if (EDGE) return (static_cast<unsigned>(col - klo) <= span) ? x : AT3_MASKED;
  return x;
}
template <bool EDGE>
__device__ __forceinline__ float at4_max(const uint32_t (&s)[32], const float* __restrict__ bias,
                                         float scale, float m, int col0, int klo, unsigned span) {
#pragma unroll
  for (int i = 0; i < 32; i += 4) {
    const float4 bz = *reinterpret_cast<const float4*>(bias + i);
    m = fmaxf(m, at4_x<EDGE>(s[i + 0], scale, bz.x, col0 + i + 0, klo, span));
    m = fmaxf(m, at4_x<EDGE>(s[i + 1], scale, bz.y, col0 + i + 1, klo, span));
    m = fmaxf(m, at4_x<EDGE>(s[i + 2], scale, bz.z, col0 + i + 2, klo, span));
    m = fmaxf(m, at4_x<EDGE>(s[i + 3], scale, bz.w, col0 + i + 3, klo, span));
  }
  return m;
}
template <bool EDGE>
__device__ __forceinline__ float at4_exp_pack(const uint32_t (&s)[32], const float* __restrict__ bias,
                                              float scale, float m, uint32_t* pk, float& xmax,
                                              int col0, int klo, unsigned span) {
  float sum = 0.0f;
#pragma unroll
  for (int i = 0; i < 32; i += 4) {
    const float4 bz = *reinterpret_cast<const float4*>(bias + i);
    const float x0 = at4_x<EDGE>(s[i + 0], scale, bz.x, col0 + i + 0, klo, span);
    const float x1 = at4_x<EDGE>(s[i + 1], scale, bz.y, col0 + i + 1, klo, span);
    const float x2 = at4_x<EDGE>(s[i + 2], scale, bz.z, col0 + i + 2, klo, span);
    const float x3 = at4_x<EDGE>(s[i + 3], scale, bz.w, col0 + i + 3, klo, span);
    xmax = fmaxf(fmaxf(xmax, fmaxf(x0, x1)), fmaxf(x2, x3));
    const float p0 = fast_exp2(x0 - m), p1 = fast_exp2(x1 - m);
    const float p2 = fast_exp2(x2 - m), p3 = fast_exp2(x3 - m);
    sum += (p0 + p1) + (p2 + p3);
    pk[i / 2] = pack_h16x2(p0, p1);
    pk[i / 2 + 1] = pack_h16x2(p2, p3);
  }
  return sum;
}

// One work item = (sequence b, query head h, pair of query tiles pr).  Chunk ranges are absolute
// 64-key chunk indices: slot s visits [lo[s], hi[s]); the ring carries [lo[0], hi_u).
struct At4Item {
  int b, h, pr, lo[2], hi[2], hi_u;
  int np;   // leading key chunks of the sequence whose 64 keys are all attended
  int row0, len;   // token layout (pack.cuh): rows [row0, row0 + len) hold the sequence; b * S and S when padded
};
// Two steps, like attention3.cuh's at3_fetch / at3_finish: the next item's per-sequence numbers are loaded at
// the top of the current item and first touched at its end (warps issue in order).
struct At4Raw {
  int b, h, pr, kvc, np, ok, row0, len;
};
__device__ __forceinline__ At4Raw at4_fetch(int item, int npairs, int heads, const int* __restrict__ kv_chunks,
                                            int n_items, int bh_total, int S, const int* __restrict__ seq_cu,
                                            const int* __restrict__ seq_len,
                                            const int* __restrict__ plain_chunks = nullptr) {
  At4Raw r;
  r.ok = item < n_items;
  const int it = r.ok ? item : 0;
  r.pr = npairs - 1 - it / bh_total;   // heaviest (latest) query tiles first
  const int bh = it % bh_total;
  r.h = bh % heads;
  r.b = bh / heads;
  r.kvc = __ldg(kv_chunks + r.b);
  r.np = plain_chunks != nullptr ? __ldg(plain_chunks + r.b) : 0;
  r.row0 = seq_cu != nullptr ? __ldg(seq_cu + r.b) : r.b * S;
  r.len = seq_len != nullptr ? __ldg(seq_len + r.b) : S;
  return r;
}
__device__ __forceinline__ At4Item at4_finish(At4Raw r, int window) {
  asm volatile("" : "+r"(r.kvc), "+r"(r.np), "+r"(r.row0), "+r"(r.len));
  At4Item it{0, 0, 0, {0, 0}, {0, 0}, 0, 0, 0, 0};
  if (r.ok) {
    it.pr = r.pr;
    it.h = r.h;
    it.b = r.b;
    it.np = r.np;
    it.row0 = r.row0;
    it.len = r.len;
    const int nq = (r.len + 127) / 128;   // query tiles of THIS sequence (tiles behind it belong to the next one)
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      const int t = 2 * it.pr + s;
      int hi = min(r.kvc, 2 * t + 2);
      int lo = (window > 0) ? max(0, 128 * t - window + 1) / AT4_KC : 0;
      if (lo >= hi) lo = hi - 1;
      if (t >= nq) { lo = 0; hi = 0; }
      it.lo[s] = lo;
      it.hi[s] = hi;
    }
    it.hi_u = max(it.hi[0], it.hi[1]);
  }
  return it;
}

__global__ void __launch_bounds__(AT4_THREADS, 1)
attention4_d128_causal_kernel(const __grid_constant__ CUtensorMap tm_q,   // [T, ld] h16, box 64 x 128
                              const __grid_constant__ CUtensorMap tm_kv,  // [T, ld] h16, box 64 x 64
                              const float* __restrict__ bias,             // [B, S_pad]
                              const int* __restrict__ kv_chunks,          // [B]
                              const int* __restrict__ plain_chunks,       // [B] or nullptr
                              const __grid_constant__ CUtensorMap tm_ctx, // [T, heads*128], box 64 x 128 (full tiles)
                              int B, int S, int S_pad, int heads, int kv_heads, int window,
                              float scale_log2e,
                              const int* __restrict__ seq_cu,    // [B] first row of each sequence, or nullptr (= b*S)
                              const int* __restrict__ seq_len,   // [B] rows of each sequence, or nullptr (= S)
                              h16* __restrict__ ctx_out) {       // [T, heads*128]: partial last tiles, row by row
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sb = smem_u32(smem);
  if ((sb & 1023u) != 0) __trap();
  const int warp = threadIdx.x >> 5;
  const int nq = (S + 127) / 128;
  const int npairs = (nq + 1) / 2;
  const int bh_total = B * heads;
  const int n_items = bh_total * npairs;
  const int k_col0 = heads * AT4_D;                 // K columns start after the query heads
  const int v_col0 = (heads + kv_heads) * AT4_D;
  const int group = heads / kv_heads;

  const uint32_t bar0 = sb + AT4_SMEM_BAR;
  const uint32_t kv_full = bar0;                       // [NST]
  const uint32_t kv_empty = kv_full + 8 * AT4_NST;     // [NST]
  const uint32_t q_full = kv_empty + 8 * AT4_NST;      // [2 slot]
  const uint32_t q_empty = q_full + 16;                // [2 slot]
  const uint32_t s_ready = q_empty + 16;               // [2 slot][2 sbuf]
  const uint32_t p_ready = s_ready + 32;               // [2][2]
  const uint32_t pv_done = p_ready + 32;               // [2][2]
  const uint32_t o_ready = pv_done + 32;               // [2 slot]
  const uint32_t o_empty = o_ready + 16;               // [2 slot]
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + AT4_SMEM_BAR + 384);

  if (warp == 8) {
    if (elect_one()) {
      tma_prefetch_desc(&tm_q);
      tma_prefetch_desc(&tm_kv);
      tma_prefetch_desc(&tm_ctx);
      for (int i = 0; i < AT4_NST; ++i) {
        mbar_init(kv_full + 8u * i, 1);
        mbar_init(kv_empty + 8u * i, 2);   // one arrival from each slot's MMA issuer
      }
      for (int i = 0; i < 4; ++i) {
        mbar_init(s_ready + 8u * i, 1);
        mbar_init(p_ready + 8u * i, 128);
        mbar_init(pv_done + 8u * i, 1);
      }
      for (int i = 0; i < 2; ++i) {
        mbar_init(q_full + 8u * i, 1);
        mbar_init(q_empty + 8u * i, 1);
        mbar_init(o_ready + 8u * i, 1);
        mbar_init(o_empty + 8u * i, 128);
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

  if (warp >= 8) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
    if (warp == 9) {
      if (elect_one()) {
        // ------------------------------------------------------------ loader
        uint32_t chunk_ctr = 0;       // ring position, runs across items
        uint32_t q_loads[2] = {0, 0}; // Q tiles loaded so far per slot
        int item = blockIdx.x;
        At4Item cur = at4_finish(at4_fetch(item, npairs, heads, kv_chunks, n_items, bh_total, S, seq_cu, seq_len), window);
        for (; item < n_items; item += gridDim.x) {
          const At4Raw nxt = at4_fetch(item + gridDim.x, npairs, heads, kv_chunks, n_items, bh_total, S, seq_cu, seq_len);
          const int row_base = cur.row0;
          const int hk = cur.h / group;
          const int n_active = (cur.hi[0] > 0 ? 1 : 0) + (cur.hi[1] > 0 ? 1 : 0);   // slot 1 implies slot 0
          int q_pending = n_active;   // slots [n_active - q_pending, n_active) still need their Q tile
          // Q tile of a slot: its buffer is free once the last Q K^T of the previous tile completed.
          auto try_q = [&](bool block) {
#pragma unroll
            for (int slot = 0; slot < 2; ++slot) {
              if (q_pending == 0 || slot != n_active - q_pending) continue;   // served in order
              if (q_loads[slot] > 0) {
                const uint32_t par = (q_loads[slot] - 1) & 1u;
                if (block) mbar_wait(q_empty + 8u * slot, par);
                else if (!mbar_test(q_empty + 8u * slot, par)) return;
              }
              const uint32_t qb = q_full + 8u * slot;
              const uint32_t dst = sb + AT4_SMEM_Q + slot * AT4_QTILE;
              const int row = row_base + (2 * cur.pr + slot) * 128;
              mbar_expect_tx(qb, AT4_QTILE);
              tma_load_2d(dst, &tm_q, qb, cur.h * AT4_D, row);
              tma_load_2d(dst + AT4_QSLAB, &tm_q, qb, cur.h * AT4_D + 64, row);
              ++q_loads[slot];
              --q_pending;
            }
          };
          for (int j = cur.lo[0]; j < cur.hi_u; ++j, ++chunk_ctr) {
            try_q(false);
            const int st = chunk_ctr % AT4_NST;
            const uint32_t use = chunk_ctr / AT4_NST;
            if (use > 0 && !mbar_test(kv_empty + 8u * st, (use - 1) & 1u)) {
              try_q(true);   // never sleep on the ring while a Q tile of this item is still owed
              mbar_wait(kv_empty + 8u * st, (use - 1) & 1u);
            }
            const uint32_t fb = kv_full + 8u * st;
            mbar_expect_tx(fb, 2 * AT4_KVTILE + AT4_KC * 4);
            const uint32_t dst = sb + AT4_SMEM_KV + st * 2 * AT4_KVTILE;
            const int krow = row_base + j * AT4_KC;
            tma_load_2d(dst, &tm_kv, fb, k_col0 + hk * AT4_D, krow);
            tma_load_2d(dst + AT4_KVSLAB, &tm_kv, fb, k_col0 + hk * AT4_D + 64, krow);
            tma_load_2d(dst + AT4_KVTILE, &tm_kv, fb, v_col0 + hk * AT4_D, krow);
            tma_load_2d(dst + AT4_KVTILE + AT4_KVSLAB, &tm_kv, fb, v_col0 + hk * AT4_D + 64, krow);
            bulk_load_1d(sb + AT4_SMEM_BIAS + st * AT4_KC * 4,
                         bias + static_cast<size_t>(cur.b) * S_pad + j * AT4_KC, AT4_KC * 4, fb);
          }
          try_q(true);
          cur = at4_finish(nxt, window);
        }
      }
    } else if (warp == 8 || warp == 10) {
      if (elect_one()) {
        // ------------------------------------------------------------ MMA issuer of ONE slot
        // (warp 8: query tile A, warp 10: query tile B).  With head_dim 128 one thread driving both
        // slots spent ~600 clk per Q K^T or P V event (polls, 8 or 4 MMAs, commits), four events per pair
        // of chunks -- more than the softmax it feeds.  Each slot's thread walks the item's ring chunks in
        // order: chunks outside its tile's range only get their ring arrival, the others Q K^T / P V.
        const int slot = (warp == 8) ? 0 : 1;
        constexpr uint32_t idesc_s = make_idesc_h16(128, AT4_KC, 0, 0);
        constexpr uint32_t idesc_o = make_idesc_h16(128, AT4_D, 0, 1);  // B (= V) is MN-major
        const uint32_t t_slot = tmem_base + static_cast<uint32_t>(slot * 256);
        const uint32_t q_addr = sb + AT4_SMEM_Q + slot * AT4_QTILE;
        uint32_t chunk_base = 0;   // ring position of this item's first chunk (lo[0])
        uint32_t q_cnt = 0;        // Q tiles consumed (parity of q_full)
        uint32_t p_par = 0;        // bit sbuf: parity of the p_ready phase to wait for
        uint32_t tile_cnt = 0;
        int item = blockIdx.x;
        At4Item cur = at4_finish(at4_fetch(item, npairs, heads, kv_chunks, n_items, bh_total, S, seq_cu, seq_len), window);
        for (; item < n_items; item += gridDim.x) {
          const At4Raw nxt = at4_fetch(item + gridDim.x, npairs, heads, kv_chunks, n_items, bh_total, S, seq_cu, seq_len);
          const int n_u = cur.hi_u - cur.lo[0];
          const int my_lo = slot ? cur.lo[1] : cur.lo[0];
          const int my_hi = slot ? cur.hi[1] : cur.hi[0];
          const bool active = my_hi > 0;
          const int c_lo = active ? my_lo - cur.lo[0] : n_u;   // positions relative to the ring's first chunk
          const int c_hi = active ? my_hi - cur.lo[0] : n_u;
          auto pass_on = [&](int c) {   // a chunk this slot does not use: arrive once it has been filled
            const uint32_t rc = chunk_base + c;
            mbar_wait(kv_full + 8u * (rc % AT4_NST), (rc / AT4_NST) & 1u);
            mbar_arrive(kv_empty + 8u * (rc % AT4_NST));
          };
          auto issue_qk = [&](int c) {
            const int sbuf = (c - c_lo) & 1;
            const uint32_t rc = chunk_base + c;
            const int st = rc % AT4_NST;
            mbar_wait(kv_full + 8u * st, (rc / AT4_NST) & 1u);
            tc_fence_after();
            const uint32_t k_addr = sb + AT4_SMEM_KV + st * 2 * AT4_KVTILE;
            const uint32_t d = t_slot + static_cast<uint32_t>(sbuf * 64);
#pragma unroll
            for (int k = 0; k < AT4_D / 16; ++k) {
              const uint64_t q_desc =
                  make_smem_desc_sw128(q_addr + (k >> 2) * AT4_QSLAB, 16, 1024) + 2u * (k & 3);
              const uint64_t k_desc =
                  make_smem_desc_sw128(k_addr + (k >> 2) * AT4_KVSLAB, 16, 1024) + 2u * (k & 3);
              tc_mma_f16_ss(d, q_desc, k_desc, idesc_s, static_cast<uint32_t>(k != 0));
            }
            tc_commit(s_ready + 8u * (slot * 2 + sbuf));
            if (c + 1 == c_hi) tc_commit(q_empty + 8u * slot);
          };
          for (int c = 0; c < c_lo; ++c) pass_on(c);
          if (active) {
            mbar_wait(q_full + 8u * slot, q_cnt & 1u);
            ++q_cnt;
            issue_qk(c_lo);
            for (int c = c_lo; c < c_hi; ++c) {
              // S of the next chunk goes out before P of this one is awaited (its buffer held the P of
              // chunk c-1, whose P V was issued one iteration ago)
              if (c + 1 < c_hi) issue_qk(c + 1);
              const int sbuf = (c - c_lo) & 1;
              mbar_wait(p_ready + 8u * (slot * 2 + sbuf), (p_par >> sbuf) & 1u);
              p_par ^= 1u << sbuf;
              // the previous tile's epilogue (o_empty) precedes this tile's first p_ready
              if (c == c_lo && tile_cnt > 0) mbar_wait(o_empty + 8u * slot, (tile_cnt - 1) & 1u);
              tc_fence_after();
              const uint32_t rc = chunk_base + c;
              const int st = rc % AT4_NST;
              const uint32_t p = t_slot + static_cast<uint32_t>(sbuf * 64);
              const uint32_t o = t_slot + 128u;
              const uint32_t v_base = sb + AT4_SMEM_KV + st * 2 * AT4_KVTILE + AT4_KVTILE;
#pragma unroll
              for (int k = 0; k < AT4_KC / 16; ++k) {
                // 16 keys = two 8-key groups 1024 B apart (SBO); the two 64-wide d slabs are 8 KiB
                // apart (LBO)
                const uint64_t v_desc = make_smem_desc_sw128(v_base + k * 16 * 128, AT4_KVSLAB, 1024);
                tc_mma_f16_ts(o, p + static_cast<uint32_t>(8 * k), v_desc, idesc_o,
                              static_cast<uint32_t>((c != c_lo) || k != 0));
              }
              tc_commit(pv_done + 8u * (slot * 2 + sbuf));
              tc_commit(kv_empty + 8u * st);   // this slot is done with the stage
              if (c + 1 == c_hi) {
                tc_commit(o_ready + 8u * slot);
                ++tile_cnt;
              }
            }
          }
          for (int c = c_hi; c < n_u; ++c) pass_on(c);
          chunk_base += static_cast<uint32_t>(n_u);
          cur = at4_finish(nxt, window);
        }
      }
    }
  } else {
    // -------------------------------------------------------------- softmax warpgroups
    asm volatile("setmaxnreg.inc.sync.aligned.u32 200;");
    const int slot = warp >> 2;
    const int r = threadIdx.x & 127;
    const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const uint32_t t_slot = tmem_base + lane_base + static_cast<uint32_t>(slot * 256);
    const uint32_t t_o = t_slot + 128u;
    uint32_t chunk_base = 0;
    uint32_t s_par = 0;   // bit sbuf: parity of the s_ready[slot][sbuf] phase to wait for
    uint32_t o_cnt = 0;
    int item = blockIdx.x;
    At4Item cur = at4_finish(at4_fetch(item, npairs, heads, kv_chunks, n_items, bh_total, S, seq_cu, seq_len, plain_chunks), window);
    uint8_t* ostage = smem + AT4_SMEM_OST + slot * AT4_QSLAB;
    const uint32_t ostage_addr = sb + AT4_SMEM_OST + slot * AT4_QSLAB;
    for (; item < n_items; item += gridDim.x) {
      const At4Raw nxt = at4_fetch(item + gridDim.x, npairs, heads, kv_chunks, n_items, bh_total, S, seq_cu, seq_len, plain_chunks);
      const int t = 2 * cur.pr + slot;
      const int lo = slot ? cur.lo[1] : cur.lo[0];
      const int hi = slot ? cur.hi[1] : cur.hi[0];
      if (hi > 0) {
        const int qi = t * 128 + r;   // this thread's query position inside the sequence
        float m_used = 0.0f, l = 0.0f;
        for (int j = lo; j < hi; ++j) {
          const int sbuf = (j - lo) & 1;
          const uint32_t c = chunk_base + static_cast<uint32_t>(j - cur.lo[0]);
          const int st = c % AT4_NST;
          mbar_wait(s_ready + 8u * (slot * 2 + sbuf), (s_par >> sbuf) & 1u);
          s_par ^= 1u << sbuf;
          // visible key columns of this row inside the chunk: [klo, klo + span]
          const int key0 = j * AT4_KC;
          const bool edge = (key0 + AT4_KC - 1 > t * 128) ||
                            (window > 0 && t * 128 + 127 - key0 >= window);
          // interior chunk of fully attended keys: neither the bias row nor a per-element visibility test is
          // needed (the same fast path as attention3.cuh: 3.5 issue slots per element, one exponential in four
          // on the FMA pipe, the general path only when a score runs away from the running maximum)
          const bool plain = !edge && j < cur.np;
          if (!plain) mbar_wait(kv_full + 8u * st, (c / AT4_NST) & 1u);  // complete: acquires the bias bytes
          tc_fence_after();
          const float* bias_j = reinterpret_cast<const float*>(smem + AT4_SMEM_BIAS + st * AT4_KC * 4);
          const uint32_t t_s = t_slot + static_cast<uint32_t>(sbuf * 64);
          uint32_t s0[32], s1[32];
          tmem_ld32(t_s, s0);
          tmem_ld32(t_s + 32u, s1);
          tmem_ld_wait();
          int klo = (window > 0) ? max(0, qi - window + 1 - key0) : 0;
          const int khi = min(AT4_KC - 1, qi - key0);
          unsigned span = static_cast<unsigned>(khi - klo);
          if (khi < klo) { klo = AT4_KC; span = 0u; }   // nothing visible: every compare fails
          uint32_t pk[32];
          float xmax = -INFINITY;
          float sum = 0.0f;
          const bool first = (j == lo);
          bool done = false;
          if (plain) {
            if (first) {
              m_used = scale_log2e * at3_smax_plain(s1, at3_smax_plain(s0, -INFINITY));
              sum = at3_exp_pack_plain<1>(s0, scale_log2e, -m_used, pk);
              sum += at3_exp_pack_plain<1>(s1, scale_log2e, -m_used, pk + 16);
              done = true;
            } else {
              sum = at3_exp_pack_plain<1>(s0, scale_log2e, -m_used, pk);
              sum += at3_exp_pack_plain<1>(s1, scale_log2e, -m_used, pk + 16);
              // every p <= row sum: a sum within 2^threshold proves that no score ran away
              done = __all_sync(0xffffffffu, sum <= 256.0f);
              if (!done) mbar_wait(kv_full + 8u * st, (c / AT4_NST) & 1u);   // the general path reads the bias row
            }
          }
          if (!done) {
          if (first) {
            // first chunk of the row: exact maximum first (finite: masked scores are finite too)
            float cmax;
            if (edge) {
              cmax = at4_max<true>(s0, bias_j, scale_log2e, -INFINITY, 0, klo, span);
              cmax = at4_max<true>(s1, bias_j + 32, scale_log2e, cmax, 32, klo, span);
            } else {
              cmax = at4_max<false>(s0, bias_j, scale_log2e, -INFINITY, 0, 0, 0u);
              cmax = at4_max<false>(s1, bias_j + 32, scale_log2e, cmax, 32, 0, 0u);
            }
            m_used = cmax;
          }
          if (edge) {
            sum = at4_exp_pack<true>(s0, bias_j, scale_log2e, m_used, pk, xmax, 0, klo, span);
            sum += at4_exp_pack<true>(s1, bias_j + 32, scale_log2e, m_used, pk + 16, xmax, 32, klo, span);
          } else {
            sum = at4_exp_pack<false>(s0, bias_j, scale_log2e, m_used, pk, xmax, 0, 0, 0u);
            sum += at4_exp_pack<false>(s1, bias_j + 32, scale_log2e, m_used, pk + 16, xmax, 32, 0, 0u);
          }
          if (!first) {
            // redo only if this chunk exceeds the running maximum by more than 2^8
            const bool need = xmax > m_used + AT3_RESCALE_THRESHOLD;
            if (__any_sync(0xffffffffu, need)) {
              const float m_new = need ? xmax : m_used;
              const float sc = fast_exp2(m_used - m_new);  // 1 for rows that keep their maximum
              m_used = m_new;
              l *= sc;
              float dummy = -INFINITY;
              if (edge) {
                sum = at4_exp_pack<true>(s0, bias_j, scale_log2e, m_used, pk, dummy, 0, klo, span);
                sum += at4_exp_pack<true>(s1, bias_j + 32, scale_log2e, m_used, pk + 16, dummy, 32, klo, span);
              } else {
                sum = at4_exp_pack<false>(s0, bias_j, scale_log2e, m_used, pk, dummy, 0, 0, 0u);
                sum += at4_exp_pack<false>(s1, bias_j + 32, scale_log2e, m_used, pk + 16, dummy, 32, 0, 0u);
              }
              // O = sum over earlier chunks must be complete before it is rescaled (attention3.cuh)
              const int pb = sbuf ^ 1;
              mbar_wait(pv_done + 8u * (slot * 2 + pb), ((s_par >> pb) & 1u) ^ 1u);
              tc_fence_after();
#pragma unroll 1
              for (int cc = 0; cc < 4; ++cc) {
                uint32_t o[32];
                tmem_ld32(t_o + static_cast<uint32_t>(cc * 32), o);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * sc);
                tmem_st32(t_o + static_cast<uint32_t>(cc * 32), o);
              }
            }
          }
          }
          l += sum;
          tmem_st32(t_s, pk);  // h16 P over the first 32 columns of S's own buffer
          tmem_st_wait();
          tc_fence_before();
          mbar_arrive(p_ready + 8u * (slot * 2 + sbuf));
        }
        // ---- epilogue: O / l -> h16 -> swizzled 128 x 64 staging tile -> TMA store, twice
        mbar_wait(o_ready + 8u * slot, o_cnt & 1u);
        ++o_cnt;
        tc_fence_after();
        const float inv_l = 1.0f / l;
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
          if (r == 0) tma_store_wait_read<0>();   // the previous store has read the staging tile
          asm volatile("bar.sync %0, 128;" ::"r"(2 + slot) : "memory");
#pragma unroll 1
          for (int cc = 0; cc < 2; ++cc) {
            uint32_t o[32];
            tmem_ld32(t_o + static_cast<uint32_t>(half * 64 + cc * 32), o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; i += 8) {
              uint4 w;
              w.x = pack_h16x2(__uint_as_float(o[i + 0]) * inv_l, __uint_as_float(o[i + 1]) * inv_l);
              w.y = pack_h16x2(__uint_as_float(o[i + 2]) * inv_l, __uint_as_float(o[i + 3]) * inv_l);
              w.z = pack_h16x2(__uint_as_float(o[i + 4]) * inv_l, __uint_as_float(o[i + 5]) * inv_l);
              w.w = pack_h16x2(__uint_as_float(o[i + 6]) * inv_l, __uint_as_float(o[i + 7]) * inv_l);
              const int unit = cc * 4 + (i >> 3);
              *reinterpret_cast<uint4*>(ostage + r * 128 + ((unit ^ (r & 7)) << 4)) = w;
            }
          }
          if (half == 1) {
            tc_fence_before();
            mbar_arrive(o_empty + 8u * slot);   // O's TMEM columns may be overwritten by the next tile
          }
          fence_proxy_async_smem();
          asm volatile("bar.sync %0, 128;" ::"r"(2 + slot) : "memory");
          const int valid = cur.len - t * 128;   // rows of this tile that belong to the sequence
          if (valid >= 128) {
            if (r == 0) {
              tma_store_2d(&tm_ctx, ostage_addr, cur.h * AT4_D + half * 64, cur.row0 + t * 128);
              tma_store_commit();
            }
          } else if (r < valid) {
            // last, partial tile of the sequence: the rows behind it belong to the NEXT sequence or do not exist;
            // every thread stores the row it staged itself
            h16* dst = ctx_out + static_cast<size_t>(cur.row0 + t * 128 + r) * (heads * AT4_D) + cur.h * AT4_D +
                       half * 64;
#pragma unroll
            for (int u = 0; u < 8; ++u)
              *reinterpret_cast<uint4*>(dst + u * 8) =
                  *reinterpret_cast<const uint4*>(ostage + r * 128 + ((u ^ (r & 7)) << 4));
          }
        }
      }
      chunk_base += static_cast<uint32_t>(cur.hi_u - cur.lo[0]);
      cur = at4_finish(nxt, window);
    }
    if (r == 0) tma_store_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem_base, 512);
}

}  // namespace b2e
