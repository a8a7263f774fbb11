// Streaming bidirectional self-attention for head_dim 64 on sm_100a (third generation, any S).
//
// Persistent CTAs (one per SM) walk work items (sequence b, head h, pair of 128-row query tiles).
// K/V of the head stream through a shared-memory ring in 64-key chunks, so the TMA loader runs
// ahead across items (no load latency at item boundaries) and S is not limited by shared memory.
//
//   warp 9      loader      Q tiles (double-buffered across items) and {K_j, V_j, mask-bias_j}
//                           ring stages: cp.async.bulk.tensor + one 256 B cp.async.bulk
//   warps 8,10  MMA issuers (one elected thread each, one per query tile / TMEM slot):
//                           S_j = Q K_j^T (SS, 128x64x16) and O += P_j V_j (TS: P read from TMEM,
//                           V_j as MN-major smem operand); S is DOUBLE-buffered per slot, so
//                           Q K_{j+1}^T is issued before softmax_j finishes and the MMA round trip
//                           leaves the softmax critical path
//   warps 0-3   softmax for query tile A (slot 0)    one row per thread; scores of a chunk live in
//   warps 4-7   softmax for query tile B (slot 1)    registers; online softmax with lazy rescale;
//                                                    h16 P written over S's own TMEM columns
//
// TMEM: slot s at column 256*s: S/P buffer 0 [0,64), S/P buffer 1 [64,128), O [128,192).
// Semantics: HF SDPA with an additive key-padding mask (transformers/models/bert/
// modeling_bert.py:192-205, :692-716; the same call pattern serves ESM's attention).  The bias row
// (0 attended / most-negative-finite padded / -inf beyond S) and the number of key chunks that hold
// an attended key are prepared once per forward pass by attn_prep_kernel.
#pragma once

#include "common.cuh"

namespace b2e {

constexpr int AT3_D = 64;
constexpr int AT3_KC = 64;                            // keys per chunk
constexpr int AT3_THREADS = 384;
constexpr int AT3_NST = 8;                            // K/V ring stages
constexpr int AT3_QTILE = 128 * AT3_D * 2;            // 16 KiB
constexpr int AT3_KVTILE = AT3_KC * AT3_D * 2;        // 8 KiB
constexpr int AT3_SMEM_Q = 0;                         // [2 item buffers][2 slots]
constexpr int AT3_SMEM_KV = AT3_SMEM_Q + 4 * AT3_QTILE;            // stage: K | V
constexpr int AT3_SMEM_BIAS = AT3_SMEM_KV + AT3_NST * 2 * AT3_KVTILE;
constexpr int AT3_SMEM_OST = AT3_SMEM_BIAS + AT3_NST * AT3_KC * 4;   // [2 slots] output staging tiles
constexpr int AT3_SMEM_BAR = AT3_SMEM_OST + 2 * AT3_QTILE;
constexpr int AT3_SMEM_BYTES = AT3_SMEM_BAR + 512;
static_assert(AT3_SMEM_OST % 1024 == 0, "swizzled staging tiles need 1024-byte alignment");
static_assert(AT3_SMEM_BYTES <= 232448, "exceeds the 227 KiB per-CTA shared memory limit");

constexpr float AT3_MASKED = -3.0e38f;
constexpr float AT3_RESCALE_THRESHOLD = 8.0f;

// bias[b, j] for j < S_pad (multiple of 64) and kv_chunks[b] = chunks holding an attended key
// (all chunks when nothing is attended, so that such a row degenerates to HF's uniform softmax).
// plain_chunks[b] (nullable) = number of LEADING chunks whose 64 keys are all attended: on those the softmax
// needs neither the bias row nor a vote over it (right-padded batches: every chunk but the last one or two).
__global__ void attn_prep_kernel(const int64_t* __restrict__ mask, float* __restrict__ bias,
                                 int* __restrict__ kv_chunks, int* __restrict__ plain_chunks, int B, int S,
                                 int S_pad) {
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  const int lane = threadIdx.x & 31;
  int last = 0;
  int first_off = S_pad;   // first key position that is NOT attended (padding beyond S counts)
  for (int j = lane; j < S_pad; j += 32) {
    float v = -INFINITY;
    bool on = false;
    if (j < S) {
      on = mask[static_cast<size_t>(b) * S + j] != 0;
      v = on ? 0.0f : AT3_MASKED;
      if (on) last = j + 1;
    }
    if (!on) first_off = min(first_off, j);
    bias[static_cast<size_t>(b) * S_pad + j] = v;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    last = max(last, __shfl_xor_sync(0xffffffffu, last, o));
    first_off = min(first_off, __shfl_xor_sync(0xffffffffu, first_off, o));
  }
  if (lane == 0) {
    kv_chunks[b] = last > 0 ? (last + AT3_KC - 1) / AT3_KC : S_pad / AT3_KC;
    if (plain_chunks != nullptr) plain_chunks[b] = first_off / AT3_KC;
  }
}

// ---- softmax helpers over 32 register-resident scores (x = scale*s + bias formed on the fly)
__device__ __forceinline__ float at3_max(const uint32_t (&s)[32], const float* __restrict__ bias,
                                         float scale, float m) {
#pragma unroll
  for (int i = 0; i < 32; i += 4) {
    const float4 bz = *reinterpret_cast<const float4*>(bias + i);
    m = fmaxf(m, fmaf(__uint_as_float(s[i + 0]), scale, bz.x));
    m = fmaxf(m, fmaf(__uint_as_float(s[i + 1]), scale, bz.y));
    m = fmaxf(m, fmaf(__uint_as_float(s[i + 2]), scale, bz.z));
    m = fmaxf(m, fmaf(__uint_as_float(s[i + 3]), scale, bz.w));
  }
  return m;
}
// p = exp2(x - m) -> h16 pairs; returns the fp32 row-sum contribution and tracks max(x)
__device__ __forceinline__ float at3_exp_pack(const uint32_t (&s)[32], const float* __restrict__ bias,
                                              float scale, float m, uint32_t* pk, float& xmax) {
  float sum = 0.0f;
#pragma unroll
  for (int i = 0; i < 32; i += 4) {
    const float4 bz = *reinterpret_cast<const float4*>(bias + i);
    const float x0 = fmaf(__uint_as_float(s[i + 0]), scale, bz.x);
    const float x1 = fmaf(__uint_as_float(s[i + 1]), scale, bz.y);
    const float x2 = fmaf(__uint_as_float(s[i + 2]), scale, bz.z);
    const float x3 = fmaf(__uint_as_float(s[i + 3]), scale, bz.w);
    xmax = fmaxf(fmaxf(xmax, fmaxf(x0, x1)), fmaxf(x2, x3));
    const float p0 = fast_exp2(x0 - m), p1 = fast_exp2(x1 - m);
    const float p2 = fast_exp2(x2 - m), p3 = fast_exp2(x3 - m);
    sum += (p0 + p1) + (p2 + p3);
    pk[i / 2] = pack_h16x2(p0, p1);
    pk[i / 2 + 1] = pack_h16x2(p2, p3);
  }
  return sum;
}

// Variants for a chunk whose 64 keys are all attended (bias == 0): x = scale*s, so the subtraction of the
// running maximum folds into the FMA (p = exp2(fma(s, scale, -m))) and no per-element maximum is tracked:
// 3.5 issue slots per element instead of 5.75.  The caller falls back to the general path when the
// chunk's row sum shows that some score may exceed the running maximum by more than the threshold.
__device__ __forceinline__ float at3_smax_plain(const uint32_t (&s)[32], float m) {
#pragma unroll
  for (int i = 0; i < 32; ++i) m = fmaxf(m, __uint_as_float(s[i]));
  return m;
}
// 2^x on the FMA / ALU pipes instead of the MUFU pipe (16 ex2 per clock and SM is what bounds head_dim-64
// attention): Cody-Waite range reduction n = round(x), f = x - n in [-0.5, 0.5], 2^f by a degree-4 least-squares
// polynomial (max relative error 2.7e-6 on the interval, the level of ex2.approx itself and far below the
// 2^-11 rounding of P to half), n added to the exponent field as an integer.  x is clamped at -126 (no
// denormal / wrap-around for very negative scores).
__device__ __forceinline__ float poly_exp2(float x) {
  x = fmaxf(x, -126.0f);
  const float t = x + 12582912.0f;   // 1.5 * 2^23: round-to-nearest integer part in the low mantissa bits
  const float n = t - 12582912.0f;
  const float f = x - n;
  float p = 0.009560510f;
  p = fmaf(p, f, 0.055917039f);
  p = fmaf(p, f, 0.240249811f);
  p = fmaf(p, f, 0.693121968f);
  p = fmaf(p, f, 0.999999191f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}

// POLY = how many of every four exponentials run on the FMA pipe (0, 1 or 2)
template <int POLY>
__device__ __forceinline__ float at3_exp_pack_plain(const uint32_t (&s)[32], float scale, float neg_m,
                                                    uint32_t* pk) {
  float sum0 = 0.0f, sum1 = 0.0f;
#pragma unroll
  for (int i = 0; i < 32; i += 4) {
    const float x0 = fmaf(__uint_as_float(s[i + 0]), scale, neg_m);
    const float x1 = fmaf(__uint_as_float(s[i + 1]), scale, neg_m);
    const float x2 = fmaf(__uint_as_float(s[i + 2]), scale, neg_m);
    const float x3 = fmaf(__uint_as_float(s[i + 3]), scale, neg_m);
    const float p0 = fast_exp2(x0);
    const float p1 = POLY >= 2 ? poly_exp2(x1) : fast_exp2(x1);
    const float p2 = fast_exp2(x2);
    const float p3 = POLY >= 1 ? poly_exp2(x3) : fast_exp2(x3);
    sum0 += p0 + p1;
    sum1 += p2 + p3;
    pk[i / 2] = pack_h16x2(p0, p1);
    pk[i / 2 + 1] = pack_h16x2(p2, p3);
  }
  return sum0 + sum1;
}

// The same two exponentials of poly_exp2 as ONE stream of packed fp32x2 instructions.
__device__ __forceinline__ void poly_exp2_x2(uint64_t x, float& p_lo, float& p_hi) {
  float x0, x1;
  f32x2_split(x, x0, x1);
  x = f32x2(fmaxf(x0, -126.0f), fmaxf(x1, -126.0f));
  const uint64_t magic = f32x2(12582912.0f, 12582912.0f);
  const uint64_t t = add_f32x2(x, magic);
  const uint64_t f = sub_f32x2(x, sub_f32x2(t, magic));
  uint64_t p = f32x2(0.009560510f, 0.009560510f);
  p = fma_f32x2(p, f, f32x2(0.055917039f, 0.055917039f));
  p = fma_f32x2(p, f, f32x2(0.240249811f, 0.240249811f));
  p = fma_f32x2(p, f, f32x2(0.693121968f, 0.693121968f));
  p = fma_f32x2(p, f, f32x2(0.999999191f, 0.999999191f));
  float t0, t1, q0, q1;
  f32x2_split(t, t0, t1);
  f32x2_split(p, q0, q1);
  p_lo = __int_as_float(__float_as_int(q0) + (__float_as_int(t0) << 23));
  p_hi = __int_as_float(__float_as_int(q1) + (__float_as_int(t1) << 23));
}
// at3_exp_pack_plain with packed fp32x2 arithmetic: x = fma(s, scale, -m) and the row sum take one FMA-pipe
// instruction per PAIR of scores; POLY16 of every 16 exponentials (whole pairs: 0, 4, 6 or 8) run on the FMA pipe.
template <int POLY16>
__device__ __forceinline__ float at3_exp_pack_plain_x2(const uint32_t (&s)[32], float scale, float neg_m,
                                                       uint32_t* pk) {
  const uint64_t sc2 = f32x2(scale, scale), nm2 = f32x2(neg_m, neg_m);
  uint64_t sum_a = 0, sum_b = 0;   // two fp32 zeros each
#pragma unroll
  for (int i = 0; i < 32; i += 16) {
#pragma unroll
    for (int k = 0; k < 16; k += 2) {
      const uint64_t x =
          fma_f32x2(f32x2(__uint_as_float(s[i + k]), __uint_as_float(s[i + k + 1])), sc2, nm2);
      const bool poly = (POLY16 >= 2 && k == 14) || (POLY16 >= 4 && k == 6) || (POLY16 >= 6 && k == 10) ||
                        (POLY16 >= 8 && k == 2);
      float p0, p1;
      if (poly) {
        poly_exp2_x2(x, p0, p1);
      } else {
        float x0, x1;
        f32x2_split(x, x0, x1);
        p0 = fast_exp2(x0);
        p1 = fast_exp2(x1);
      }
      if (k & 2) sum_b = add_f32x2(sum_b, f32x2(p0, p1)); else sum_a = add_f32x2(sum_a, f32x2(p0, p1));
      pk[(i + k) / 2] = pack_h16x2(p0, p1);
    }
  }
  float a, b;
  f32x2_split(add_f32x2(sum_a, sum_b), a, b);
  return a + b;
}

// One work item = (sequence b, head h, pair of query tiles pr); n = key chunks to visit.
struct At3Item {
  int b, h, pr, n, np;   // n: key chunks to visit; np: leading chunks whose 64 keys are all attended
  int j0;                // first key chunk (0 unless a sliding window narrows the range)
  int row0, nq, len;     // first row of the sequence in the token arrays, its 128-row query tiles, its rows
};
// window > 0 (bidirectional sliding window, |q - k| <= window): the pair of query tiles [256 pr, 256 pr + 255]
// only needs the key chunks that overlap [256 pr - window, 256 pr + 255 + window]; BOTH tiles walk that same
// range (the band is applied per element), so every role sees the same chunk list.  At least one chunk is
// always visited: rows of padding tiles must still come out finite.
// Decoding an item is split in two so that the global loads of the NEXT item's per-sequence numbers are issued
// at the top of the current item and first touched at its end: warps issue in order, so arithmetic placed right
// behind the loads would stall every role for a full L2 / DRAM round trip once per item
// (profiles/r02_ncu_att3_source.md: `long_sb` on the per-item branches).
struct At3Raw {
  int b, h, pr;
  int n, np, row0, len;   // loaded: key chunks, plain chunks, first row, rows
  int ok;                 // item < n_items
};
// Position of a CTA's current work item, advanced by gridDim.x per step WITHOUT divisions: three integer divisions
// by run-time values are ~450 clk of dependent instructions for a single thread, and every role paid them between
// two items (the ~700-800 clk gap between an item's last event and the next item's first in
// profiles/r02_att3_timeline_*.log).
struct At3Walk {
  int item, pr, h, b;   // item = (b * heads + h) * npairs + pr
  int G, gr, gh, gb;    // the stride gridDim.x in the same mixed radix
  __device__ __forceinline__ void init(int first, int stride, int npairs, int heads) {
    item = first;
    pr = first % npairs;
    const int bh = first / npairs;
    h = bh % heads;
    b = bh / heads;
    G = stride;
    gr = stride % npairs;
    const int gq = stride / npairs;
    gh = gq % heads;
    gb = gq / heads;
  }
  __device__ __forceinline__ void step(int npairs, int heads) {
    item += G;
    pr += gr;
    int c = pr >= npairs ? 1 : 0;
    pr -= c ? npairs : 0;
    h += gh + c;
    c = h >= heads ? 1 : 0;
    h -= c ? heads : 0;
    b += gb + c;
  }
};
__device__ __forceinline__ At3Raw at3_fetch(const At3Walk& w, const int* __restrict__ kv_chunks,
                                            const int* __restrict__ plain_chunks, int n_items, int S,
                                            const int* __restrict__ seq_cu, const int* __restrict__ seq_len) {
  At3Raw r;
  r.ok = w.item < n_items;
  r.pr = r.ok ? w.pr : 0;
  r.h = r.ok ? w.h : 0;
  r.b = r.ok ? w.b : 0;   // always a valid index: the loads need no branch
  r.n = __ldg(kv_chunks + r.b);
  r.np = plain_chunks != nullptr ? __ldg(plain_chunks + r.b) : 0;
  // token layout (pack.cuh): rows [row0, row0 + len) hold the sequence; padded layout when seq_cu is null
  r.row0 = seq_cu != nullptr ? __ldg(seq_cu + r.b) : r.b * S;
  r.len = seq_len != nullptr ? __ldg(seq_len + r.b) : S;
  return r;
}
__device__ __forceinline__ At3Item at3_finish(At3Raw r, int window) {
  // the loaded values become visible to the arithmetic below only here
  asm volatile("" : "+r"(r.n), "+r"(r.np), "+r"(r.row0), "+r"(r.len));
  At3Item it{0, 0, 0, 0, 0, 0, 0, 0, 0};
  if (r.ok) {
    it.b = r.b;
    it.h = r.h;
    it.pr = r.pr;
    it.n = r.n;
    it.np = r.np;
    it.row0 = r.row0;
    it.len = r.len;
    it.nq = (it.len + 127) / 128;
    if (2 * it.pr >= it.nq) it.n = 0;   // both query tiles lie beyond the sequence: nothing to do
    if (window > 0 && it.n > 0) {
      int lo = (256 * it.pr - window) / AT3_KC;
      if (256 * it.pr - window < 0) lo = 0;
      int hi = (256 * it.pr + 255 + window) / AT3_KC;
      if (hi > it.n - 1) hi = it.n - 1;
      if (lo > it.n - 1) lo = it.n - 1;
      if (hi < lo) hi = lo;
      it.j0 = lo;
      it.n = hi - lo + 1;
      it.np = 0;   // every chunk needs the band test
    }
  }
  return it;
}

// profiling aid (b2e_debug_set_clock_buffer): CTA 0 records (clock64, code) pairs per role
__device__ long long* g_att3_clock = nullptr;
__device__ int g_att3_flags = 2;   // 0 free-running, 1 strict ping-pong of the exp phase, 2 (default) de-phase once per item

// V (softmax variant, measured side by side in tools/att_bench.py):
//   bit 0  plain chunks come from plain_chunks[b]: no wait for / vote over the bias row on them
//   bit 1  S_{j+1} is fetched from TMEM right behind the store of P_j, so that tcgen05.ld's latency runs
//          under the publish (st wait, fence, arrive) instead of in front of the next chunk's exponentials
//   bits 2-3  exponentials per four that run on the FMA pipe (plain chunks only): 0, 1 or 2
//   bit 5  packed fp32x2 arithmetic on plain chunks; bits 2-3 then mean 0, 4, 6 or 8 exponentials per 16 on
//          the FMA pipe
//   bit 8  timeline stamps compiled in (profiling instantiation, tools/att3_timeline.py)
//   bit 4  bidirectional sliding window (`window` > 0: ModernBERT's local layers): chunk range per item narrowed
//          to the band, scores outside |q - k| <= window masked per element
template <int V>
__global__ void __launch_bounds__(AT3_THREADS, 1)
attention3_d64_kernel(const __grid_constant__ CUtensorMap tm_q,   // [T, 3H] h16, box 64 x 128
                      const __grid_constant__ CUtensorMap tm_kv,  // [T, 3H] h16, box 64 x 64
                      const float* __restrict__ bias,             // [B, S_pad]
                      const int* __restrict__ kv_chunks,          // [B]
                      const int* __restrict__ plain_chunks,       // [B] or nullptr
                      const __grid_constant__ CUtensorMap tm_ctx, // [T, H] h16, box 64 x 128 (full tiles)
                      int B, int S, int S_pad, int heads, float scale_log2e, int window,
                      const int* __restrict__ seq_cu,    // [B] first row of each sequence, or nullptr (= b*S)
                      const int* __restrict__ seq_len,   // [B] rows of each sequence, or nullptr (= S)
                      h16* __restrict__ ctx_out) {       // [T, H]: partial last tiles are stored row by row
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sb = smem_u32(smem);
  if ((sb & 1023u) != 0) __trap();
  const int warp = threadIdx.x >> 5;
  const int H = heads * AT3_D;
  const int nq = (S + 127) / 128;
  const int npairs = (nq + 1) / 2;
  const int n_items = B * heads * npairs;

  // barriers (8 B each)
  const uint32_t bar0 = sb + AT3_SMEM_BAR;
  const uint32_t kv_full = bar0;                       // [NST]
  const uint32_t kv_empty = kv_full + 8 * AT3_NST;     // [NST]
  const uint32_t q_full = kv_empty + 8 * AT3_NST;      // [2 buf][2 slot]
  const uint32_t q_empty = q_full + 32;                // [2][2]
  const uint32_t s_ready = q_empty + 32;               // [2 slot][2 sbuf]
  const uint32_t p_ready = s_ready + 32;               // [2][2]
  const uint32_t pv_done = p_ready + 32;               // [2][2]  P_j V_j has completed
  const uint32_t o_ready = pv_done + 32;               // [2 slot]
  const uint32_t o_empty = o_ready + 16;               // [2 slot]
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + AT3_SMEM_BAR + 384);

  if (warp == 8) {
    if (elect_one()) {
      tma_prefetch_desc(&tm_q);
      tma_prefetch_desc(&tm_kv);
      tma_prefetch_desc(&tm_ctx);
      for (int i = 0; i < AT3_NST; ++i) {
        mbar_init(kv_full + 8u * i, 1);
        mbar_init(kv_empty + 8u * i, 2);   // one arrival from each slot's MMA issuer
      }
      for (int i = 0; i < 4; ++i) {
        mbar_init(q_full + 8u * i, 1);
        mbar_init(q_empty + 8u * i, 1);
        mbar_init(s_ready + 8u * i, 1);
        mbar_init(p_ready + 8u * i, 128);
        mbar_init(pv_done + 8u * i, 1);
      }
      for (int i = 0; i < 2; ++i) {
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
  long long* const clk = (blockIdx.x == 0) ? g_att3_clock : nullptr;
  int clk_n = 0;
  // role 0/1: softmax slot A/B (thread 0 of the warpgroup), role 2: MMA issuer; 2 x 256 int64 each.  Compiled in
  // only with V bit 8: the volatile clock reads pin the instruction schedule around them and cost the
  // four-warpgroup kernel 8 % when they were unconditional (profiles/r02_notes.md section 6c).
  constexpr bool kStamp = (V & 256) != 0;
#define AT3_STAMP(role, code)                                          \
  do {                                                                 \
    if constexpr (kStamp) {                                            \
      if (clk != nullptr && clk_n < 256) {                             \
        clk[(role) * 512 + clk_n] = clock64();                         \
        clk[(role) * 512 + 256 + clk_n] = (code);                      \
        ++clk_n;                                                       \
      }                                                                \
    }                                                                  \
  } while (0)

  // waits of the single-thread roles (loader, MMA issuers): flag bit 2 parks the thread in hardware instead of
  // polling -- those threads sit on the sub-partitions of softmax warps 0-2 and 4-6 and their polling loops
  // executed more warp instructions than the kernel has MUFU.EX2 (profiles/r02_ncu_att3_source.md)
  const bool park = (g_att3_flags & 4) != 0;
  auto bg_wait = [&](uint32_t bar, uint32_t parity) {
    if (park) mbar_wait_parked(bar, parity); else mbar_wait(bar, parity);
  };
  if (warp >= 8) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
    if (warp == 9) {
      if (elect_one()) {
        // ------------------------------------------------------------ loader
        uint32_t chunk_ctr = 0;          // ring position, runs across items
        uint32_t q_par = 0, q_any = 0;   // per (buf,slot) bit: (#loads so far) & 1 / #loads > 0
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
            // the buffer is free once the last Q K^T of its previous tile has completed
            if (q_any & bit) bg_wait(q_empty + 8u * idx, ((q_par >> idx) & 1u) ^ 1u);
            const uint32_t qb = q_full + 8u * idx;
            mbar_expect_tx(qb, AT3_QTILE);
            tma_load_2d(sb + AT3_SMEM_Q + idx * AT3_QTILE, &tm_q, qb, h * AT3_D, row_base + t * 128);
            q_par ^= bit;
            q_any |= bit;
          }
          const int n = cur.n;
          for (int j = 0; j < n; ++j, ++chunk_ctr) {
            const int st = chunk_ctr % AT3_NST;
            const uint32_t use = chunk_ctr / AT3_NST;
            if (use > 0) bg_wait(kv_empty + 8u * st, (use - 1) & 1u);
            const uint32_t fb = kv_full + 8u * st;
            mbar_expect_tx(fb, 2 * AT3_KVTILE + AT3_KC * 4);
            const uint32_t dst = sb + AT3_SMEM_KV + st * 2 * AT3_KVTILE;
            const int jk = (cur.j0 + j) * AT3_KC;   // first key of the chunk
            tma_load_2d(dst, &tm_kv, fb, H + h * AT3_D, row_base + jk);
            tma_load_2d(dst + AT3_KVTILE, &tm_kv, fb, 2 * H + h * AT3_D, row_base + jk);
            bulk_load_1d(sb + AT3_SMEM_BIAS + st * AT3_KC * 4,
                         bias + static_cast<size_t>(b) * S_pad + jk, AT3_KC * 4, fb);
            AT3_STAMP(3, it * 100 + j);
          }
          cur = at3_finish(nxt, window);
        }
      }
    } else if (warp == 8 || warp == 10) {
      if (elect_one()) {
        // ------------------------------------------------------------ MMA issuer of ONE slot
        // (warp 8: query tile A, warp 10: query tile B).  One thread driving both slots needed ~450 clk
        // per QK or PV event (polls + 4 MMAs + commits), four events per pair of chunks: 1800 clk,
        // more than the softmax itself.  Two threads run their slots side by side with blocking waits.
        const int slot = (warp == 8) ? 0 : 1;
        constexpr uint32_t idesc_s = make_idesc_h16(128, AT3_KC, 0, 0);
        constexpr uint32_t idesc_o = make_idesc_h16(128, AT3_D, 0, 1);  // B (= V) is MN-major
        const uint32_t t_slot = tmem_base + static_cast<uint32_t>(slot * 256);
        uint32_t chunk_base = 0;   // ring position of this item's chunk 0
        uint32_t q_cnt[2] = {0, 0};   // Q tiles consumed per item buffer (parity of q_full)
        uint32_t p_par = 0;           // bit sbuf: parity of the p_ready phase to wait for
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
          if (slot == 0) AT3_STAMP(2, 9000 + n);
          if (!active) {
            // this slot has no query tile in the item: it still owes the ring one arrival per chunk,
            // and may give it only once the stage has been filled for THIS use
            for (int j = 0; j < n; ++j) {
              const uint32_t c = chunk_base + j;
              bg_wait(kv_full + 8u * (c % AT3_NST), (c / AT3_NST) & 1u);
              mbar_arrive(kv_empty + 8u * (c % AT3_NST));
            }
          } else {
            const int qidx = buf * 2 + slot;
            const uint64_t q_desc = make_smem_desc_sw128(sb + AT3_SMEM_Q + qidx * AT3_QTILE, 16, 1024);
            auto issue_qk = [&](int j) {
              const uint32_t c = chunk_base + j;
              const int st = c % AT3_NST;
              bg_wait(kv_full + 8u * st, (c / AT3_NST) & 1u);
              tc_fence_after();
              const uint64_t k_desc =
                  make_smem_desc_sw128(sb + AT3_SMEM_KV + st * 2 * AT3_KVTILE, 16, 1024);
              const uint32_t d = t_slot + static_cast<uint32_t>((j & 1) * 64);
#pragma unroll
              for (int k = 0; k < AT3_D / 16; ++k)
                tc_mma_f16_ss(d, q_desc + 2u * k, k_desc + 2u * k, idesc_s, static_cast<uint32_t>(k != 0));
              tc_commit(s_ready + 8u * (slot * 2 + (j & 1)));
              if (slot == 0) AT3_STAMP(2, j * 10 + 1);
              if (j + 1 == n) tc_commit(q_empty + 8u * qidx);
            };
            bg_wait(q_full + 8u * qidx, q_cnt[buf] & 1u);
            ++q_cnt[buf];
            issue_qk(0);
            for (int j = 0; j < n; ++j) {
              // S_{j+1} = Q K_{j+1}^T goes out before P_j is awaited: its buffer held P_{j-1}, whose
              // P V was issued one iteration ago (tcgen05 ops of one thread execute in order)
              if (j + 1 < n) issue_qk(j + 1);
              const int sbuf = j & 1;
              bg_wait(p_ready + 8u * (slot * 2 + sbuf), (p_par >> sbuf) & 1u);
              p_par ^= 1u << sbuf;
              // the previous tile's epilogue (o_empty) precedes this tile's first p_ready
              if (j == 0 && tile_cnt > 0) bg_wait(o_empty + 8u * slot, (tile_cnt - 1) & 1u);
              tc_fence_after();
              const uint32_t c = chunk_base + j;
              const int st = c % AT3_NST;
              const uint32_t p = t_slot + static_cast<uint32_t>(sbuf * 64);
              const uint32_t o = t_slot + 128u;
              const uint32_t v_base = sb + AT3_SMEM_KV + st * 2 * AT3_KVTILE + AT3_KVTILE;
#pragma unroll
              for (int k = 0; k < AT3_KC / 16; ++k) {
                const uint64_t v_desc = make_smem_desc_sw128(v_base + k * 16 * 128, 1024, 1024);
                tc_mma_f16_ts(o, p + static_cast<uint32_t>(8 * k), v_desc, idesc_o,
                              static_cast<uint32_t>((j | k) != 0));
              }
              tc_commit(pv_done + 8u * (slot * 2 + sbuf));
              tc_commit(kv_empty + 8u * st);   // this slot is done with the stage (K by Q K^T, V by P V)
              if (slot == 0) AT3_STAMP(2, j * 10 + 2);
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
    // Item parameters are decoded one item AHEAD (two integer divisions and a dependent global load
    // cost ~2000 clk when they sit between two items; here they overlap the current item's work).
    int item = blockIdx.x;
    At3Walk wk;
        wk.init(item, gridDim.x, npairs, heads);
        At3Item cur = at3_finish(at3_fetch(wk, kv_chunks, plain_chunks, n_items, S, seq_cu, seq_len), window);
    // strict alternation A, B, A, B ...: both slots see the same number of chunks in every item
    // bit 0: strict ping-pong on every chunk (measured 3 % slower); bit 1: only the FIRST chunk of an item
    // is ordered (A before B), which merely de-phases the two warpgroups
    const int pp_mode = g_att3_flags & 3;
    const bool pingpong = pp_mode != 0;
    if (pingpong && slot == 1) asm volatile("bar.arrive %0, 256;" ::"r"(4) : "memory");  // A goes first
    uint8_t* ostage = smem + AT3_SMEM_OST + slot * AT3_QTILE;
    const uint32_t ostage_addr = sb + AT3_SMEM_OST + slot * AT3_QTILE;
    for (; item < n_items; item += gridDim.x) {
      wk.step(npairs, heads);
          const At3Raw nxt = at3_fetch(wk, kv_chunks, plain_chunks, n_items, S, seq_cu, seq_len);
      const int pr = cur.pr, h = cur.h, n = cur.n;
      const int t = 2 * pr + slot;
      if (t >= cur.nq && pingpong) {
        // no query tile for this slot in the item: keep the other warpgroup's turns coming
        for (int j = 0; j < ((pp_mode & 1) ? n : 1); ++j) {
          asm volatile("bar.sync %0, 256;" ::"r"(4 + slot) : "memory");
          asm volatile("bar.arrive %0, 256;" ::"r"(4 + (slot ^ 1)) : "memory");
        }
      }
      if (t < cur.nq) {
        constexpr bool kPlainCount = (V & 1) != 0;
        constexpr bool kPrefetch = (V & 2) != 0;
        constexpr int kPoly = (V >> 2) & 3;
        const int n_plain = cur.np;
        float m_used = 0.0f, l = 0.0f;
        uint32_t s0[32], s1[32];
        // wait for S_j = Q K_j^T and start moving it from TMEM into registers (completed by tmem_ld_wait)
        auto fetch_scores = [&](int j) {
          const int sbuf = j & 1;
          mbar_wait(s_ready + 8u * (slot * 2 + sbuf), (s_par >> sbuf) & 1u);
          s_par ^= 1u << sbuf;
          tc_fence_after();
          const uint32_t t_s = t_slot + static_cast<uint32_t>(sbuf * 64);
          tmem_ld32(t_s, s0);
          tmem_ld32(t_s + 32u, s1);
        };
        if (kPrefetch) {
          fetch_scores(0);
          tmem_ld_wait();
        }
        for (int j = 0; j < n; ++j) {
          const int sbuf = j & 1;
          const uint32_t c = chunk_base + j;
          const int st = c % AT3_NST;
          if (r == 0) AT3_STAMP(slot, j * 10 + 0);
          if (!kPrefetch) {
            fetch_scores(j);
            tmem_ld_wait();
          }
          if (r == 0) AT3_STAMP(slot, j * 10 + 1);
          const float* bias_j = reinterpret_cast<const float*>(smem + AT3_SMEM_BIAS + st * AT3_KC * 4);
          const uint32_t t_s = t_slot + static_cast<uint32_t>(sbuf * 64);
          // all 64 keys of the chunk attended?  Either known from the prepared per-sequence count (no
          // shared-memory traffic at all on such chunks), or by a warp vote over the chunk's bias row
          bool plain;
          if (kPlainCount) {
            plain = j < n_plain;
            if (!plain) mbar_wait(kv_full + 8u * st, (c / AT3_NST) & 1u);  // complete: acquires the bias bytes
          } else {
            mbar_wait(kv_full + 8u * st, (c / AT3_NST) & 1u);  // already complete: acquires the bias bytes
            const float2 bz = *reinterpret_cast<const float2*>(bias_j + 2 * (threadIdx.x & 31));
            plain = !__any_sync(0xffffffffu, bz.x != 0.0f || bz.y != 0.0f);
          }
          // Optional ping-pong (experiment, b2e_debug_set_att3_flags bit 0): the exp-heavy part of a chunk
          // runs in ONE warpgroup at a time.  Measured: a warpgroup alone still needs ~950 clk for the 64
          // exponentials per thread (issue/latency bound, MUFU alone would be 512), so strict alternation
          // gives 2 x 950 per pair of chunks, no better than the ~1950 of the free-running version.
          const bool turn = (pp_mode & 1) || (pp_mode == 2 && j == 0);
          if (turn) asm volatile("bar.sync %0, 256;" ::"r"(4 + slot) : "memory");
          uint32_t pk[32];
          if ((V & 16) != 0) {
            // sliding window: key (j0 + j) * 64 + i is visible to query row q iff |q - key| <= window; scores
            // outside the band become a large negative FINITE number (so that a chunk with no visible key
            // degenerates to a uniform softmax that the next visible chunk's rescale wipes out)
            const int q_abs = t * 128 + r;
            const int ilo = q_abs - window - (cur.j0 + j) * AT3_KC;
            const unsigned span = static_cast<unsigned>(2 * window);
            constexpr uint32_t kOut = 0xfcf0bdc2u;   // -1e37f
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              if (static_cast<unsigned>(i - ilo) > span) s0[i] = kOut;
              if (static_cast<unsigned>(i + 32 - ilo) > span) s1[i] = kOut;
            }
          }
          bool done = false;
          auto plain_exp = [&](const uint32_t (&sv)[32], uint32_t* out) {
            if constexpr ((V & 32) != 0) {
              constexpr int kPoly16 = kPoly == 0 ? 0 : 2 + 2 * kPoly;   // 0, 4, 6, 8
              return at3_exp_pack_plain_x2<kPoly16>(sv, scale_log2e, -m_used, out);
            } else {
              return at3_exp_pack_plain<(kPoly > 2 ? 2 : kPoly)>(sv, scale_log2e, -m_used, out);
            }
          };
          if (plain) {
            if (j == 0) {
              // exact maximum of the raw scores first (scale > 0: max commutes with the scaling)
              m_used = scale_log2e * at3_smax_plain(s1, at3_smax_plain(s0, -INFINITY));
              l = plain_exp(s0, pk);
              l += plain_exp(s1, pk + 16);
              done = true;
            } else {
              float sum = plain_exp(s0, pk);
              sum += plain_exp(s1, pk + 16);
              // every p <= row sum: a sum within 2^threshold proves that no score ran away
              // (a NaN sum -- inf - inf cannot occur here -- would fail the test and take the general path)
              const bool calm = sum <= 256.0f;   // 2^AT3_RESCALE_THRESHOLD
              if (__all_sync(0xffffffffu, calm)) {
                l += sum;
                done = true;
              } else if (kPlainCount) {
                mbar_wait(kv_full + 8u * st, (c / AT3_NST) & 1u);   // the general path reads the bias row
              }
            }
          }
          if (done) {
          } else if (j == 0) {
            // first chunk of the row: exact maximum first (always finite: key 0 exists)
            float cmax = at3_max(s0, bias_j, scale_log2e, -INFINITY);
            cmax = at3_max(s1, bias_j + 32, scale_log2e, cmax);
            m_used = cmax;
            float dummy = -INFINITY;
            l = at3_exp_pack(s0, bias_j, scale_log2e, m_used, pk, dummy);
            l += at3_exp_pack(s1, bias_j + 32, scale_log2e, m_used, pk + 16, dummy);
          } else {
            // single pass with the running maximum; redo only if this chunk exceeds it by > 2^8
            float xmax = -INFINITY;
            float sum = at3_exp_pack(s0, bias_j, scale_log2e, m_used, pk, xmax);
            sum += at3_exp_pack(s1, bias_j + 32, scale_log2e, m_used, pk + 16, xmax);
            const bool need = xmax > m_used + AT3_RESCALE_THRESHOLD;
            if (__any_sync(0xffffffffu, need)) {
              const float m_new = need ? xmax : m_used;
              const float sc = fast_exp2(m_used - m_new);  // 1 for rows that keep their maximum
              m_used = m_new;
              l *= sc;
              float dummy = -INFINITY;
              sum = at3_exp_pack(s0, bias_j, scale_log2e, m_used, pk, dummy);
              sum += at3_exp_pack(s1, bias_j + 32, scale_log2e, m_used, pk + 16, dummy);
              // O = sum_{i<j} P_i V_i must be complete before it is rescaled: S_j was issued ahead
              // of P_{j-1} V_{j-1}, so wait for that MMA explicitly (its barrier has seen exactly
              // as many phases as this warpgroup has published P chunks on that buffer)
              const int pb = sbuf ^ 1;
              mbar_wait(pv_done + 8u * (slot * 2 + pb), ((s_par >> pb) & 1u) ^ 1u);
              tc_fence_after();
#pragma unroll 1
              for (int cc = 0; cc < 2; ++cc) {
                uint32_t o[32];
                tmem_ld32(t_o + static_cast<uint32_t>(cc * 32), o);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * sc);
                tmem_st32(t_o + static_cast<uint32_t>(cc * 32), o);
              }
            }
            l += sum;
          }
          if (turn) asm volatile("bar.arrive %0, 256;" ::"r"(4 + (slot ^ 1)) : "memory");
          if (r == 0) AT3_STAMP(slot, j * 10 + 2);
          tmem_st32(t_s, pk);  // h16 P over the first 32 columns of S's own buffer
          // S_{j+1} (the other buffer; its Q K^T was issued before P_{j-1} V_{j-1}) starts to move into the
          // score registers now: the load runs under the store's wait, the fence and the arrive
          if (kPrefetch && j + 1 < n) fetch_scores(j + 1);
          tmem_st_wait();
          tc_fence_before();
          mbar_arrive(p_ready + 8u * (slot * 2 + sbuf));
          if (kPrefetch && j + 1 < n) tmem_ld_wait();
          if (r == 0) AT3_STAMP(slot, j * 10 + 3);
        }
        // ---- epilogue: O / l -> h16 -> swizzled staging tile -> one TMA store per tile
        if (r == 0) AT3_STAMP(slot, 900);
        if (r == 0) tma_store_wait_read<0>();   // the previous tile's store has read the staging
        asm volatile("bar.sync %0, 128;" ::"r"(2 + slot) : "memory");
        mbar_wait(o_ready + 8u * slot, o_cnt & 1u);
        if (r == 0) AT3_STAMP(slot, 901);
        ++o_cnt;
        tc_fence_after();
        const float inv_l = 1.0f / l;
#pragma unroll 1
        for (int cc = 0; cc < 2; ++cc) {
          uint32_t o[32];
          tmem_ld32(t_o + static_cast<uint32_t>(cc * 32), o);
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
        tc_fence_before();
        mbar_arrive(o_empty + 8u * slot);   // O's TMEM columns may be overwritten by the next tile
        fence_proxy_async_smem();
        asm volatile("bar.sync %0, 128;" ::"r"(2 + slot) : "memory");
        const int valid = cur.len - t * 128;   // rows of this tile that belong to the sequence
        if (valid >= 128) {
          if (r == 0) {
            tma_store_2d(&tm_ctx, ostage_addr, h * AT3_D, cur.row0 + t * 128);
            tma_store_commit();
            AT3_STAMP(slot, 902);
          }
        } else if (r < valid) {
          // last, partial tile of the sequence: the rows behind it belong to the NEXT sequence (packed layout)
          // or do not exist -- every thread stores its own row (it staged that row itself: no hazard)
          h16* dst = ctx_out + static_cast<size_t>(cur.row0 + t * 128 + r) * H + h * AT3_D;
#pragma unroll
          for (int u = 0; u < 8; ++u)
            *reinterpret_cast<uint4*>(dst + u * 8) =
                *reinterpret_cast<const uint4*>(ostage + r * 128 + ((u ^ (r & 7)) << 4));
        }
      }
      chunk_base += static_cast<uint32_t>(n);
      cur = at3_finish(nxt, window);
    }
    if (r == 0) tma_store_wait_all();
  }

#undef AT3_STAMP
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem_base, 512);
}

}  // namespace b2e
