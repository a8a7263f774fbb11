// Fused bidirectional self-attention for head_dim 64, S <= 512 (BERT family) on sm_100a.
//
// One CTA = one (sequence b, head h, 128-query tile).  Everything for the tile stays on chip:
//   S = Q.K^T   tcgen05.mma 128 x 128 x 16 per 128-key tile, fp32 scores in TMEM (<= 512 columns)
//   softmax     4 warps, one query row per thread: tcgen05.ld -> scale+mask -> max -> exp2 -> bf16
//               P written to 128B-swizzled smem (K-major A operand), 64 keys per chunk
//   O = P.V     tcgen05.mma 128 x 64 x 16, V consumed straight from its TMA tile as an MN-major
//               B operand; O accumulates in TMEM columns [0,64) (the S columns already consumed)
//   epilogue    tcgen05.ld O -> * 1/rowsum -> bf16 -> ctx[b*S+q, h*64 ..]
//
// Semantics follow HF BERT's SDPA path (transformers/models/bert/modeling_bert.py:192-205 with the
// additive padding mask built at :692-716): masked keys get the most negative finite score, so a
// sequence whose mask is all zero degenerates to a uniform distribution exactly like the reference.
#pragma once

#include "common.cuh"

namespace b2e {

constexpr int ATT_D = 64;
constexpr int ATT_BQ = 128;
constexpr int ATT_MAX_S = 512;
constexpr int ATT_THREADS = 160;  // 4 softmax warps + 1 control warp
constexpr int ATT_TILE_BYTES = 128 * ATT_D * 2;  // 16 KiB: 128 rows x 128 B
constexpr int ATT_SMEM_Q = 0;
constexpr int ATT_SMEM_K = ATT_SMEM_Q + ATT_TILE_BYTES;
constexpr int ATT_SMEM_V = ATT_SMEM_K + 4 * ATT_TILE_BYTES;
constexpr int ATT_SMEM_X = ATT_SMEM_V + 4 * ATT_TILE_BYTES;     // P chunks 4..7
constexpr int ATT_SMEM_BIAS = ATT_SMEM_X + 4 * ATT_TILE_BYTES;  // 512 floats
constexpr int ATT_SMEM_BAR = ATT_SMEM_BIAS + ATT_MAX_S * 4;
constexpr int ATT_SMEM_BYTES = ATT_SMEM_BAR + 256 + 1024;

constexpr float ATT_MASKED = -3.0e38f;

__global__ void __launch_bounds__(ATT_THREADS, 1)
attention_d64_tcgen05_kernel(const __grid_constant__ CUtensorMap tm_qkv,  // [T, 3H] bf16, box 64x128
                             const int64_t* __restrict__ attn_mask,        // [B, S] (0 = padded)
                             bf16* __restrict__ ctx,                       // [T, H]
                             int S, int H, float scale_log2e, float* __restrict__ dbg_scores) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t pad = ((raw + 1023u) & ~1023u) - raw;
  uint8_t* smem = smem_raw + pad;
  const uint32_t sb = raw + pad;

  const int q_tile = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int q0 = q_tile * ATT_BQ;
  const int nkt = (S + 127) / 128;  // 128-key tiles
  const int nk = nkt * 128;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  float* sbias = reinterpret_cast<float*>(smem + ATT_SMEM_BIAS);
  const uint32_t bar_qk = sb + ATT_SMEM_BAR;
  const uint32_t bar_v = bar_qk + 8;
  const uint32_t bar_s = bar_qk + 16;
  const uint32_t bar_o = bar_qk + 24;
  const uint32_t bar_p = bar_qk + 32;  // 8 barriers
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + ATT_SMEM_BAR + 96);

  if (warp == 4) {
    if (elect_one()) {
      tma_prefetch_desc(&tm_qkv);
      mbar_init(bar_qk, 1);
      mbar_init(bar_v, 1);
      mbar_init(bar_s, 1);
      mbar_init(bar_o, 1);
      for (int c = 0; c < 8; ++c) mbar_init(bar_p + 8u * c, 128);
      mbar_fence_init();
    }
    __syncwarp();
    tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_slot)), 512);
  } else {
    // additive key bias in the exp2 domain: 0 valid, most-negative-finite padded, -inf beyond S
    for (int j = threadIdx.x; j < nk; j += 128) {
      float v = -INFINITY;
      if (j < S) v = (attn_mask[static_cast<size_t>(b) * S + j] != 0) ? 0.0f : ATT_MASKED;
      sbias[j] = v;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int row_base = b * S;

  if (warp == 4) {
    if (elect_one()) {
      // ---- loads
      mbar_expect_tx(bar_qk, (1 + nkt) * ATT_TILE_BYTES);
      tma_load_2d(sb + ATT_SMEM_Q, &tm_qkv, bar_qk, h * ATT_D, row_base + q0);
      for (int kt = 0; kt < nkt; ++kt)
        tma_load_2d(sb + ATT_SMEM_K + kt * ATT_TILE_BYTES, &tm_qkv, bar_qk, H + h * ATT_D,
                    row_base + kt * 128);
      mbar_expect_tx(bar_v, nkt * ATT_TILE_BYTES);
      for (int kt = 0; kt < nkt; ++kt)
        tma_load_2d(sb + ATT_SMEM_V + kt * ATT_TILE_BYTES, &tm_qkv, bar_v, 2 * H + h * ATT_D,
                    row_base + kt * 128);
      // ---- S = Q K^T
      mbar_wait(bar_qk, 0);
      tc_fence_after();
      constexpr uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0);
      const uint64_t q_desc = make_smem_desc_sw128(sb + ATT_SMEM_Q, 16, 1024);
      for (int kt = 0; kt < nkt; ++kt) {
        const uint64_t k_desc =
            make_smem_desc_sw128(sb + ATT_SMEM_K + kt * ATT_TILE_BYTES, 16, 1024);
#pragma unroll
        for (int k = 0; k < ATT_D / 16; ++k)
          tc_mma_f16_ss(tmem_base + static_cast<uint32_t>(kt * 128), q_desc + 2u * k,
                        k_desc + 2u * k, idesc_s, static_cast<uint32_t>(k != 0));
      }
      tc_commit(bar_s);
      // ---- O = P V, one 64-key chunk at a time as the softmax warps publish P
      mbar_wait(bar_v, 0);
      constexpr uint32_t idesc_o = make_idesc_bf16(128, ATT_D, 0, 1);  // B (=V) is MN-major
      for (int c = 0; c < 2 * nkt; ++c) {
        mbar_wait(bar_p + 8u * c, 0);
        tc_fence_after();
        const uint32_t p_addr =
            sb + (c < 4 ? ATT_SMEM_K + c * ATT_TILE_BYTES : ATT_SMEM_X + (c - 4) * ATT_TILE_BYTES);
        const uint64_t p_desc = make_smem_desc_sw128(p_addr, 16, 1024);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          // 16 keys = 16 rows of 128 B = two 8-row swizzle atoms
          const uint64_t v_desc =
              make_smem_desc_sw128(sb + ATT_SMEM_V + (c * 64 + k * 16) * 128, 1024, 1024);
          tc_mma_f16_ss(tmem_base, p_desc + 2u * k, v_desc, idesc_o,
                        static_cast<uint32_t>((c | k) != 0));
        }
      }
      tc_commit(bar_o);
    }
  } else {
    const int r = threadIdx.x;  // query row inside the tile == TMEM lane
    const uint32_t t_row = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    mbar_wait(bar_s, 0);
    tc_fence_after();

    // ---- pass 1: row max of scale*s + bias
    float m = -INFINITY;
    for (int c = 0; c < nk / 32; ++c) {
      uint32_t s[32];
      tmem_ld32(t_row + static_cast<uint32_t>(c * 32), s);
      tmem_ld_wait();
      if (dbg_scores != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) {
#pragma unroll
        for (int j = 0; j < 32; ++j) dbg_scores[r * ATT_MAX_S + c * 32 + j] = __uint_as_float(s[j]);
      }
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float4 bz = *reinterpret_cast<const float4*>(sbias + c * 32 + j);
        m = fmaxf(m, fmaf(__uint_as_float(s[j + 0]), scale_log2e, bz.x));
        m = fmaxf(m, fmaf(__uint_as_float(s[j + 1]), scale_log2e, bz.y));
        m = fmaxf(m, fmaf(__uint_as_float(s[j + 2]), scale_log2e, bz.z));
        m = fmaxf(m, fmaf(__uint_as_float(s[j + 3]), scale_log2e, bz.w));
      }
    }

    // ---- pass 2: p = exp2(x - m) -> bf16 P chunks in smem, row sum in fp32
    float l = 0.0f;
    for (int c = 0; c < nk / 32; ++c) {
      uint32_t s[32];
      tmem_ld32(t_row + static_cast<uint32_t>(c * 32), s);
      tmem_ld_wait();
      const int pc = c >> 1;  // 64-key chunk
      uint8_t* p_tile =
          smem + (pc < 4 ? ATT_SMEM_K + pc * ATT_TILE_BYTES : ATT_SMEM_X + (pc - 4) * ATT_TILE_BYTES);
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        const float4 b0 = *reinterpret_cast<const float4*>(sbias + c * 32 + j);
        const float4 b1 = *reinterpret_cast<const float4*>(sbias + c * 32 + j + 4);
        float p[8];
        p[0] = fast_exp2(fmaf(__uint_as_float(s[j + 0]), scale_log2e, b0.x) - m);
        p[1] = fast_exp2(fmaf(__uint_as_float(s[j + 1]), scale_log2e, b0.y) - m);
        p[2] = fast_exp2(fmaf(__uint_as_float(s[j + 2]), scale_log2e, b0.z) - m);
        p[3] = fast_exp2(fmaf(__uint_as_float(s[j + 3]), scale_log2e, b0.w) - m);
        p[4] = fast_exp2(fmaf(__uint_as_float(s[j + 4]), scale_log2e, b1.x) - m);
        p[5] = fast_exp2(fmaf(__uint_as_float(s[j + 5]), scale_log2e, b1.y) - m);
        p[6] = fast_exp2(fmaf(__uint_as_float(s[j + 6]), scale_log2e, b1.z) - m);
        p[7] = fast_exp2(fmaf(__uint_as_float(s[j + 7]), scale_log2e, b1.w) - m);
        l += ((p[0] + p[1]) + (p[2] + p[3])) + ((p[4] + p[5]) + (p[6] + p[7]));
        uint4 o;
        o.x = pack_bf16x2(p[0], p[1]);
        o.y = pack_bf16x2(p[2], p[3]);
        o.z = pack_bf16x2(p[4], p[5]);
        o.w = pack_bf16x2(p[6], p[7]);
        const int unit = (c & 1) * 4 + (j >> 3);  // 16-byte unit inside the 128-byte row
        *reinterpret_cast<uint4*>(p_tile + r * 128 + ((unit ^ (r & 7)) << 4)) = o;
      }
      if (c & 1) {
        fence_proxy_async_smem();
        tc_fence_before();
        mbar_arrive(bar_p + 8u * pc);
      }
    }

    // ---- epilogue: O / l -> ctx
    mbar_wait(bar_o, 0);
    tc_fence_after();
    const float inv_l = 1.0f / l;
    const int q = q0 + r;
    bf16* dst = ctx + static_cast<size_t>(row_base + q) * H + h * ATT_D;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t o[32];
      tmem_ld32(t_row + static_cast<uint32_t>(c * 32), o);
      tmem_ld_wait();
      if (q < S) {
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          uint4 w;
          w.x = pack_bf16x2(__uint_as_float(o[j + 0]) * inv_l, __uint_as_float(o[j + 1]) * inv_l);
          w.y = pack_bf16x2(__uint_as_float(o[j + 2]) * inv_l, __uint_as_float(o[j + 3]) * inv_l);
          w.z = pack_bf16x2(__uint_as_float(o[j + 4]) * inv_l, __uint_as_float(o[j + 5]) * inv_l);
          w.w = pack_bf16x2(__uint_as_float(o[j + 6]) * inv_l, __uint_as_float(o[j + 7]) * inv_l);
          *reinterpret_cast<uint4*>(dst + c * 32 + j) = w;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem_base, 512);
}

}  // namespace b2e
