// Binary ("ubinary") retrieval with float rescoring over a device-resident packed-bit corpus.
//
// Replaces, for precision='ubinary' / search_algorithm='exact', what distllm/rag/search.py:202-260 builds
// (faiss.IndexBinaryFlat over sentence_transformers' packbits(x > 0)) and what :280-336 runs through
// semantic_search_faiss(rescore=True, rescore_multiplier): Hamming top-(k * multiplier) on the packed bits,
// then  score = sum_j q[j] * bit[j]  of the FLOAT query against each candidate's unpacked bits, top k by score.
//
// HBM-bound integer / byte work (H/8 bytes per row: 96 B at H = 768, 10 M rows = 0.96 GB): no tensor cores.
//   pack_ubinary_kernel      fp32 rows -> packed bits (first dimension in the most significant bit)
//   hamming_hist_kernel      pass 1: one 16-byte vector load per lane and word group, XOR + POPC, per-CTA
//                            shared-memory histogram of the distances (0..H), merged with one atomic per bin
//   hamming_threshold_kernel distance t of the K-th nearest row and how many rows are strictly closer
//   hamming_select_kernel    pass 2: rows with d <= t appended to the candidate list (warp-aggregated)
//   binary_rescore_kernel    one CTA per query: sort candidates by (distance, id) -- among equal distances the
//                            smaller ids stay, as IndexBinaryFlat's strictly-closer replacement implies --
//                            keep K, rescore with the float query, sort by descending score, emit top k
#pragma once

#include "common.cuh"

namespace b2e {

constexpr int BIN_THREADS = 256;
constexpr int BIN_MAX_Q = 8;          // queries per pass over the corpus
constexpr int BIN_MAX_WORDS = 256;    // H <= 8192
constexpr int BIN_MAX_CAND = 4096;    // candidates per query handed to the rescoring kernel

// out[row, j/8] bit (7 - j%8) = emb[row, j] > 0      (np.packbits, bitorder 'big')
__global__ void pack_ubinary_kernel(const float* __restrict__ emb, uint8_t* __restrict__ out, long long N,
                                    int H) {
  const long long total = N * (H / 8);
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float4 a = *reinterpret_cast<const float4*>(emb + i * 8);
    const float4 b = *reinterpret_cast<const float4*>(emb + i * 8 + 4);
    uint32_t v = 0;
    v |= (a.x > 0.0f) << 7; v |= (a.y > 0.0f) << 6; v |= (a.z > 0.0f) << 5; v |= (a.w > 0.0f) << 4;
    v |= (b.x > 0.0f) << 3; v |= (b.y > 0.0f) << 2; v |= (b.z > 0.0f) << 1; v |= (b.w > 0.0f) << 0;
    out[i] = static_cast<uint8_t>(v);
  }
}

// Hamming distance of one row (W 32-bit words, 16-byte aligned when W % 4 == 0) to Q queries held in shared
// memory as [Q][W] words.
template <int Q>
__device__ __forceinline__ void row_distances(const uint32_t* __restrict__ row, const uint32_t* __restrict__ qs,
                                              int W, int (&d)[Q]) {
#pragma unroll
  for (int q = 0; q < Q; ++q) d[q] = 0;
  if ((W & 3) == 0) {
    for (int w = 0; w < W; w += 4) {
      const uint4 r = __ldg(reinterpret_cast<const uint4*>(row + w));
#pragma unroll
      for (int q = 0; q < Q; ++q) {
        const uint4 x = *reinterpret_cast<const uint4*>(qs + q * W + w);
        d[q] += __popc(r.x ^ x.x) + __popc(r.y ^ x.y) + __popc(r.z ^ x.z) + __popc(r.w ^ x.w);
      }
    }
  } else {
    for (int w = 0; w < W; ++w) {
      const uint32_t r = __ldg(row + w);
#pragma unroll
      for (int q = 0; q < Q; ++q) d[q] += __popc(r ^ qs[q * W + w]);
    }
  }
}

// pass 1: hist[q][d] += 1 for every row.  Dynamic shared memory: Q*W words of query bits + Q*(H+1) counters.
template <int Q>
__global__ void __launch_bounds__(BIN_THREADS)
hamming_hist_kernel(const uint32_t* __restrict__ corpus, const uint32_t* __restrict__ queries, long long N,
                    int W, int H, unsigned* __restrict__ hist) {
  extern __shared__ __align__(16) uint32_t bs_smem[];
  uint32_t* qs = bs_smem;
  unsigned* h = reinterpret_cast<unsigned*>(bs_smem + Q * W);
  for (int i = threadIdx.x; i < Q * W; i += blockDim.x) qs[i] = queries[i];
  for (int i = threadIdx.x; i < Q * (H + 1); i += blockDim.x) h[i] = 0;
  __syncthreads();
  for (long long r = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; r < N;
       r += static_cast<long long>(gridDim.x) * blockDim.x) {
    int d[Q];
    row_distances<Q>(corpus + r * W, qs, W, d);
#pragma unroll
    for (int q = 0; q < Q; ++q) atomicAdd(&h[q * (H + 1) + d[q]], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < Q * (H + 1); i += blockDim.x)
    if (h[i]) atomicAdd(&hist[i], h[i]);
}

// per query: t = smallest distance with count(d <= t) >= K (K clipped to N); thr[q] = {t, count(d < t)}
__global__ void hamming_threshold_kernel(const unsigned* __restrict__ hist, int H, long long K, int Q,
                                         int* __restrict__ thr) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= Q) return;
  long long acc = 0;
  int t = H;
  long long below = 0;
  for (int d = 0; d <= H; ++d) {
    const long long c = hist[q * (H + 1) + d];
    if (acc + c >= K) {
      t = d;
      below = acc;
      break;
    }
    acc += c;
    below = acc;
  }
  thr[2 * q] = t;
  thr[2 * q + 1] = static_cast<int>(below < 0x7fffffff ? below : 0x7fffffff);
}

// pass 2: every row with d <= t[q] goes to cand[q][...] as (distance << 40 | row id); n_cand[q] counts them
// (it may exceed `cap`: the caller checks).  Warp-aggregated appends.
template <int Q>
__global__ void __launch_bounds__(BIN_THREADS)
hamming_select_kernel(const uint32_t* __restrict__ corpus, const uint32_t* __restrict__ queries, long long N,
                      int W, const int* __restrict__ thr, unsigned long long* __restrict__ cand,
                      unsigned* __restrict__ n_cand, unsigned cap) {
  extern __shared__ __align__(16) uint32_t bs_smem[];
  uint32_t* qs = bs_smem;
  for (int i = threadIdx.x; i < Q * W; i += blockDim.x) qs[i] = queries[i];
  __syncthreads();
  int t[Q];
#pragma unroll
  for (int q = 0; q < Q; ++q) t[q] = thr[2 * q];
  const int lane = threadIdx.x & 31;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  const long long first = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  // whole warps iterate together (the ballots below need every lane): loop on the warp's first row
  for (long long base = first - lane; base < N; base += stride) {
    const long long r = base + lane;
    int d[Q];
    if (r < N) {
      row_distances<Q>(corpus + r * W, qs, W, d);
    } else {
#pragma unroll
      for (int q = 0; q < Q; ++q) d[q] = 0x7fffffff;
    }
#pragma unroll
    for (int q = 0; q < Q; ++q) {
      const bool take = d[q] <= t[q];
      const unsigned m = __ballot_sync(0xffffffffu, take);
      if (m == 0) continue;
      unsigned pos = 0;
      if (lane == __ffs(m) - 1) pos = atomicAdd(&n_cand[q], __popc(m));
      pos = __shfl_sync(0xffffffffu, pos, __ffs(m) - 1) + __popc(m & ((1u << lane) - 1u));
      if (take && pos < cap)
        cand[static_cast<size_t>(q) * cap + pos] =
            (static_cast<unsigned long long>(d[q]) << 40) | static_cast<unsigned long long>(r);
    }
  }
}

__device__ __forceinline__ void bitonic_sort_u64(unsigned long long* a, int n_pow2) {
  for (int k = 2; k <= n_pow2; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < n_pow2; i += blockDim.x) {
        const int p = i ^ j;
        if (p > i) {
          const unsigned long long x = a[i], y = a[p];
          const bool up = (i & k) == 0;
          if ((x > y) == up) {
            a[i] = y;
            a[p] = x;
          }
        }
      }
      __syncthreads();
    }
}

// One CTA per query.  cand holds n (<= cap) keys (distance << 40 | id); the K smallest by (distance, id) are
// rescored with the float query and the top k by (descending score, candidate rank) are written.
// Dynamic shared memory: n_pow2 * 8 bytes (keys) + H * 4 (query) .
__global__ void __launch_bounds__(BIN_THREADS)
binary_rescore_kernel(const unsigned long long* __restrict__ cand, const unsigned* __restrict__ n_cand,
                      unsigned cap, int n_pow2, const uint32_t* __restrict__ corpus, int W, int H,
                      const float* __restrict__ queries, long long K, int k, float* __restrict__ out_score,
                      long long* __restrict__ out_index) {
  extern __shared__ __align__(16) unsigned long long rs_keys[];
  float* qf = reinterpret_cast<float*>(rs_keys + n_pow2);
  const int q = blockIdx.x;
  unsigned n = n_cand[q];
  if (n > cap) {
    // more rows tie at the threshold distance than the candidate buffer holds (a corpus of duplicates): the
    // result would depend on the order of the atomics -- flag it instead (scores NaN, indices -2)
    for (int i = threadIdx.x; i < k; i += blockDim.x) {
      out_score[static_cast<size_t>(q) * k + i] = __int_as_float(0x7fc00000);
      out_index[static_cast<size_t>(q) * k + i] = -2;
    }
    return;
  }
  for (int i = threadIdx.x; i < n_pow2; i += blockDim.x)
    rs_keys[i] = i < static_cast<int>(n) ? cand[static_cast<size_t>(q) * cap + i] : ~0ull;
  for (int i = threadIdx.x; i < H; i += blockDim.x) qf[i] = queries[static_cast<size_t>(q) * H + i];
  __syncthreads();
  bitonic_sort_u64(rs_keys, n_pow2);          // ascending (distance, id)
  const int kk = static_cast<int>(K < static_cast<long long>(n) ? K : n);
  // rescoring: one warp per candidate, lane l owns words l, l+32, ...; key := (~score bits, rank) for a
  // DEscending sort by score with ties in candidate order
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int c = warp; c < kk; c += nwarps) {
    const unsigned long long id = rs_keys[c] & ((1ull << 40) - 1ull);
    const uint32_t* row = corpus + id * W;
    float s = 0.0f;
    for (int w = lane; w < W; w += 32) {
      // bytes are packed MSB-first: dimension 32*w + 8*b + j is bit (7 - j) of byte b of the word
      const uint32_t word = __ldg(row + w);
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const uint32_t byte = (word >> (8 * b)) & 0xffu;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (byte & (0x80u >> j)) s += qf[32 * w + 8 * b + j];
      }
    }
    s = warp_sum(s);
    __syncwarp();
    if (lane == 0) {
      // order-preserving map of the float to an unsigned key, inverted for descending order
      uint32_t u = __float_as_uint(s);
      u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
      rs_keys[c] = (static_cast<unsigned long long>(~u) << 32) | (static_cast<unsigned long long>(c) << 8) |
                   0ull;
      // the id is recovered through a side array: reuse the upper half of the key buffer
      rs_keys[n_pow2 / 2 + c] = id;   // safe: kk <= n_pow2 / 2 is guaranteed by the launcher
    }
  }
  __syncthreads();
  for (int i = kk + threadIdx.x; i < n_pow2 / 2; i += blockDim.x) rs_keys[i] = ~0ull;
  __syncthreads();
  // sort the first n_pow2/2 keys (scores); ids sit untouched in the upper half
  bitonic_sort_u64(rs_keys, n_pow2 / 2);
  for (int i = threadIdx.x; i < k; i += blockDim.x) {
    if (i < kk) {
      const unsigned long long key = rs_keys[i];
      const int c = static_cast<int>((key >> 8) & 0xffffffu);
      uint32_t u = ~static_cast<uint32_t>(key >> 32);
      u = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
      out_score[static_cast<size_t>(q) * k + i] = __uint_as_float(u);
      out_index[static_cast<size_t>(q) * k + i] = static_cast<long long>(rs_keys[n_pow2 / 2 + c]);
    } else {
      out_score[static_cast<size_t>(q) * k + i] = -INFINITY;
      out_index[static_cast<size_t>(q) * k + i] = -1;
    }
  }
}

}  // namespace b2e
