// Padding-free ("packed") token layout for the pooled forward pass.
//
// The reference pads every batch to its longest sequence (distllm/embed/datasets/utils.py:43-50) and runs the
// encoder over all B*S positions (full_sequence.py:57-59).  Only attended tokens can influence a pooled
// embedding: padded KEYS are masked in attention and padded ROWS are dropped by both poolers.  When every mask
// row is a non-empty prefix (right padding, what `tokenizer(..., padding=True)` produces) the forward pass
// therefore runs on the T' = sum(len_b) attended tokens only, stored back to back:
//
//   rows of sequence b:   cu[b] .. cu[b] + len[b] - 1        (cu = exclusive prefix sum of len)
//   tok_src[t]        :   b*S + s of packed row t            (token id / position / pooling weight lookups)
//   t_real[0]         :   T'                                  (row count of every GEMM / row kernel, read ON DEVICE:
//                                                              nothing is synchronised with the host)
//
// The batch keeps its composition (order, size), so the mean pooler's cross-row quirk is untouched.  Masks with
// holes or left padding, empty rows, or pack == 0 give the identity layout (cu[b] = b*S, len[b] = S): the same
// kernels then reproduce the padded computation.
#pragma once

#include "common.cuh"

namespace b2e {

// one warp per sequence: len_raw[b] = sum(mask[b,:]), ok[b] = mask row is 1...1 0...0 with at least one 1
__global__ void pack_lengths_kernel(const int64_t* __restrict__ mask, int* __restrict__ len_raw,
                                    int* __restrict__ ok, int B, int S) {
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  const int lane = threadIdx.x & 31;
  int n = 0, last = 0;
  for (int s = lane; s < S; s += 32) {
    if (mask[static_cast<size_t>(b) * S + s] != 0) {
      ++n;
      last = s + 1;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    n += __shfl_xor_sync(0xffffffffu, n, o);
    last = max(last, __shfl_xor_sync(0xffffffffu, last, o));
  }
  if (lane == 0) {
    len_raw[b] = n;
    ok[b] = (n > 0 && last == n) ? 1 : 0;   // as many ones as the position of the last one: a prefix
  }
}

// single block: packed iff `enable` and every row is a non-empty prefix mask; len, cu (B + 1 entries), t_real
__global__ void pack_scan_kernel(const int* __restrict__ len_raw, const int* __restrict__ ok,
                                 int* __restrict__ len, int* __restrict__ cu, int* __restrict__ t_real,
                                 int B, int S, int enable) {
  __shared__ int all_ok;
  if (threadIdx.x == 0) all_ok = enable;
  __syncthreads();
  for (int b = threadIdx.x; b < B; b += blockDim.x)
    if (!ok[b]) atomicAnd(&all_ok, 0);
  __syncthreads();
  const int packed = all_ok;
  for (int b = threadIdx.x; b < B; b += blockDim.x) len[b] = packed ? len_raw[b] : S;
  __syncthreads();
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int b = 0; b < B; ++b) {
      cu[b] = acc;
      acc += len[b];
    }
    cu[B] = acc;
    t_real[0] = acc;
    t_real[1] = packed;
  }
}

// tok_src[cu[b] + s] = b*S + s
__global__ void pack_fill_kernel(const int* __restrict__ len, const int* __restrict__ cu,
                                 int* __restrict__ tok_src, int B, int S) {
  const int b = blockIdx.y;
  const int n = len[b], base = cu[b];
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < n; s += gridDim.x * blockDim.x)
    tok_src[base + s] = b * S + s;
}

}  // namespace b2e
