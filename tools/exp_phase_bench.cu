// Micro-benchmark of the softmax "exp phase" of the attention kernels, one SM at a time.
//
// Question (profiles/r01_notes.md): a softmax warp ALONE on its SMSP needs ~950 clk for the 64 exponentials
// of a 64-key chunk although the MUFU pipe (16 ex2/clk/SM = one warp instruction per 8 clk per SMSP) would
// allow 512.  Which instruction class costs the rest, and what does moving part of the exponentials to the
// FMA pipe (Cody-Waite range reduction + degree-3 polynomial) buy?
//
// Each warp keeps 64 values per thread in registers and repeats the chunk body ITER times; results feed back
// into the next iteration (64 independent chains, so the ILP of the real loop is kept).  Reported: clk per
// chunk body (64 elements per thread) for 1 warp per SMSP (4 warps) and 2 warps per SMSP (8 warps).
//
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/bin/exp_phase_bench tools/exp_phase_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdint>

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
// 2^x for x in [-126, 126] on the FMA / ALU pipes: n = round(x), f = x - n in [-0.5, 0.5],
// 2^f by a degree-4 polynomial (rel. error 2.7e-6), exponent added as integer.
__device__ __forceinline__ float ex2_poly(float x) {
  x = fmaxf(x, -126.0f);
  const float t = x + 12582912.0f;          // 1.5 * 2^23: the integer part lands in the low mantissa bits
  const float n = t - 12582912.0f;
  const float f = x - n;
  float p = 0.009560510f;                    // degree-4 least-squares fit of 2^f on [-0.5, 0.5] (2.7e-6)
  p = fmaf(p, f, 0.055917039f);
  p = fmaf(p, f, 0.240249811f);
  p = fmaf(p, f, 0.693121968f);
  p = fmaf(p, f, 0.999999191f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}

// ---- packed fp32x2 arithmetic (sm_100: FFMA2 / FADD2): one issue slot for two elements
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk2(float lo, float hi) {
  u64 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpk2(u64 v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
  u64 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ u64 add2(u64 a, u64 b) {
  u64 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ u64 sub2(u64 a, u64 b) {
  u64 r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
// two exponentials on the FMA pipe, pairwise
__device__ __forceinline__ void ex2_poly2(u64 x, float& p_lo, float& p_hi) {
  float x0, x1;
  unpk2(x, x0, x1);
  x = pk2(fmaxf(x0, -126.0f), fmaxf(x1, -126.0f));
  const u64 magic = pk2(12582912.0f, 12582912.0f);
  const u64 t = add2(x, magic);
  const u64 n = sub2(t, magic);
  const u64 f = sub2(x, n);
  u64 p = pk2(0.009560510f, 0.009560510f);
  p = fma2(p, f, pk2(0.055917039f, 0.055917039f));
  p = fma2(p, f, pk2(0.240249811f, 0.240249811f));
  p = fma2(p, f, pk2(0.693121968f, 0.693121968f));
  p = fma2(p, f, pk2(0.999999191f, 0.999999191f));
  float t0, t1, q0, q1;
  unpk2(t, t0, t1);
  unpk2(p, q0, q1);
  p_lo = __int_as_float(__float_as_int(q0) + (__float_as_int(t0) << 23));
  p_hi = __int_as_float(__float_as_int(q1) + (__float_as_int(t1) << 23));
}

// MODE 7..10: the shipped chunk body with fp32x2 arithmetic; POLY16 = exponentials per 16 on the FMA pipe
template <int POLY16>
__global__ void __launch_bounds__(256, 1) bench2(const float* __restrict__ in, float* __restrict__ out,
                                                long long* __restrict__ clk, int iters, float scale, float negm) {
  float s[64];
#pragma unroll
  for (int i = 0; i < 64; ++i) s[i] = in[(threadIdx.x * 64 + i) & 4095];
  u64 sum2 = 0;
  uint32_t acc = 0;
  const u64 sc2 = pk2(scale, scale), nm2 = pk2(negm, negm);
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 64; i += 16) {
#pragma unroll
      for (int k = 0; k < 16; k += 2) {
        const u64 x = fma2(pk2(s[i + k], s[i + k + 1]), sc2, nm2);
        float p0, p1;
        // pairs 7, 3, 5, 1 of the eight go to the FMA pipe first (spread over the group)
        const bool poly = (POLY16 >= 2 && k == 14) || (POLY16 >= 4 && k == 6) || (POLY16 >= 6 && k == 10) ||
                          (POLY16 >= 8 && k == 2);
        if (poly) {
          ex2_poly2(x, p0, p1);
        } else {
          float x0, x1;
          unpk2(x, x0, x1);
          p0 = ex2(x0);
          p1 = ex2(x1);
        }
        sum2 = add2(sum2, pk2(p0, p1));
        if (k & 2) acc ^= pack2(p0, p1); else acc += pack2(p0, p1);
        s[i + k] = p0;
        s[i + k + 1] = p1;
      }
    }
  }
  const long long t1 = clock64();
  float a, b;
  unpk2(sum2, a, b);
  float r = a + b + __uint_as_float(acc & 0x3fffffffu);
#pragma unroll
  for (int i = 0; i < 64; ++i) r += s[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
  if ((threadIdx.x & 31) == 0) clk[blockIdx.x * 8 + (threadIdx.x >> 5)] = t1 - t0;
}

template <int POLY16>
void run2(const char* name, const float* in, float* out, long long* clk, int warps) {
  const int iters = 2000;
  long long h[8];
  bench2<POLY16><<<1, warps * 32>>>(in, out, clk, 10, 0.18f, -0.5f);
  bench2<POLY16><<<1, warps * 32>>>(in, out, clk, iters, 0.18f, -0.5f);
  cudaDeviceSynchronize();
  cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (int w = 0; w < warps; ++w) mx = h[w] > mx ? h[w] : mx;
  printf("%-44s warps/SM=%d (%d per SMSP): %7.1f clk per 64-element chunk per warp\n", name, warps, warps / 4,
         (double)mx / iters);
}

template <int MODE>
__global__ void __launch_bounds__(256, 1) bench(const float* __restrict__ in, float* __restrict__ out,
                                               long long* __restrict__ clk, int iters, float scale, float negm) {
  float s[64];
#pragma unroll
  for (int i = 0; i < 64; ++i) s[i] = in[(threadIdx.x * 64 + i) & 4095];
  float sum0 = 0.0f, sum1 = 0.0f;
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 64; i += 4) {
      float p0, p1, p2, p3;
      if constexpr (MODE == 0) {            // MUFU only
        p0 = ex2(s[i]); p1 = ex2(s[i + 1]); p2 = ex2(s[i + 2]); p3 = ex2(s[i + 3]);
      } else if constexpr (MODE <= 3) {     // FFMA + MUFU (+ FADD, + pack below)
        p0 = ex2(fmaf(s[i], scale, negm)); p1 = ex2(fmaf(s[i + 1], scale, negm));
        p2 = ex2(fmaf(s[i + 2], scale, negm)); p3 = ex2(fmaf(s[i + 3], scale, negm));
      } else if constexpr (MODE == 4) {     // 1 of 4 on the FMA pipe
        p0 = ex2(fmaf(s[i], scale, negm)); p1 = ex2(fmaf(s[i + 1], scale, negm));
        p2 = ex2(fmaf(s[i + 2], scale, negm)); p3 = ex2_poly(fmaf(s[i + 3], scale, negm));
      } else if constexpr (MODE == 5) {     // 2 of 4 on the FMA pipe
        p0 = ex2(fmaf(s[i], scale, negm)); p1 = ex2_poly(fmaf(s[i + 1], scale, negm));
        p2 = ex2(fmaf(s[i + 2], scale, negm)); p3 = ex2_poly(fmaf(s[i + 3], scale, negm));
      } else {                              // MODE 6: all on the FMA pipe
        p0 = ex2_poly(fmaf(s[i], scale, negm)); p1 = ex2_poly(fmaf(s[i + 1], scale, negm));
        p2 = ex2_poly(fmaf(s[i + 2], scale, negm)); p3 = ex2_poly(fmaf(s[i + 3], scale, negm));
      }
      if constexpr (MODE >= 2) {
        sum0 += p0 + p1;
        sum1 += p2 + p3;
      }
      if constexpr (MODE >= 3) {
        acc ^= pack2(p0, p1);
        acc += pack2(p2, p3);
      }
      s[i] = p0; s[i + 1] = p1; s[i + 2] = p2; s[i + 3] = p3;
    }
  }
  const long long t1 = clock64();
  float r = sum0 + sum1 + __uint_as_float(acc & 0x3fffffffu);
#pragma unroll
  for (int i = 0; i < 64; ++i) r += s[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
  if ((threadIdx.x & 31) == 0) clk[blockIdx.x * 8 + (threadIdx.x >> 5)] = t1 - t0;
}

template <int MODE>
void run(const char* name, const float* in, float* out, long long* clk, int warps) {
  const int iters = 2000;
  long long h[8];
  bench<MODE><<<1, warps * 32>>>(in, out, clk, 10, 0.18f, -0.5f);
  bench<MODE><<<1, warps * 32>>>(in, out, clk, iters, 0.18f, -0.5f);
  cudaDeviceSynchronize();
  cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (int w = 0; w < warps; ++w) mx = h[w] > mx ? h[w] : mx;
  printf("%-44s warps/SM=%d (%d per SMSP): %7.1f clk per 64-element chunk per warp\n", name, warps, warps / 4,
         (double)mx / iters);
}

int main() {
  float *in, *out;
  long long* clk;
  cudaMalloc(&in, 4096 * 4);
  cudaMalloc(&out, 256 * 4);
  cudaMalloc(&clk, 64 * 8);
  float h[4096];
  for (int i = 0; i < 4096; ++i) h[i] = -3.0f + 6.0f * (float)((i * 2654435761u) % 1000) / 1000.0f;
  cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
  for (int warps = 4; warps <= 8; warps += 4) {
    run<0>("MUFU.EX2 only", in, out, clk, warps);
    run<1>("FFMA + MUFU", in, out, clk, warps);
    run<2>("FFMA + MUFU + FADD (row sum)", in, out, clk, warps);
    run<3>("FFMA + MUFU + FADD + bf16 pack (as shipped)", in, out, clk, warps);
    run<4>("same, 1 of 4 exponentials on the FMA pipe", in, out, clk, warps);
    run<5>("same, 2 of 4 exponentials on the FMA pipe", in, out, clk, warps);
    run<6>("same, all exponentials on the FMA pipe", in, out, clk, warps);
    run2<0>("fp32x2 arithmetic, all MUFU", in, out, clk, warps);
    run2<2>("fp32x2, 2 of 16 on the FMA pipe", in, out, clk, warps);
    run2<4>("fp32x2, 4 of 16 on the FMA pipe", in, out, clk, warps);
    run2<6>("fp32x2, 6 of 16 on the FMA pipe", in, out, clk, warps);
    run2<8>("fp32x2, 8 of 16 on the FMA pipe", in, out, clk, warps);
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
  return 0;
}
