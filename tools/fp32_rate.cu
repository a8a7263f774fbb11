// Issue rate of the FP32 instruction forms the softmax uses, one SM, 1 or 2 warps per SMSP:
// clk per warp instruction (64 independent chains per thread, so latency is hidden).
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/bin/fp32_rate tools/fp32_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long u64;

template <int OP>
__global__ void __launch_bounds__(256, 1) rate(const float* __restrict__ in, float* __restrict__ out,
                                              long long* __restrict__ clk, int iters, float a, float b) {
  float s[64];
#pragma unroll
  for (int i = 0; i < 64; ++i) s[i] = in[(threadIdx.x * 64 + i) & 4095];
  float breg = in[threadIdx.x & 4095] * 1e-3f + b;   // a per-thread register operand
  float areg = in[(threadIdx.x + 7) & 4095] * 1e-3f + a;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if constexpr (OP == 0) {          // FFMA, three register operands
#pragma unroll
      for (int i = 0; i < 64; ++i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(s[i]) : "f"(areg), "f"(breg));
    } else if constexpr (OP == 1) {   // FFMA, immediate addend
#pragma unroll
      for (int i = 0; i < 64; ++i) asm volatile("fma.rn.f32 %0, %0, %1, 0f3F000000;" : "+f"(s[i]) : "f"(areg));
    } else if constexpr (OP == 2) {   // FFMA, uniform (kernel parameter) multiplier, register addend
#pragma unroll
      for (int i = 0; i < 64; ++i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(s[i]) : "f"(a), "f"(breg));
    } else if constexpr (OP == 3) {   // FADD, two registers
#pragma unroll
      for (int i = 0; i < 64; ++i) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(s[i]) : "f"(breg));
    } else if constexpr (OP == 4) {   // FMUL
#pragma unroll
      for (int i = 0; i < 64; ++i) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(s[i]) : "f"(areg));
    } else if constexpr (OP == 5) {   // FFMA2 (32 instructions = 64 elements)
      u64 aa, bb;
      asm("mov.b64 %0, {%1, %1};" : "=l"(aa) : "f"(areg));
      asm("mov.b64 %0, {%1, %1};" : "=l"(bb) : "f"(breg));
#pragma unroll
      for (int i = 0; i < 64; i += 2) {
        u64 v;
        asm volatile("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(s[i]), "f"(s[i + 1]));
        asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(v) : "l"(aa), "l"(bb));
        asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(s[i]), "=f"(s[i + 1]) : "l"(v));
      }
    } else if constexpr (OP == 6) {   // FADD2
      u64 bb;
      asm("mov.b64 %0, {%1, %1};" : "=l"(bb) : "f"(breg));
#pragma unroll
      for (int i = 0; i < 64; i += 2) {
        u64 v;
        asm volatile("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(s[i]), "f"(s[i + 1]));
        asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(v) : "l"(bb));
        asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(s[i]), "=f"(s[i + 1]) : "l"(v));
      }
    } else if constexpr (OP == 7) {   // MUFU.EX2
#pragma unroll
      for (int i = 0; i < 64; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(s[i]));
    } else if constexpr (OP == 8) {   // FMNMX
#pragma unroll
      for (int i = 0; i < 64; ++i) asm volatile("max.f32 %0, %0, %1;" : "+f"(s[i]) : "f"(breg));
    } else if constexpr (OP == 9) {   // F2FP pack (two floats -> bf16x2), result folded back
#pragma unroll
      for (int i = 0; i < 64; i += 2) {
        uint32_t p;
        asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(p) : "f"(s[i + 1]), "f"(s[i]));
        s[i] = __uint_as_float(p);
      }
    } else if constexpr (OP == 10) {  // 3 MUFU : 2 FFMA2 interleaved (the x2 softmax mix without the row sum)
      u64 aa, bb;
      asm("mov.b64 %0, {%1, %1};" : "=l"(aa) : "f"(areg));
      asm("mov.b64 %0, {%1, %1};" : "=l"(bb) : "f"(breg));
#pragma unroll
      for (int i = 0; i < 64; i += 4) {
        u64 v;
        asm volatile("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(s[i]), "f"(s[i + 1]));
        asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(v) : "l"(aa), "l"(bb));
        asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(s[i]), "=f"(s[i + 1]) : "l"(v));
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(s[i + 2]));
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(s[i + 3]));
      }
    } else if constexpr (OP == 11) {  // 2 MUFU : 2 FFMA interleaved
#pragma unroll
      for (int i = 0; i < 64; i += 4) {
        asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(s[i]) : "f"(areg), "f"(breg));
        asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(s[i + 1]) : "f"(areg), "f"(breg));
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(s[i + 2]));
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(s[i + 3]));
      }
    }
  }
  const long long t1 = clock64();
  float r = 0.0f;
#pragma unroll
  for (int i = 0; i < 64; ++i) r += s[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
  if ((threadIdx.x & 31) == 0) clk[threadIdx.x >> 5] = t1 - t0;
}

template <int OP>
void run(const char* name, int per_iter, const float* in, float* out, long long* clk) {
  for (int warps = 4; warps <= 8; warps += 4) {
    const int iters = 4000;
    long long h[8];
    rate<OP><<<1, warps * 32>>>(in, out, clk, 10, 0.999f, 1e-4f);
    rate<OP><<<1, warps * 32>>>(in, out, clk, iters, 0.999f, 1e-4f);
    cudaDeviceSynchronize();
    cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int w = 0; w < warps; ++w) mx = h[w] > mx ? h[w] : mx;
    printf("%-46s %d warp(s)/SMSP: %6.2f clk per warp instruction per SMSP (%d instr/iter)\n", name, warps / 4,
           (double)mx / iters / per_iter / (warps / 4), per_iter);
  }
}

int main() {
  float *in, *out;
  long long* clk;
  cudaMalloc(&in, 4096 * 4);
  cudaMalloc(&out, 256 * 4);
  cudaMalloc(&clk, 64 * 8);
  float h[4096];
  for (int i = 0; i < 4096; ++i) h[i] = -1.0f + 2.0f * (float)((i * 2654435761u) % 1000) / 1000.0f;
  cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
  run<0>("FFMA r,r,r", 64, in, out, clk);
  run<1>("FFMA r,r,imm", 64, in, out, clk);
  run<2>("FFMA r,uniform,r", 64, in, out, clk);
  run<3>("FADD r,r", 64, in, out, clk);
  run<4>("FMUL r,r", 64, in, out, clk);
  run<5>("FFMA2 (two elements each)", 32, in, out, clk);
  run<6>("FADD2 (two elements each)", 32, in, out, clk);
  run<7>("MUFU.EX2", 64, in, out, clk);
  run<8>("FMNMX", 64, in, out, clk);
  run<9>("F2FP.BF16 pack", 32, in, out, clk);
  run<10>("mix: 1 FFMA2 + 2 MUFU (x16)", 48, in, out, clk);
  run<11>("mix: 2 FFMA + 2 MUFU (x16)", 64, in, out, clk);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
  return 0;
}
