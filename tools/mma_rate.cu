// Micro-benchmark: issue rate of tcgen05.mma (SS, bf16, N=256, K=16) with NO loads in flight.
//   mode 1: cta_group::1, M=128 (one CTA per SM)
//   mode 2: cta_group::2, M=256 (CTA pairs), optional concurrent TMA-like smem fill disabled
// Prints cycles per MMA as seen by the issuing thread (clock64 around issue + final commit wait).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I distllm_b200/csrc -o /tmp/mma_rate tools/mma_rate.cu
#include <cstdio>
#include <cuda_runtime.h>

#include "common.cuh"

using namespace b2e;

constexpr int STAGE = 32768 + 16384;  // A 16K + B up to 32K

template <int PAIR>
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(long long* out, int iters) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t pad = ((raw + 1023u) & ~1023u) - raw;
  const uint32_t sb = raw + pad;
  uint8_t* smem = smem_raw + pad;
  for (int i = threadIdx.x; i < 4 * STAGE / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  const uint32_t bar = sb + 4 * STAGE;
  volatile uint32_t* slot = reinterpret_cast<volatile uint32_t*>(smem + 4 * STAGE + 64);
  const int warp = threadIdx.x >> 5;
  if (warp == 1 && elect_one()) {
    mbar_init(bar, 1);
    mbar_fence_init();
  }
  if (PAIR) cluster_sync_all();
  if (warp == 2) {
    if (PAIR) tmem_alloc_pair(smem_u32(const_cast<uint32_t*>(slot)), 512);
    else tmem_alloc(smem_u32(const_cast<uint32_t*>(slot)), 512);
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  const bool leader = !PAIR || cluster_ctarank() == 0;
  if (warp == 1 && leader && elect_one()) {
    constexpr uint32_t idesc = make_idesc_bf16(PAIR ? 256 : 128, 256, 0, 0);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const uint32_t a_addr = sb + (it & 3) * STAGE;
      const uint64_t a_desc = make_smem_desc_sw128(a_addr, 16, 1024);
      const uint64_t b_desc = make_smem_desc_sw128(a_addr + 16384, 16, 1024);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (PAIR) tc_mma_f16_ss_pair(tmem, a_desc + 2u * k, b_desc + 2u * k, idesc, 1u);
        else tc_mma_f16_ss(tmem, a_desc + 2u * k, b_desc + 2u * k, idesc, 1u);
      }
    }
    const long long t1 = clock64();
    if (PAIR) tc_commit_pair(bar, 1);
    else tc_commit(bar);
    mbar_wait(bar, 0);
    const long long t2 = clock64();
    if (blockIdx.x < 2) {
      out[blockIdx.x * 2] = t1 - t0;
      out[blockIdx.x * 2 + 1] = t2 - t0;
    }
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all();
  else __syncthreads();
  if (warp == 2) {
    if (PAIR) tmem_dealloc_pair(tmem, 512);
    else tmem_dealloc(tmem, 512);
  }
}

int main() {
  long long* d;
  cudaMalloc(&d, 64);
  const int iters = 2000;
  const int smem = 4 * STAGE + 2048;
  cudaFuncSetAttribute(mma_rate_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(mma_rate_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int grid : {1, 148}) {
    long long h[4] = {0, 0, 0, 0};
    cudaMemset(d, 0, 64);
    mma_rate_kernel<0><<<grid, 128, smem>>>(d, iters);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
    printf("cta_group::1 grid=%3d: %s issue %.1f clk/MMA, complete %.1f clk/MMA (128x256x16)\n", grid,
           cudaGetErrorString(e), h[0] / (4.0 * iters), h[1] / (4.0 * iters));
  }
  for (int grid : {2, 148}) {
    long long h[4] = {0, 0, 0, 0};
    cudaMemset(d, 0, 64);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(128);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, mma_rate_kernel<1>, d, iters);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
    printf("cta_group::2 grid=%3d: %s issue %.1f clk/MMA, complete %.1f clk/MMA (256x256x16 per pair)\n", grid,
           cudaGetErrorString(e), h[0] / (4.0 * iters), h[1] / (4.0 * iters));
  }
  return 0;
}
