// Micro-benchmark: tcgen05.mma issue rate WHILE TMA keeps filling the same shared-memory ring
// (unsynchronised: results are garbage, rates are real).  Separates "the tensor core is slow" from
// "the shared-memory fill is slow" for the single-CTA (128x256) and CTA-pair (256x256) tile shapes.
//   per CTA and K block of 64:  single = A 16 KiB + W 32 KiB ; pair = A 16 KiB + half W 16 KiB
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I distllm_b200/csrc \
//             -o tools/bin/mma_fill_rate tools/mma_fill_rate.cu
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>

#include "common.cuh"

using namespace b2e;

constexpr int NSTAGE = 4;

template <int PAIR>
__global__ void __launch_bounds__(128, 1)
fill_rate_kernel(const __grid_constant__ CUtensorMap tm, const __grid_constant__ CUtensorMap tm_a,
                 const __grid_constant__ CUtensorMap tm_w, long long* out, int iters, int do_mma,
                 int do_fill, int M, int N, int K) {
  constexpr int STAGE = PAIR ? 32768 : 49152;
  constexpr int BOXES = STAGE / 16384;  // 16 KiB TMA boxes (64 cols x 128 rows) per stage
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sb = smem_u32(smem);
  for (int i = threadIdx.x; i < NSTAGE * STAGE / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  const uint32_t bars = sb + NSTAGE * STAGE;  // NSTAGE fill barriers + 1 mma barrier
  volatile uint32_t* slot = reinterpret_cast<volatile uint32_t*>(smem + NSTAGE * STAGE + 128);
  const int warp = threadIdx.x >> 5;
  if (warp == 1 && elect_one()) {
    for (int i = 0; i <= NSTAGE; ++i) mbar_init(bars + 8u * i, 1);
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
  if (warp == 0 && do_fill && elect_one()) {
    // free-running loader: stage s is reloaded as soon as its previous load has landed
    const long long t0 = clock64();
    uint32_t phase = 0;
    int stage = 0;
    const int rows_total = 1 << 20;
    int row = (blockIdx.x * 4096) % rows_total;
    for (int it = 0; it < iters; ++it) {
      if (it >= NSTAGE) mbar_wait(bars + 8u * stage, phase ^ 1u);
      mbar_expect_tx(bars + 8u * stage, STAGE);
      if (do_fill == 1) {
        for (int b = 0; b < BOXES; ++b) {
          tma_load_2d(sb + stage * STAGE + b * 16384, &tm, bars + 8u * stage, 0, row);
          row = (row + 128) % rows_total;
        }
      } else {
        // GEMM-like pattern: tile = unit + i*units, K blocks innermost
        const int kblocks = K / 64;
        const int n_tiles = N / 256;
        const int unit = PAIR ? (blockIdx.x >> 1) : blockIdx.x;
        const int units = PAIR ? (gridDim.x >> 1) : gridDim.x;
        const int m_rows = PAIR ? 256 : 128;
        const int total = (M / m_rows) * n_tiles;
        const int tile = (unit + (it / kblocks) * units) % total;
        const int kb = it % kblocks;
        const int m_blk = tile / n_tiles, n_blk = tile % n_tiles;
        const int rank = PAIR ? (blockIdx.x & 1) : 0;
        tma_load_2d(sb + stage * STAGE, &tm_a, bars + 8u * stage, kb * 64, m_blk * m_rows + rank * 128);
        if (PAIR) {
          tma_load_2d(sb + stage * STAGE + 16384, &tm_w, bars + 8u * stage, kb * 64, n_blk * 256 + rank * 128);
        } else {
          tma_load_2d(sb + stage * STAGE + 16384, &tm_w, bars + 8u * stage, kb * 64, n_blk * 256);
          tma_load_2d(sb + stage * STAGE + 32768, &tm_w, bars + 8u * stage, kb * 64, n_blk * 256 + 128);
        }
      }
      if (++stage == NSTAGE) { stage = 0; phase ^= 1u; }
    }
    for (int s = 0; s < NSTAGE; ++s) {  // drain
      const int st = (stage + s) % NSTAGE;
      const uint32_t ph = (st >= stage) ? (phase ^ 1u) : phase;
      mbar_wait(bars + 8u * st, ph);
    }
    if (blockIdx.x < 2) out[blockIdx.x * 4 + 2] = clock64() - t0;
  }
  if (warp == 1 && do_mma && leader && elect_one()) {
    constexpr uint32_t idesc = make_idesc_bf16(PAIR ? 256 : 128, 256, 0, 0);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const uint32_t a_addr = sb + (it % NSTAGE) * STAGE;
      const uint64_t a_desc = make_smem_desc_sw128(a_addr, 16, 1024);
      const uint64_t b_desc = make_smem_desc_sw128(a_addr + 16384, 16, 1024);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (PAIR) tc_mma_f16_ss_pair(tmem, a_desc + 2u * k, b_desc + 2u * k, idesc, 1u);
        else tc_mma_f16_ss(tmem, a_desc + 2u * k, b_desc + 2u * k, idesc, 1u);
      }
    }
    if (PAIR) tc_commit_pair(bars + 8u * NSTAGE, 1);
    else tc_commit(bars + 8u * NSTAGE);
    mbar_wait(bars + 8u * NSTAGE, 0);
    if (blockIdx.x < 2) out[blockIdx.x * 4 + 1] = clock64() - t0;
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all();
  else __syncthreads();
  if (warp == 2) {
    if (PAIR) tmem_dealloc_pair(tmem, 512);
    else tmem_dealloc(tmem, 512);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

template <int PAIR>
void run(const CUtensorMap& tm, const CUtensorMap& tm_a, const CUtensorMap& tm_w, long long* d, int grid,
         int iters, int do_mma, int do_fill, int M, int N, int K) {
  constexpr int STAGE = PAIR ? 32768 : 49152;
  const int smem = NSTAGE * STAGE + 1024;
  cudaFuncSetAttribute(fill_rate_kernel<PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaMemset(d, 0, 64);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(128);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = PAIR ? 2 : 1;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, fill_rate_kernel<PAIR>, tm, tm_a, tm_w, d, iters, do_mma, do_fill, M, N, K);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[8];
  cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
  printf("%-6s grid=%3d mma=%d fill=%d : %s | mma %.1f clk per K-block (ideal 512) | fill %.1f B/clk/SM (%.1f clk per K-block)\n",
         PAIR ? "pair" : "single", grid, do_mma, do_fill, cudaGetErrorString(e),
         do_mma ? h[1] / double(iters) : 0.0, do_fill && h[2] ? double(STAGE) * iters / h[2] : 0.0,
         do_fill ? h[2] / double(iters) : 0.0);
}

int main() {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  auto encode = reinterpret_cast<EncodeTiledFn>(p);
  const size_t rows = 1 << 20, cols = 64;  // 128 MiB: exceeds nothing much, mostly L2 resident
  void* g;
  cudaMalloc(&g, rows * cols * 2);
  cudaMemset(g, 0, rows * cols * 2);
  CUtensorMap tm;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * 2};
  cuuint32_t box[2] = {64, 128};
  cuuint32_t es[2] = {1, 1};
  encode(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, g, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  long long* d;
  cudaMalloc(&d, 64);
  const int iters = 2000;
  const int M = 262144, N = 2304;
  for (int K : {768, 3072}) {
    void *ga, *gw;
    cudaMalloc(&ga, size_t(M) * K * 2);
    cudaMalloc(&gw, size_t(N) * K * 2);
    cudaMemset(ga, 0, size_t(M) * K * 2);
    cudaMemset(gw, 0, size_t(N) * K * 2);
    CUtensorMap tm_a, tm_w;
    cuuint64_t da[2] = {cuuint64_t(K), cuuint64_t(M)}, dw[2] = {cuuint64_t(K), cuuint64_t(N)};
    cuuint64_t st[1] = {cuuint64_t(K) * 2};
    encode(&tm_a, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, ga, da, st, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    encode(&tm_w, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, gw, dw, st, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("--- GEMM-like access pattern, A [%d,%d], W [%d,%d]\n", M, K, N, K);
    for (int grid : {148}) {
      run<0>(tm, tm_a, tm_w, d, grid, iters, 0, 1, M, N, K);   // dense reference
      run<0>(tm, tm_a, tm_w, d, grid, iters, 0, 2, M, N, K);
      run<0>(tm, tm_a, tm_w, d, grid, iters, 1, 2, M, N, K);
      run<1>(tm, tm_a, tm_w, d, grid, iters, 0, 2, M, N, K);
      run<1>(tm, tm_a, tm_w, d, grid, iters, 1, 2, M, N, K);
    }
    cudaFree(ga);
    cudaFree(gw);
  }
  return 0;
}
