// Microbenchmark 2: does the ~120-cycle cost per tcgen05.mma overlap across issuing warps?
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../advise_video_ssl_b200/csrc/sm100_ptx.cuh"
using namespace avssl::ptx;

__global__ void bench(int n_issuers, int N, int iters, int unroll_desc, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar[4];
  __shared__ uint32_t tbase;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0;
  if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) mbar_init(&bar[i], 1); mbar_fence_init(); }
  if (threadIdx.x < 32) tmem_alloc(&tbase, 512);
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = tbase;
  const int w = threadIdx.x >> 5;
  long long t0 = clock64();
  if ((threadIdx.x & 31) == 0 && w < n_issuers) {
    const uint32_t b_addr = smem_u32(smem + 64 * 1024);
    const uint32_t idesc = umma_idesc_tf32(128, N, 0, 0);
    const uint32_t d = tmem + w * 64;
    uint64_t bd = umma_smem_desc(b_addr, 16, 1024, kUmmaSwizzle128B);
    for (int i = 0; i < iters; ++i) {
      if (unroll_desc) bd = umma_smem_desc(b_addr + (i & 3) * 32, 16, 1024, kUmmaSwizzle128B);
      mma_tf32_ts(d, tmem + 480, bd, idesc, 1);
    }
    tc_commit(&bar[w]);
    mbar_wait(&bar[w], 0);
  }
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0) out[0] = t1 - t0;
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

int main() {
  long long* d; cudaMalloc(&d, 8);
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int iters = 2000;
  for (int N : {16, 32, 64, 128})
    for (int W : {1, 2, 4}) {
      bench<<<1, 128, 200 * 1024>>>(W, N, iters, 0, d);
      cudaError_t e = cudaDeviceSynchronize();
      long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
      printf("tf32 TS N=%3d issuers=%d: %7.1f clk per MMA-per-issuer, %7.1f clk per MMA overall (%s)\n", N, W,
             (double)h / iters, (double)h / iters / W, cudaGetErrorString(e));
    }
  return 0;
}
