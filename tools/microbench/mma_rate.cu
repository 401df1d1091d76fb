// Microbenchmark: issue-to-retire cost of back-to-back tcgen05.mma for several
// (kind, N, A source, accumulator pattern) combinations on one SM.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../advise_video_ssl_b200/csrc/sm100_ptx.cuh"
using namespace avssl::ptx;

__device__ __forceinline__ void mma_f16_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_f16_ts_local(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
               ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N, int bmn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)bmn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// mode: 0 tf32 SS, 1 tf32 TS, 2 bf16 SS, 3 bf16 TS;  dpat: 0 same D, 1 alternate 2 D tiles;  bmn: B MN-major
__global__ void bench(int mode, int N, int dpat, int bmn, int iters, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tbase;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  if (threadIdx.x < 32) tmem_alloc(&tbase, 512);
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = tbase;
  if (threadIdx.x == 0) {
    const uint32_t a_addr = smem_u32(smem), b_addr = smem_u32(smem + 64 * 1024);
    const bool tf32 = mode < 2, ts = (mode & 1);
    const uint32_t idesc = tf32 ? umma_idesc_tf32(128, N, 0, bmn) : idesc_bf16(128, N, bmn);
    const uint64_t ad = umma_smem_desc(a_addr, 16, 1024, kUmmaSwizzle128B);
    const uint64_t bd = bmn ? umma_smem_desc(b_addr, 8192, tf32 ? 512 : 1024, tf32 ? kUmmaSwizzle128BBase32B : kUmmaSwizzle128B)
                            : umma_smem_desc(b_addr, 16, 1024, kUmmaSwizzle128B);
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const uint32_t d = tmem + ((dpat && (i & 1)) ? 256 : 0);
      if (mode == 0) mma_tf32_ss(d, ad, bd, idesc, 1);
      else if (mode == 1) mma_tf32_ts(d, tmem + 480, bd, idesc, 1);
      else if (mode == 2) mma_f16_ss(d, ad, bd, idesc, 1);
      else mma_f16_ts_local(d, tmem + 480, bd, idesc, 1);
    }
    tc_commit(&bar);
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    out[0] = t1 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

int main() {
  long long* d; cudaMalloc(&d, 8);
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const char* names[] = {"tf32 SS", "tf32 TS", "bf16 SS", "bf16 TS"};
  const int iters = 2000;
  for (int mode = 0; mode < 4; ++mode)
    for (int bmn = 0; bmn < 2; ++bmn)
      for (int N : {64, 128, 256})
        for (int dpat = 0; dpat < 2; ++dpat) {
          if (dpat && N > 128) continue;
          bench<<<1, 128, 200 * 1024>>>(mode, N, dpat, bmn, iters, d);
          cudaError_t e = cudaDeviceSynchronize();
          long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
          printf("%s B-%s N=%3d %s: %7.1f clk/MMA  (%s)\n", names[mode], bmn ? "MN" : "K ", N, dpat ? "alt-D " : "same-D",
                 (double)h / iters, cudaGetErrorString(e));
        }
  return 0;
}
