// K4 — MoCo queue ring write with a device-resident pointer
// (replaces models/contrastive.py:263-292; bit-exact copy, int64 pointer).
#include "common.cuh"

namespace avssl {

// Every CTA reads the old pointer, copies its share of rows, and the last CTA to
// finish advances the pointer (so no CTA can observe the new value).  n*D floats
// is tiny (32 KiB at B=64, D=128): the launch is latency-bound by construction.
__global__ void __launch_bounds__(1024)
queue_enqueue_kernel(float* __restrict__ queue, int64_t* ptr_dev, const float* __restrict__ keys, int n,
                     int K, int D, uint32_t* status) {
  __shared__ int64_t s_ptr;
  if (threadIdx.x == 0) s_ptr = *reinterpret_cast<volatile int64_t*>(ptr_dev);
  __syncthreads();
  const int64_t ptr = s_ptr;
  const bool ok = ptr >= 0 && ptr + n <= K;  // models/contrastive.py:285
  const int64_t total = (int64_t)n * D;
  if (ok) {
    float* dst = queue + ptr * D;
    if (((D & 3) == 0) && ((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(keys)) & 15u) == 0) {
      const int64_t t4 = total >> 2;
      for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < t4; i += (int64_t)gridDim.x * blockDim.x)
        reinterpret_cast<float4*>(dst)[i] = reinterpret_cast<const float4*>(keys)[i];
    } else {
      for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
        dst[i] = keys[i];
    }
  }
  // The grid is a single CTA unless n*D is large (> 1 MiB); with several CTAs the
  // pointer is advanced by a follow-up launch (see the host code below).
  if (gridDim.x == 1) {
    __syncthreads();
    if (threadIdx.x == 0) {
      if (ok) {
        int64_t np = ptr + n;
        if (np == K) np = 0;  // wrap only when landing exactly on K (:290-291)
        *ptr_dev = np;
      } else if (status) {
        atomicOr(status, AVSSL_DEVFLAG_QUEUE_OVERRUN);
      }
    }
  }
}

__global__ void queue_advance_kernel(int64_t* ptr_dev, int n, int K, uint32_t* status) {
  const int64_t ptr = *ptr_dev;
  if (ptr >= 0 && ptr + n <= K) {
    int64_t np = ptr + n;
    if (np == K) np = 0;
    *ptr_dev = np;
  } else if (status) {
    atomicOr(status, AVSSL_DEVFLAG_QUEUE_OVERRUN);
  }
}

}  // namespace avssl

using namespace avssl;

extern "C" int avssl_queue_enqueue(float* queue, int64_t* ptr_dev, const float* keys, int n, int K, int D,
                                   uint32_t* status_dev, void* stream) {
  AVSSL_REQUIRE(queue && ptr_dev && keys, AVSSL_ERR_INVALID_ARGUMENT, "queue_enqueue: null pointer");
  AVSSL_REQUIRE(n > 0 && K > 0 && D > 0, AVSSL_ERR_INVALID_ARGUMENT, "queue_enqueue: bad sizes n=%d K=%d D=%d", n, K, D);
  // models/contrastive.py:284  assert self.k % num_items == 0
  AVSSL_REQUIRE(K % n == 0, AVSSL_ERR_INVALID_ARGUMENT, "queue_enqueue: K=%d is not a multiple of n=%d", K, n);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int64_t total4 = ((int64_t)n * D + 3) / 4;
  int blocks = 1;
  if (total4 > 65536) {
    blocks = (int)((total4 + 4095) / 4096);
    if (blocks > 1184) blocks = 1184;
  }
  queue_enqueue_kernel<<<blocks, 1024, 0, s>>>(queue, ptr_dev, keys, n, K, D, status_dev);
  AVSSL_LAUNCH_OK("queue_enqueue_kernel");
  if (blocks > 1) {
    queue_advance_kernel<<<1, 1, 0, s>>>(ptr_dev, n, K, status_dev);
    AVSSL_LAUNCH_OK("queue_advance_kernel");
  }
  return AVSSL_OK;
}
