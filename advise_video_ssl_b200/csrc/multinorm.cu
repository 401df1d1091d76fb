// Multi-tensor L2 norm: replaces get_grad_norm_
// (models/optimizer.py:375-397, called every step from utils/solver.py:109-111) and the
// per-parameter torch.norm pairs of LARS.step (models/optimizer.py:351-352).
//
// Same flat chunk table as the momentum update (K1): one 256-thread CTA per four 4096-element
// chunks, sixteen independent 128-bit streaming loads per thread, fp32 sum of squares per chunk written to a
// partial array; a second one-CTA launch folds the partials per tensor (chunk order) and the
// per-tensor norms into the total, all in a fixed order (deterministic, no fp atomics).
// HBM-bound: 4 bytes per element, read once.
#include "common.cuh"

namespace avssl {

constexpr int kNormThreads = 256;
constexpr int kNormChunk = 4096;  // == avssl_ema_chunk_elems()

constexpr int kNormChunksPerCta = 4;  // 64 KiB per CTA: amortises the CTA turn-over and the block reduction

template <int kN>
__device__ __forceinline__ void norm_block_sum_n(float (&v)[kN], float* scratch /* >= 8 * kN */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < kN; ++i) v[i] = warp_sum(v[i]);
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < kN; ++i) scratch[i * 8 + warp] = v[i];
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < kN; ++i) {
    float r = lane < kNormThreads / 32 ? scratch[i * 8 + lane] : 0.f;
    v[i] = warp_sum(r);
  }
}

__global__ void __launch_bounds__(kNormThreads)
multi_l2norm_kernel(const avssl_ema_chunk* __restrict__ table, int n_chunks, float* __restrict__ partial) {
  __shared__ float s_red[8 * kNormChunksPerCta];
  const int tid = threadIdx.x;
  const int c0 = blockIdx.x * kNormChunksPerCta;
  avssl_ema_chunk c[kNormChunksPerCta];
  bool fast = true;
#pragma unroll
  for (int u = 0; u < kNormChunksPerCta; ++u) {
    if (c0 + u < n_chunks) {
      c[u] = table[c0 + u];
    } else {
      c[u].online = nullptr;
      c[u].n = 0;
      c[u].flags = 1u;
    }
    fast = fast && (c[u].flags & 1u) && c[u].n == (uint32_t)kNormChunk;
  }
  float ss[kNormChunksPerCta];
  if (fast) {  // all loads of the CTA's chunks in flight before the first use
    float4 v[kNormChunksPerCta][4];
#pragma unroll
    for (int u = 0; u < kNormChunksPerCta; ++u) {
      const float4* x4 = reinterpret_cast<const float4*>(c[u].online) + tid;
#pragma unroll
      for (int k = 0; k < 4; ++k) v[u][k] = ldg_stream(x4 + k * kNormThreads);
    }
#pragma unroll
    for (int u = 0; u < kNormChunksPerCta; ++u) {
      float a[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) a[k] = v[u][k].x * v[u][k].x + v[u][k].y * v[u][k].y + v[u][k].z * v[u][k].z + v[u][k].w * v[u][k].w;
      ss[u] = (a[0] + a[1]) + (a[2] + a[3]);
    }
  } else {
#pragma unroll
    for (int u = 0; u < kNormChunksPerCta; ++u) {
      float a = 0.f;
      for (uint32_t i = tid; i < c[u].n; i += kNormThreads) {
        const float x = c[u].online[i];
        a = fmaf(x, x, a);
      }
      ss[u] = a;
    }
  }
  norm_block_sum_n<kNormChunksPerCta>(ss, s_red);
  if (tid < kNormChunksPerCta && c0 + tid < n_chunks) {
    float out = ss[0];
#pragma unroll
    for (int u = 1; u < kNormChunksPerCta; ++u) out = tid == u ? ss[u] : out;
    partial[c0 + tid] = out;
  }
}

// Second launch: one warp per tensor (lanes stride over its chunks in a fixed order) across
// ceil(n_tensors / 8) CTAs; the last CTA to finish adds the squared norms in tensor order.
constexpr int kFoldThreads = 256;
__global__ void __launch_bounds__(kFoldThreads)
multi_l2norm_fold_kernel(const int* __restrict__ first_chunk, int n_tensors, const float* __restrict__ partial,
                         float* __restrict__ sq, float* __restrict__ per_tensor, float* __restrict__ total,
                         unsigned* counter) {
  __shared__ float s_red[32];
  __shared__ unsigned s_last;
  const int tid = threadIdx.x, lane = tid & 31;
  const int t = blockIdx.x * (kFoldThreads / 32) + (tid >> 5);
  if (t < n_tensors) {
    const int k0 = __ldg(first_chunk + t), k1 = __ldg(first_chunk + t + 1);
    float s = 0.f;
    for (int k = k0 + lane; k < k1; k += 32) s += __ldcg(partial + k);
    s = warp_sum(s);
    if (lane == 0) {
      sq[t] = s;
      if (per_tensor) per_tensor[t] = sqrtf(s);
    }
  }
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    s_last = (atomicAdd(counter, 1u) == gridDim.x - 1) ? 1u : 0u;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  float a = 0.f;
  for (int k = tid; k < n_tensors; k += kFoldThreads) a += __ldcg(sq + k);
  a = block_sum(a, s_red);
  if (tid == 0) {
    *total = sqrtf(a);  // norm(stack(norm_t)) = sqrt(sum_t norm_t^2)
    *counter = 0u;
  }
}

__global__ void multi_l2norm_empty_kernel(float* total) { *total = 0.f; }

}  // namespace avssl

using namespace avssl;

extern "C" size_t avssl_multi_l2norm_workspace_bytes(int64_t n_chunks, int n_tensors) {
  if (n_chunks < 0 || n_tensors < 0) return 0;
  return 256 + 4 * (size_t)n_chunks + 4 * (size_t)n_tensors + 256;  // counter | partial[n_chunks] | sq[n_tensors]
}

extern "C" int avssl_multi_l2norm(const avssl_ema_chunk* table_dev, int64_t n_chunks, const int32_t* first_chunk_dev,
                                  int n_tensors, float* per_tensor_norm_out, float* total_norm_out, void* workspace,
                                  size_t workspace_bytes, void* stream) {
  AVSSL_REQUIRE(total_norm_out && workspace, AVSSL_ERR_INVALID_ARGUMENT, "multi_l2norm: null pointer");
  AVSSL_REQUIRE(n_chunks >= 0 && n_chunks < (1ll << 31) && n_tensors >= 0, AVSSL_ERR_INVALID_ARGUMENT,
                "multi_l2norm: bad sizes n_chunks=%lld n_tensors=%d", (long long)n_chunks, n_tensors);
  AVSSL_REQUIRE(workspace_bytes >= avssl_multi_l2norm_workspace_bytes(n_chunks, n_tensors), AVSSL_ERR_WORKSPACE,
                "multi_l2norm: workspace too small");
  AVSSL_REQUIRE(sm_count() > 0, AVSSL_ERR_CUDA, "multi_l2norm: no CUDA device (there is no CPU fallback)");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (n_chunks == 0) {  // get_grad_norm_ returns 0.0 for an empty list (models/optimizer.py:380-381)
    multi_l2norm_empty_kernel<<<1, 1, 0, s>>>(total_norm_out);
    AVSSL_LAUNCH_OK("multi_l2norm_empty_kernel");
    return AVSSL_OK;
  }
  AVSSL_REQUIRE(table_dev && first_chunk_dev && n_tensors > 0, AVSSL_ERR_INVALID_ARGUMENT, "multi_l2norm: null table");
  float* partial = reinterpret_cast<float*>(static_cast<char*>(workspace) + 256);
  const unsigned grid = (unsigned)((n_chunks + kNormChunksPerCta - 1) / kNormChunksPerCta);
  multi_l2norm_kernel<<<grid, kNormThreads, 0, s>>>(table_dev, (int)n_chunks, partial);
  AVSSL_LAUNCH_OK("multi_l2norm_kernel");
  float* sq = partial + n_chunks;
  unsigned* counter = static_cast<unsigned*>(workspace);
  multi_l2norm_fold_kernel<<<(n_tensors + kFoldThreads / 32 - 1) / (kFoldThreads / 32), kFoldThreads, 0, s>>>(
      first_chunk_dev, n_tensors, partial, sq, per_tensor_norm_out, total_norm_out, counter);
  AVSSL_LAUNCH_OK("multi_l2norm_fold_kernel");
  return AVSSL_OK;
}
