// Multi-tensor L2 norm: replaces get_grad_norm_
// (models/optimizer.py:375-397, called every step from utils/solver.py:109-111) and the
// per-parameter torch.norm pairs of LARS.step (models/optimizer.py:351-352).
//
// Same flat chunk table as the momentum update (K1): one 256-thread CTA per four 4096-element
// chunks, sixteen independent 128-bit streaming loads per thread, fp32 sum of squares per chunk written to a
// partial array.  The fold rides in the same launch: the CTA that completes the LAST chunk of a tensor (a per-tensor
// arrival counter) sums that tensor's partials in chunk order, and the CTA that completes the last tensor adds the
// squared norms in tensor order -- every sum has a fixed shape, so the result is deterministic (no fp atomics), and
// there is no second launch behind the 144 MB stream (round 1: two launches, 0.59 of HBM).
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

// Persistent kernel: 2 CTAs per SM walk groups of kNormChunksPerCta chunks.  Per group: 16 independent 128-bit
// streaming loads per thread, a block reduction per chunk, the partials to memory -- and then ONE thread does the
// arrival bookkeeping (a counter per tensor; the chunk table carries the tensor index in flags[31:8]) while the
// other threads already issue the loads of the next group, so the bookkeeping latency hides under the stream.
// Whatever tensors the arrivals completed are folded (by the whole CTA, in chunk order) one iteration later.
__device__ __forceinline__ void norm_fold_tensors(const int* s_fold, const int* __restrict__ first_chunk, int n_tensors,
                                                  const float* __restrict__ partial, float* __restrict__ sq,
                                                  unsigned* __restrict__ tensor_done, unsigned* counter,
                                                  float* __restrict__ per_tensor, float* __restrict__ total, float* s_red,
                                                  int* s_flag, int extra_arrivals) {
  const int tid = threadIdx.x;
  int n_done = extra_arrivals;
  for (int k = 0; k < kNormChunksPerCta; ++k) {
    const int t = s_fold[k];
    if (t < 0) break;  // uniform over the CTA
    __threadfence();
    const int k0 = __ldg(first_chunk + t), k1 = __ldg(first_chunk + t + 1);
    float a = 0.f;
    for (int j = k0 + tid; j < k1; j += kNormThreads) a += __ldcg(partial + j);  // fixed thread <- chunk assignment
    __syncthreads();
    a = block_sum(a, s_red);
    if (tid == 0) {
      sq[t] = a;
      if (per_tensor) per_tensor[t] = sqrtf(a);
      tensor_done[t] = 0u;  // reusable: every chunk of t has arrived
    }
    ++n_done;
  }
  if (n_done == 0) return;
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    const unsigned before = atomicAdd(counter, (unsigned)n_done);
    *s_flag = (before + (unsigned)n_done == (unsigned)n_tensors) ? 1 : 0;
  }
  __syncthreads();
  if (*s_flag) {  // the last tensor is complete: norm(stack(norm_t)) = sqrt(sum_t norm_t^2), tensor order
    __threadfence();
    float a = 0.f;
    for (int t = tid; t < n_tensors; t += kNormThreads) a += __ldcg(sq + t);
    __syncthreads();
    a = block_sum(a, s_red);
    if (tid == 0) {
      *total = sqrtf(a);
      *counter = 0u;
    }
  }
}

__global__ void __launch_bounds__(kNormThreads, 2)
multi_l2norm_kernel(const avssl_ema_chunk* __restrict__ table, int n_chunks, const int* __restrict__ first_chunk,
                    int n_tensors, float* __restrict__ partial, float* __restrict__ sq, unsigned* __restrict__ tensor_done,
                    unsigned* counter, float* __restrict__ per_tensor, float* __restrict__ total) {
  __shared__ float s_red[8 * kNormChunksPerCta + 32];
  __shared__ int s_fold[2][kNormChunksPerCta];  // tensors completed by this CTA's arrivals of the previous iteration
  __shared__ int s_flag;
  const int tid = threadIdx.x;
  const int n_groups = (n_chunks + kNormChunksPerCta - 1) / kNormChunksPerCta;
  if (tid < 2 * kNormChunksPerCta) (&s_fold[0][0])[tid] = -1;

  // tensors without elements own no chunk and would never arrive: CTA 0 reports them (norm 0)
  int empties = 0;
  if (blockIdx.x == 0) {
    for (int t = tid; t < n_tensors; t += kNormThreads) {
      if (__ldg(first_chunk + t + 1) == __ldg(first_chunk + t)) {
        ++empties;
        sq[t] = 0.f;
        if (per_tensor) per_tensor[t] = 0.f;
      }
    }
    __syncthreads();
    empties = (int)(block_sum((float)empties, s_red) + 0.5f);
  }
  __syncthreads();

  int it = 0;
  for (int g = blockIdx.x; g < n_groups; g += gridDim.x, ++it) {
    const int c0 = g * kNormChunksPerCta;
    // ---- issue this group's loads
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
    float4 v[kNormChunksPerCta][4];
    if (fast) {
#pragma unroll
      for (int u = 0; u < kNormChunksPerCta; ++u) {
        const float4* x4 = reinterpret_cast<const float4*>(c[u].online) + tid;
#pragma unroll
        for (int k = 0; k < 4; ++k) v[u][k] = ldg_stream(x4 + k * kNormThreads);
      }
    }
    // ---- fold what the previous iteration's arrivals completed (its bookkeeping ran under the loads above)
    __syncthreads();
    norm_fold_tensors(s_fold[(it + 1) & 1], first_chunk, n_tensors, partial, sq, tensor_done, counter, per_tensor, total,
                      s_red + 8 * kNormChunksPerCta, &s_flag, 0);
    // ---- this group's sums of squares
    float ss[kNormChunksPerCta];
    if (fast) {
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
    __syncthreads();  // s_red of the fold above is dead
    norm_block_sum_n<kNormChunksPerCta>(ss, s_red);
    if (tid < kNormChunksPerCta && c0 + tid < n_chunks) {
      float out = ss[0];
#pragma unroll
      for (int u = 1; u < kNormChunksPerCta; ++u) out = tid == u ? ss[u] : out;
      partial[c0 + tid] = out;
    }
    __syncthreads();  // the partials of this group are written (ordered before thread 0's fence)
    if (tid == 0) {   // arrival bookkeeping; everybody else runs ahead into the next group's loads
      __threadfence();
      int n_fold = 0, u = 0;
      int* out = s_fold[it & 1];
      while (u < kNormChunksPerCta && c0 + u < n_chunks) {
        const int t = (int)(c[u].flags >> 8);
        int cnt = 1;
        while (u + cnt < kNormChunksPerCta && c0 + u + cnt < n_chunks && (int)(c[u + cnt].flags >> 8) == t) ++cnt;
        const unsigned len = (unsigned)(__ldg(first_chunk + t + 1) - __ldg(first_chunk + t));
        if (atomicAdd(tensor_done + t, (unsigned)cnt) + (unsigned)cnt == len) out[n_fold++] = t;
        u += cnt;
      }
      for (int k = n_fold; k < kNormChunksPerCta; ++k) out[k] = -1;
    }
  }
  // ---- what the last iteration completed (and CTA 0's empty tensors)
  __syncthreads();
  norm_fold_tensors(s_fold[(it + 1) & 1], first_chunk, n_tensors, partial, sq, tensor_done, counter, per_tensor, total,
                    s_red + 8 * kNormChunksPerCta, &s_flag, blockIdx.x == 0 ? empties : 0);
}

__global__ void multi_l2norm_empty_kernel(float* total) { *total = 0.f; }

}  // namespace avssl

using namespace avssl;

extern "C" size_t avssl_multi_l2norm_workspace_bytes(int64_t n_chunks, int n_tensors) {
  if (n_chunks < 0 || n_tensors < 0) return 0;
  // counter | partial[n_chunks] | sq[n_tensors] | tensor_done[n_tensors]
  return 256 + 4 * (size_t)n_chunks + 8 * (size_t)n_tensors + 256;
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
  const unsigned n_groups = (unsigned)((n_chunks + kNormChunksPerCta - 1) / kNormChunksPerCta);
  const unsigned cap = 2u * (unsigned)sm_count();  // persistent: two CTAs per SM (64 KiB in flight each)
  const unsigned grid = n_groups < cap ? n_groups : cap;
  float* sq = partial + n_chunks;
  unsigned* tensor_done = reinterpret_cast<unsigned*>(sq + n_tensors);  // zero-filled once with the workspace, self-resetting
  unsigned* counter = static_cast<unsigned*>(workspace);
  multi_l2norm_kernel<<<grid, kNormThreads, 0, s>>>(table_dev, (int)n_chunks, first_chunk_dev, n_tensors, partial, sq,
                                                    tensor_done, counter, per_tensor_norm_out, total_norm_out);
  AVSSL_LAUNCH_OK("multi_l2norm_kernel");
  return AVSSL_OK;
}
