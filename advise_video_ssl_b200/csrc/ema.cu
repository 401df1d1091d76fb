// K1 — multi-tensor momentum EMA (replaces models/contrastive.py:158-172).
//
// One launch updates every parameter tensor of the key encoder.  The work list is
// a device table of fixed-size chunks (pointer pair + length) so the grid is flat:
// one 256-thread CTA per 4096-element chunk, each thread keeping 8 independent
// 128-bit loads in flight (4 online + 4 history) before the first use.  Pure
// HBM-bound streaming: 12 bytes per parameter (read online, read hist, write hist).
//
// Bit-exactness: the reference evaluates `online * (1 - m) + hist * m` as three
// ATen kernels, i.e. three separately rounded fp32 operations.  __fmul_rn/__fadd_rn
// are never contracted into an FMA by nvcc, so every element matches bit for bit.
#include "common.cuh"
#include "peer.cuh"

namespace avssl {

constexpr int kEmaThreads = 256;
constexpr int kEmaUnroll = 4;
constexpr int kEmaChunk = kEmaThreads * 4 * kEmaUnroll;  // 4096 floats = 16 KiB per array

__device__ __forceinline__ float ema_blend(float o, float h, float m, float om) {
  return __fadd_rn(__fmul_rn(o, om), __fmul_rn(h, m));
}
__device__ __forceinline__ float4 ema_blend4(const float4& o, const float4& h, float m, float om) {
  return make_float4(ema_blend(o.x, h.x, m, om), ema_blend(o.y, h.y, m, om),
                     ema_blend(o.z, h.z, m, om), ema_blend(o.w, h.w, m, om));
}

// kFirstMode: 0 = not the first step, 1 = first step (host knows `iter`), 2 = read `iter` on
// the device.  kBump (only with a host-known `iter`): CTA 0 performs `self.iter += 1`
// (models/contrastive.py:314) -- nobody reads `iter` inside the kernel then, so there is no
// ordering to enforce.
//
// `chunk` is this CTA's entry of the table, `n_ctas` the number of CTAs working on the table.
template <int kFirstMode, bool kBump>
__device__ __forceinline__ void ema_cta(const avssl_ema_chunk* __restrict__ table, unsigned chunk, unsigned n_ctas,
                                        float m, float om, int64_t* iter, uint32_t* done_counter) {
  const avssl_ema_chunk c = table[chunk];
  // iter == 0: history := online first (models/contrastive.py:167-169)
  const bool first = kFirstMode == 2 ? (*reinterpret_cast<volatile int64_t*>(iter) == 0) : (kFirstMode == 1);
  const int tid = threadIdx.x;
  if (kBump && kFirstMode != 2 && chunk == 0 && tid == 0) *iter += 1;

  if ((c.flags & 1u) && c.n == (uint32_t)kEmaChunk) {
    const float4* o4 = reinterpret_cast<const float4*>(c.online) + tid;
    float4* h4 = reinterpret_cast<float4*>(c.hist) + tid;
    float4 o[kEmaUnroll], h[kEmaUnroll];
#pragma unroll
    for (int u = 0; u < kEmaUnroll; ++u) o[u] = ldg_stream(o4 + u * kEmaThreads);
    if (!first) {
#pragma unroll
      for (int u = 0; u < kEmaUnroll; ++u) h[u] = ld_stream(h4 + u * kEmaThreads);
    } else {
#pragma unroll
      for (int u = 0; u < kEmaUnroll; ++u) h[u] = o[u];
    }
#pragma unroll
    for (int u = 0; u < kEmaUnroll; ++u) st_stream(h4 + u * kEmaThreads, ema_blend4(o[u], h[u], m, om));
  } else if (c.flags & 1u) {
    // aligned tail chunk: vector body + scalar remainder
    const uint32_t n4 = c.n >> 2;
    const float4* o4 = reinterpret_cast<const float4*>(c.online);
    float4* h4 = reinterpret_cast<float4*>(c.hist);
    for (uint32_t i = tid; i < n4; i += kEmaThreads) {
      const float4 o = ldg_stream(o4 + i);
      const float4 h = first ? o : ld_stream(h4 + i);
      st_stream(h4 + i, ema_blend4(o, h, m, om));
    }
    for (uint32_t i = (n4 << 2) + tid; i < c.n; i += kEmaThreads) {
      const float o = c.online[i];
      const float h = first ? o : c.hist[i];
      c.hist[i] = ema_blend(o, h, m, om);
    }
  } else {
    for (uint32_t i = tid; i < c.n; i += kEmaThreads) {
      const float o = c.online[i];
      const float h = first ? o : c.hist[i];
      c.hist[i] = ema_blend(o, h, m, om);
    }
  }

  if (kBump && kFirstMode == 2) {
    // device-read mode: `self.iter += 1` by the last CTA to finish (every CTA has read `iter`
    // before it arrives here)
    __syncthreads();
    if (tid == 0) {
      __threadfence();
      const uint32_t prev = atomicAdd(done_counter, 1u);
      if (prev == n_ctas - 1) {
        *iter = *reinterpret_cast<volatile int64_t*>(iter) + 1;
        *done_counter = 0u;
        __threadfence();
      }
    }
  }
}

template <int kFirstMode, bool kBump>
__global__ void __launch_bounds__(kEmaThreads)
ema_multi_tensor_kernel(const avssl_ema_chunk* __restrict__ table, float m, float om,
                        int64_t* iter, uint32_t* done_counter) {
  ema_cta<kFirstMode, kBump>(table, blockIdx.x, gridDim.x, m, om, iter, done_counter);
}

// K1 + C3 push in one launch: the first `world` CTAs store this rank's key rows into every peer's
// exchange buffer over NVLink (peer.cuh) while the others stream the parameters; the transfer
// rides under the 70 us of EMA traffic and costs no launch of its own.
template <int kFirstMode, bool kBump>
__global__ void __launch_bounds__(kEmaThreads)
ema_multi_tensor_push_kernel(const avssl_ema_chunk* __restrict__ table, float m, float om, int64_t* iter,
                             uint32_t* done_counter, const avssl_peer_xchg x, const float* __restrict__ rows) {
  __shared__ unsigned long long s_epoch;
  if (blockIdx.x < (unsigned)x.world) {
    peer_push_cta<false>(x, rows, blockIdx.x, &s_epoch);
    return;
  }
  ema_cta<kFirstMode, kBump>(table, blockIdx.x - x.world, gridDim.x - x.world, m, om, iter, done_counter);
}

__global__ void bump_iter_kernel(int64_t* iter) { *iter += 1; }

}  // namespace avssl

using namespace avssl;

extern "C" int64_t avssl_ema_chunk_elems(void) { return kEmaChunk; }

extern "C" int64_t avssl_ema_plan_chunks(const int64_t* numel_host, int n_tensors) {
  if (!numel_host || n_tensors < 0) return -1;
  int64_t total = 0;
  for (int t = 0; t < n_tensors; ++t) {
    if (numel_host[t] < 0) return -1;
    total += (numel_host[t] + kEmaChunk - 1) / kEmaChunk;
  }
  return total;
}

extern "C" int avssl_ema_plan_fill(const uint64_t* online_ptrs_host, const uint64_t* hist_ptrs_host,
                                   const int64_t* numel_host, int n_tensors,
                                   avssl_ema_chunk* table_host, int64_t n_chunks) {
  AVSSL_REQUIRE(online_ptrs_host && hist_ptrs_host && numel_host && (table_host || n_chunks == 0),
                AVSSL_ERR_INVALID_ARGUMENT, "ema_plan_fill: null argument");
  AVSSL_REQUIRE(avssl_ema_plan_chunks(numel_host, n_tensors) == n_chunks, AVSSL_ERR_INVALID_ARGUMENT,
                "ema_plan_fill: table has %lld entries, plan needs %lld", (long long)n_chunks,
                (long long)avssl_ema_plan_chunks(numel_host, n_tensors));
  int64_t k = 0;
  for (int t = 0; t < n_tensors; ++t) {
    const uint64_t po = online_ptrs_host[t], ph = hist_ptrs_host[t];
    AVSSL_REQUIRE(numel_host[t] == 0 || (po && ph), AVSSL_ERR_INVALID_ARGUMENT,
                  "ema_plan_fill: tensor %d has a null pointer", t);
    AVSSL_REQUIRE((po & 3u) == 0 && (ph & 3u) == 0, AVSSL_ERR_INVALID_ARGUMENT,
                  "ema_plan_fill: tensor %d is not 4-byte aligned", t);
    for (int64_t off = 0; off < numel_host[t]; off += kEmaChunk) {
      avssl_ema_chunk& c = table_host[k++];
      const int64_t n = numel_host[t] - off < kEmaChunk ? numel_host[t] - off : kEmaChunk;
      c.online = reinterpret_cast<const float*>(po + 4ull * off);
      c.hist = reinterpret_cast<float*>(ph + 4ull * off);
      c.n = (uint32_t)n;
      c.flags = (((po + 4ull * off) | (ph + 4ull * off)) & 15u) == 0 ? 1u : 0u;
    }
  }
  return AVSSL_OK;
}

namespace {

int ema_run(const avssl_ema_chunk* table_dev, int64_t n_chunks, float m, float one_minus_m, int64_t* iter_dev,
            int first_iter, int bump_iter, uint32_t* done_counter_dev, const avssl_peer_xchg* x, const float* rows,
            void* stream) {
  AVSSL_REQUIRE(iter_dev, AVSSL_ERR_INVALID_ARGUMENT, "ema_multi_tensor: iter_dev is null");
  AVSSL_REQUIRE(first_iter >= -1 && first_iter <= 1, AVSSL_ERR_INVALID_ARGUMENT, "ema_multi_tensor: first_iter must be -1, 0 or 1");
  AVSSL_REQUIRE(n_chunks >= 0 && n_chunks < (1ll << 31), AVSSL_ERR_INVALID_ARGUMENT,
                "ema_multi_tensor: bad n_chunks %lld", (long long)n_chunks);
  AVSSL_REQUIRE(!(bump_iter && first_iter < 0) || done_counter_dev, AVSSL_ERR_INVALID_ARGUMENT,
                "ema_multi_tensor: bump_iter with a device-read iter needs done_counter_dev");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (n_chunks == 0) {
    if (bump_iter) {
      bump_iter_kernel<<<1, 1, 0, s>>>(iter_dev);
      AVSSL_LAUNCH_OK("bump_iter_kernel");
    }
    return x ? avssl_peer_push_rows(x, rows, stream) : AVSSL_OK;
  }
  AVSSL_REQUIRE(table_dev, AVSSL_ERR_INVALID_ARGUMENT, "ema_multi_tensor: table_dev is null");
  const unsigned grid = (unsigned)n_chunks + (x ? (unsigned)x->world : 0u);
#define AVSSL_EMA_LAUNCH(MODE, BUMP)                                                                              \
  do {                                                                                                            \
    if (x)                                                                                                        \
      ema_multi_tensor_push_kernel<MODE, BUMP><<<grid, kEmaThreads, 0, s>>>(table_dev, m, one_minus_m, iter_dev, \
                                                                             done_counter_dev, *x, rows);         \
    else                                                                                                          \
      ema_multi_tensor_kernel<MODE, BUMP><<<grid, kEmaThreads, 0, s>>>(table_dev, m, one_minus_m, iter_dev,      \
                                                                        done_counter_dev);                        \
  } while (0)
  if (first_iter < 0) {
    if (bump_iter) AVSSL_EMA_LAUNCH(2, true); else AVSSL_EMA_LAUNCH(2, false);
  } else if (first_iter == 1) {
    if (bump_iter) AVSSL_EMA_LAUNCH(1, true); else AVSSL_EMA_LAUNCH(1, false);
  } else {
    if (bump_iter) AVSSL_EMA_LAUNCH(0, true); else AVSSL_EMA_LAUNCH(0, false);
  }
#undef AVSSL_EMA_LAUNCH
  AVSSL_LAUNCH_OK("ema_multi_tensor_kernel");
  return AVSSL_OK;
}

}  // namespace

extern "C" int avssl_ema_multi_tensor(const avssl_ema_chunk* table_dev, int64_t n_chunks, float m,
                                      float one_minus_m, int64_t* iter_dev, int first_iter, int bump_iter,
                                      uint32_t* done_counter_dev, void* stream) {
  return ema_run(table_dev, n_chunks, m, one_minus_m, iter_dev, first_iter, bump_iter, done_counter_dev, nullptr,
                 nullptr, stream);
}

extern "C" int avssl_ema_multi_tensor_push(const avssl_ema_chunk* table_dev, int64_t n_chunks, float m,
                                           float one_minus_m, int64_t* iter_dev, int first_iter, int bump_iter,
                                           uint32_t* done_counter_dev, const avssl_peer_xchg* x, const float* rows,
                                           void* stream) {
  int rc = peer_check(x, "ema_multi_tensor_push");
  if (rc != AVSSL_OK) return rc;
  AVSSL_REQUIRE(rows && (reinterpret_cast<uintptr_t>(rows) & 15u) == 0, AVSSL_ERR_INVALID_ARGUMENT,
                "ema_multi_tensor_push: rows is null or not 16-byte aligned");
  return ema_run(table_dev, n_chunks, m, one_minus_m, iter_dev, first_iter, bump_iter, done_counter_dev, x, rows,
                 stream);
}
