// kNN evaluation top-k (second half): eval_knn, models/contrastive.py:232-241
//     dist = einsum("nc,mc->nm", q, bank);  yd, yi = dist.topk(knn_k, dim=1, largest=True, sorted=True)
// The similarity matrix comes from the head's own tcgen05 mainloop (avssl_moco_infonce_sweep with the bank in the
// queue's place, T = 1, writes logits[:, 1:]); this file is the top-k behind it: every row of `dist` is read ONCE,
// and the result is exact (no approximation, ties broken towards the smaller index, deterministic).
//
//   pass 1  grid (G segments, N rows): the CTA stages its segment of the row in shared memory as order-preserving
//           32-bit keys, finds the segment's k-th largest key by a 4 x 8-bit radix select on shared-memory histograms,
//           collects the keys above it plus the needed ties (smallest indices first) and sorts those k candidates
//           (bitonic, 64-bit composites key << 32 | ~index).
//   pass 2  grid (N rows): bitonic sort of the row's G x KP candidates, the first k are the answer; values are
//           rescaled by ||q_row|| when the similarities were computed on normalised queries.
#include "common.cuh"

namespace avssl {
namespace {

constexpr int kKnnThreads = 256;
constexpr int kKnnMaxKP = 1024;         // k is padded to a power of two KP <= 1024
constexpr int kKnnMergeMax = 8192;      // pass 2 sorts at most this many candidates per row (64 KB of composites)
constexpr int kKnnSegMin = 16384;       // elements per segment (64 KB of keys) unless the merge limit asks for more
constexpr int kKnnSegMaxBytes = 192 * 1024;

__device__ __forceinline__ uint32_t order_key(float f) {  // larger float <-> larger key
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_value(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}
__device__ __forceinline__ unsigned long long composite(uint32_t key, uint32_t idx) {
  return ((unsigned long long)key << 32) | (unsigned long long)(0xffffffffu - idx);  // key desc, then index asc
}

// Descending bitonic sort of n (power of two) composites in shared memory by the whole CTA.
__device__ void bitonic_desc(unsigned long long* s, int n) {
  for (int size = 2; size <= n; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = threadIdx.x; i < (n >> 1); i += blockDim.x) {
        const int pos = 2 * i - (i & (stride - 1));
        const unsigned long long a = s[pos], b = s[pos + stride];
        const bool desc = (pos & size) == 0;
        if (desc ? (a < b) : (a > b)) {
          s[pos] = b;
          s[pos + stride] = a;
        }
      }
      __syncthreads();
    }
  }
}

struct KnnArgs {
  const float* dist;  // [N, ld], the row's M similarities start at column 0
  int64_t ld;
  int N, M, k, KP, G, seg;
  unsigned long long* cand;  // [N, G, KP]
  const float* q;            // [N, D] or null: yd *= ||q_row||
  int D;
  float* yd;       // [N, k]
  long long* yi;   // [N, k]
};

__global__ void __launch_bounds__(kKnnThreads) knn_segment_topk_kernel(const KnnArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned long long* cand_s = reinterpret_cast<unsigned long long*>(smem_raw);       // [KP]
  uint32_t* keys = reinterpret_cast<uint32_t*>(cand_s + a.KP);                        // [seg]
  __shared__ unsigned hist[256];
  __shared__ unsigned s_bin, s_need, s_cnt, s_run;
  __shared__ unsigned s_warp[kKnnThreads / 32];

  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int g = blockIdx.x, row = blockIdx.y;
  const int base = g * a.seg;
  const int L = min(a.seg, a.M - base);  // >= 1 by construction of G
  const float* src = a.dist + (int64_t)row * a.ld + base;
  for (int i = t; i < L; i += kKnnThreads) keys[i] = order_key(src[i]);
  for (int i = t; i < a.KP; i += kKnnThreads) cand_s[i] = 0ull;  // below every real candidate
  if (t == 0) {
    s_cnt = 0;
    s_run = 0;
  }
  __syncthreads();

  const int k = min(a.k, L);  // a short segment contributes all its elements
  uint32_t kth = 0;           // k-th largest key of the segment
  unsigned need = (unsigned)k;
  if (L > k) {
    uint32_t prefix = 0, mask = 0;
    for (int shift = 24; shift >= 0; shift -= 8) {
      hist[t] = 0;  // kKnnThreads == 256 bins
      __syncthreads();
      for (int i = t; i < L; i += kKnnThreads) {
        const uint32_t key = keys[i];
        if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1u);
      }
      __syncthreads();
      if (warp == 0) {
        // lane l owns bins 255-8l .. 248-8l (highest first); find the bin where the count from the top reaches `need`
        unsigned c[8], tot = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          c[j] = hist[255 - 8 * lane - j];
          tot += c[j];
        }
        unsigned incl = tot;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const unsigned n = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += n;
        }
        const unsigned excl = incl - tot;
        if (excl < need && need <= incl) {
          unsigned cum = excl;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            if (cum + c[j] >= need) {
              s_bin = 255u - 8u * lane - j;
              s_need = need - cum;
              break;
            }
            cum += c[j];
          }
        }
      }
      __syncthreads();
      prefix |= s_bin << shift;
      mask |= 255u << shift;
      need = s_need;
      __syncthreads();  // s_bin / s_need / hist are rewritten by the next pass
    }
    kth = prefix;
  }
  // collect: everything above kth (any order), then `need` elements equal to kth, smallest indices first
  if (L > k) {
    for (int i = t; i < L; i += kKnnThreads) {
      const uint32_t key = keys[i];
      if (key > kth) cand_s[atomicAdd(&s_cnt, 1u)] = composite(key, (uint32_t)(base + i));
    }
    __syncthreads();
    const unsigned n_gt = s_cnt;  // == k - need
    for (int i0 = 0; i0 < L; i0 += kKnnThreads) {
      const int i = i0 + t;
      const bool eq = i < L && keys[i] == kth;
      const unsigned bal = __ballot_sync(0xffffffffu, eq);
      if (lane == 0) s_warp[warp] = __popc(bal);
      __syncthreads();
      unsigned before = s_run, total = 0;
#pragma unroll
      for (int w = 0; w < kKnnThreads / 32; ++w) {
        const unsigned cw = s_warp[w];
        if (w < warp) before += cw;
        total += cw;
      }
      const unsigned rank = before + __popc(bal & ((1u << lane) - 1u));
      if (eq && rank < need) cand_s[n_gt + rank] = composite(kth, (uint32_t)(base + i));
      __syncthreads();
      if (t == 0) s_run += total;
      __syncthreads();
      if (s_run >= need) break;  // uniform: s_run is shared
    }
  } else {
    for (int i = t; i < L; i += kKnnThreads) cand_s[i] = composite(keys[i], (uint32_t)(base + i));
  }
  __syncthreads();
  bitonic_desc(cand_s, a.KP);
  unsigned long long* out = a.cand + ((int64_t)row * a.G + g) * a.KP;
  for (int i = t; i < a.KP; i += kKnnThreads) out[i] = cand_s[i];
}

__global__ void __launch_bounds__(kKnnThreads) knn_merge_topk_kernel(const KnnArgs a, int P2) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned long long* s = reinterpret_cast<unsigned long long*>(smem_raw);  // [P2]
  __shared__ float s_scale;
  const int t = threadIdx.x, row = blockIdx.x;
  const int n = a.G * a.KP;
  const unsigned long long* in = a.cand + (int64_t)row * n;
  for (int i = t; i < P2; i += kKnnThreads) s[i] = i < n ? in[i] : 0ull;
  if (t < 32) {
    const float sc = a.q ? sqrtf(row_sumsq(a.q + (int64_t)row * a.D, a.D, t)) : 1.f;
    if (t == 0) s_scale = sc;
  }
  __syncthreads();
  if (a.G > 1) bitonic_desc(s, P2);  // one segment: already sorted
  const float scale = s_scale;
  for (int i = t; i < a.k; i += kKnnThreads) {
    const unsigned long long c = s[i];
    const float v = key_value((uint32_t)(c >> 32));
    a.yd[(int64_t)row * a.k + i] = a.q ? v * scale : v;
    a.yi[(int64_t)row * a.k + i] = (long long)(0xffffffffu - (uint32_t)(c & 0xffffffffu));
  }
}

// Segment plan shared by the workspace query and the launch. Returns false when the shape is unsupported.
bool knn_plan(int M, int k, int* KP, int* G, int* seg) {
  if (M < 1 || k < 1 || k > M || k > kKnnMaxKP) return false;
  int kp = 32;
  while (kp < k) kp <<= 1;
  const int gmax = kKnnMergeMax / kp;
  int g = (M + kKnnSegMin - 1) / kKnnSegMin;
  if (g > gmax) g = gmax;
  if (g < 1) g = 1;
  int sg = (M + g - 1) / g;
  g = (M + sg - 1) / sg;  // no empty segment
  if ((size_t)sg * 4 + (size_t)kp * 8 > (size_t)kKnnSegMaxBytes) return false;
  *KP = kp;
  *G = g;
  *seg = sg;
  return true;
}

}  // namespace
}  // namespace avssl

using namespace avssl;

extern "C" size_t avssl_topk_rows_workspace_bytes(int N, int M, int k) {
  int KP, G, seg;
  if (N < 1 || !knn_plan(M, k, &KP, &G, &seg)) return 0;
  return (size_t)N * G * KP * sizeof(unsigned long long);
}

extern "C" int avssl_topk_rows(const float* dist, int64_t ld, int N, int M, int k, const float* q_scale_rows, int D,
                               float* yd_out, int64_t* yi_out, void* workspace, size_t workspace_bytes, void* stream) {
  AVSSL_REQUIRE(dist && yd_out && yi_out && workspace, AVSSL_ERR_INVALID_ARGUMENT, "topk_rows: null pointer");
  AVSSL_REQUIRE(N >= 1 && N <= 65535 && M >= 1 && ld >= M && k >= 1 && k <= M, AVSSL_ERR_INVALID_ARGUMENT,
                "topk_rows: bad sizes N=%d M=%d k=%d", N, M, k);
  AVSSL_REQUIRE(!q_scale_rows || D > 0, AVSSL_ERR_INVALID_ARGUMENT, "topk_rows: D must be positive with q_scale_rows");
  KnnArgs a;
  AVSSL_REQUIRE(knn_plan(M, k, &a.KP, &a.G, &a.seg), AVSSL_ERR_UNSUPPORTED,
                "topk_rows: k=%d (<= %d) over M=%d does not fit the two-pass plan", k, kKnnMaxKP, M);
  AVSSL_REQUIRE(workspace_bytes >= avssl_topk_rows_workspace_bytes(N, M, k) && (reinterpret_cast<uintptr_t>(workspace) & 7u) == 0,
                AVSSL_ERR_WORKSPACE, "topk_rows: workspace too small or misaligned");
  a.dist = dist;
  a.ld = ld;
  a.N = N;
  a.M = M;
  a.k = k;
  a.cand = static_cast<unsigned long long*>(workspace);
  a.q = q_scale_rows;
  a.D = D;
  a.yd = yd_out;
  a.yi = reinterpret_cast<long long*>(yi_out);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  static unsigned long long configured = 0;
  if (first_use_on_device(configured)) {
    AVSSL_CUDA_OK(cudaFuncSetAttribute(knn_segment_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       kKnnSegMaxBytes));
    AVSSL_CUDA_OK(cudaFuncSetAttribute(knn_merge_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)(kKnnMergeMax * sizeof(unsigned long long))));
  }
  const size_t smem1 = (size_t)a.KP * 8 + (size_t)a.seg * 4;
  knn_segment_topk_kernel<<<dim3(a.G, N), kKnnThreads, smem1, s>>>(a);
  AVSSL_LAUNCH_OK("knn_segment_topk_kernel");
  int P2 = a.KP;
  while (P2 < a.G * a.KP) P2 <<= 1;
  knn_merge_topk_kernel<<<N, kKnnThreads, (size_t)P2 * 8, s>>>(a, P2);
  AVSSL_LAUNCH_OK("knn_merge_topk_kernel");
  return AVSSL_OK;
}
