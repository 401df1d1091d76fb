// kNN evaluation top-k (second half): eval_knn, models/contrastive.py:232-241
//     dist = einsum("nc,mc->nm", q, bank);  yd, yi = dist.topk(knn_k, dim=1, largest=True, sorted=True)
// The similarity matrix comes from the head's own tcgen05 mainloop (avssl_moco_infonce_sweep with the bank in the
// queue's place, T = 1, writes logits[:, 1:]); this file is the top-k behind it: every row of `dist` comes from HBM
// once (a second traversal is served by L2), and the result is exact (no approximation, ties broken towards the
// smaller index, deterministic).
//
//   pass 1  grid (G segments, N rows): the CTA streams its segment of the row (128-bit loads, order-preserving 32-bit
//           keys) while every thread tracks the maximum of the elements it loads.  The k-th largest of those 256
//           (k > 256: 1024) group maxima is a lower bound of the segment's k-th largest element -- k groups hold an
//           element at least that large -- so a second traversal (out of L2) appends only the few hundred elements
//           at or above it to a candidate list, which is sorted (bitonic, 64-bit composites key << 32 | ~index).  A
//           segment whose candidates overflow the list (heavy ties) takes the general path instead: 4 x 8-bit radix
//           select on shared-memory histograms, then the keys above the k-th plus the needed ties, smallest indices
//           first.  Segments are sized by the row count (about three CTAs per SM in all): the sorts are per segment.
//   pass 2  grid (N rows): with m = ceil(k / G), the k-th largest of the row is at least the smallest of the lists'
//           m-th entries, so only the lists' prefixes down to that value are packed and sorted; values are rescaled by
//           ||q_row|| when the similarities were computed on normalised queries.
#include "common.cuh"

namespace avssl {
namespace {

constexpr int kKnnThreads = 256;
constexpr int kKnnMaxKP = 1024;         // k is padded to a power of two KP <= 1024
constexpr int kKnnMergeMax = 8192;      // pass 2 holds at most this many candidates per row (64 KB of composites)
constexpr int kKnnSegMin = 8192;        // shortest segment worth its own pair of sorts

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

// Descending bitonic sort of n (power of two) elements in shared memory by the whole CTA.
template <typename T>
__device__ void bitonic_desc(T* s, int n) {
  for (int size = 2; size <= n; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = threadIdx.x; i < (n >> 1); i += blockDim.x) {
        const int pos = 2 * i - (i & (stride - 1));
        const T a = s[pos], b = s[pos + stride];
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
  int N, M, k, KP, G, seg, cap;  // cap: candidate list entries (power of two >= 2 KP)
  unsigned long long* cand;  // [N, G, KP]
  const float* q;            // [N, D] or null: yd *= ||q_row||
  int D;
  float* yd;       // [N, k]
  long long* yi;   // [N, k]
};

// Visits every element of a row segment: f(key, index in the segment, load round).  The row starts at an arbitrary
// 4-byte offset (logits[:, 1:]): scalar head up to the first 16-byte boundary, 128-bit body with four loads in flight
// per thread, scalar tail.  The same traversal serves every pass, so an element always lands in the same thread.
template <typename F>
__device__ __forceinline__ void knn_scan(const float* __restrict__ src, int L, int t, F f) {
  const int head = min(L, (int)((16u - (unsigned)(reinterpret_cast<uintptr_t>(src) & 15u)) & 15u) >> 2);
  const int nvec = (L - head) >> 2;
  if (t < head) f(order_key(__ldg(src + t)), t, 0);
  const float4* src4 = reinterpret_cast<const float4*>(src + head);
  int round = 0;
  for (int v0 = 0; v0 < nvec; v0 += 4 * kKnnThreads) {
    float4 x[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int v = v0 + u * kKnnThreads + t;
      x[u] = v < nvec ? __ldg(src4 + v) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u, ++round) {
      const int v = v0 + u * kKnnThreads + t;
      if (v < nvec) {
        const int i = head + 4 * v;
        f(order_key(x[u].x), i, round);
        f(order_key(x[u].y), i + 1, round);
        f(order_key(x[u].z), i + 2, round);
        f(order_key(x[u].w), i + 3, round);
      }
    }
  }
  const int done = head + 4 * nvec;
  if (t < L - done) f(order_key(__ldg(src + done + t)), done + t, 0);
}

// Pass 1.  Nothing of the segment is staged: it is streamed from HBM once for the group maxima and a second time --
// out of L2, the row was just written by the similarity kernel or read by the first traversal -- for the candidates,
// so a segment can be as long as the row and the per-CTA sorts are paid once per (row, segment).
__global__ void __launch_bounds__(kKnnThreads) knn_segment_topk_kernel(const KnnArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned long long* cand_s = reinterpret_cast<unsigned long long*>(smem_raw);  // [cap]
  uint32_t* gm = reinterpret_cast<uint32_t*>(cand_s + a.cap);                    // [256 or 1024] group maxima
  __shared__ unsigned hist[256];
  __shared__ unsigned s_bin, s_need, s_cnt, s_run;
  __shared__ unsigned s_warp[kKnnThreads / 32];

  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int g = blockIdx.x, row = blockIdx.y;
  const int base = g * a.seg;
  const int L = min(a.seg, a.M - base);  // >= 1 by construction of G
  const float* src = a.dist + (int64_t)row * a.ld + base;
  const int gpt = a.k > kKnnThreads ? 4 : 1;  // groups per thread: at least k groups in all
  unsigned long long* out = a.cand + ((int64_t)row * a.G + g) * a.KP;
  if (t == 0) {
    s_cnt = 0;
    s_run = 0;
  }

  // ---- fast path: threshold from the group maxima, candidates appended and sorted
  uint32_t bound = 0u;  // L <= cap: everything is a candidate
  if (L > a.cap) {      // uniform.  (A group without elements has maximum 0: it can only loosen the bound.)
    uint32_t gmax[4] = {0u, 0u, 0u, 0u};
    knn_scan(src, L, t, [&](uint32_t key, int, int round) {
      const int gi = gpt == 4 ? (round & 3) : 0;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (j == gi) gmax[j] = max(gmax[j], key);
    });
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (j < gpt) gm[t * gpt + j] = gmax[j];
    __syncthreads();
    bitonic_desc(gm, kKnnThreads * gpt);
    bound = gm[a.k - 1];
  }
  __syncthreads();
  knn_scan(src, L, t, [&](uint32_t key, int i, int) {
    if (key >= bound) {
      const unsigned slot = atomicAdd(&s_cnt, 1u);
      if (slot < (unsigned)a.cap) cand_s[slot] = composite(key, (uint32_t)(base + i));
    }
  });
  __syncthreads();
  const unsigned total = s_cnt;
  if (total <= (unsigned)a.cap) {  // uniform
    int n = a.KP;
    while (n < (int)total) n <<= 1;
    for (int i = (int)total + t; i < n; i += kKnnThreads) cand_s[i] = 0ull;  // below every real candidate
    __syncthreads();
    bitonic_desc(cand_s, n);
    for (int i = t; i < a.KP; i += kKnnThreads) out[i] = cand_s[i];
    return;
  }

  // ---- general path (the candidate list overflowed: heavy ties): radix select of the k-th largest key, 8 bits a pass
  __syncthreads();
  for (int i = t; i < a.KP; i += kKnnThreads) cand_s[i] = 0ull;
  if (t == 0) s_cnt = 0;
  __syncthreads();
  const int k = a.k;  // L > cap >= 2 k here
  uint32_t prefix = 0, mask = 0;
  unsigned need = (unsigned)k;
  for (int shift = 24; shift >= 0; shift -= 8) {
    hist[t] = 0;  // kKnnThreads == 256 bins
    __syncthreads();
    knn_scan(src, L, t, [&](uint32_t key, int, int) {
      if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1u);
    });
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
  const uint32_t kth = prefix;  // k-th largest key of the segment; `need` of the elements equal to it are wanted
  // collect: everything above kth (any order), then `need` elements equal to kth, smallest indices first
  knn_scan(src, L, t, [&](uint32_t key, int i, int) {
    if (key > kth) cand_s[atomicAdd(&s_cnt, 1u)] = composite(key, (uint32_t)(base + i));
  });
  __syncthreads();
  const unsigned n_gt = s_cnt;  // == k - need
  for (int i0 = 0; i0 < L; i0 += kKnnThreads) {  // index order: 256 consecutive elements per round
    const int i = i0 + t;
    const bool eq = i < L && order_key(__ldg(src + i)) == kth;
    const unsigned bal = __ballot_sync(0xffffffffu, eq);
    if (lane == 0) s_warp[warp] = __popc(bal);
    __syncthreads();
    unsigned before = s_run, total_eq = 0;
#pragma unroll
    for (int w = 0; w < kKnnThreads / 32; ++w) {
      const unsigned cw = s_warp[w];
      if (w < warp) before += cw;
      total_eq += cw;
    }
    const unsigned rank = before + __popc(bal & ((1u << lane) - 1u));
    if (eq && rank < need) cand_s[n_gt + rank] = composite(kth, (uint32_t)(base + i));
    __syncthreads();
    if (t == 0) s_run += total_eq;
    __syncthreads();
    if (s_run >= need) break;  // uniform: s_run is shared
  }
  __syncthreads();
  bitonic_desc(cand_s, a.KP);
  for (int i = t; i < a.KP; i += kKnnThreads) out[i] = cand_s[i];
}

// Pass 2.  With m = ceil(k / G), every list holds m entries at or above T = min over lists of their m-th entry, G m >= k
// entries in all: the k-th largest of the row is >= T and only the lists' prefixes down to T -- a few hundred entries
// for similar segments -- can be part of the answer.  The prefixes are packed into shared memory and sorted.
__global__ void __launch_bounds__(kKnnThreads) knn_merge_topk_kernel(const KnnArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned long long* s = reinterpret_cast<unsigned long long*>(smem_raw);  // survivors, <= G * KP
  __shared__ float s_scale;
  __shared__ unsigned long long s_thr;
  __shared__ int s_cnt[32], s_off[33];
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int row = blockIdx.x;
  const unsigned long long* in = a.cand + (int64_t)row * a.G * a.KP;  // [G][KP], every list sorted descending, 0-padded
  const int m = (a.k + a.G - 1) / a.G;
  if (warp == 0) {
    unsigned long long thr = lane < a.G ? in[lane * a.KP + m - 1] : ~0ull;  // G <= 32
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long other = __shfl_xor_sync(0xffffffffu, thr, o);
      thr = other < thr ? other : thr;
    }
    if (lane == 0) s_thr = thr;  // 0 when a list is shorter than m: nothing is pruned then
    const float sc = a.q ? sqrtf(row_sumsq(a.q + (int64_t)row * a.D, a.D, lane)) : 1.f;
    if (lane == 0) s_scale = sc;
  }
  __syncthreads();
  const unsigned long long thr = s_thr;
  for (int g = warp; g < a.G; g += kKnnThreads / 32) {  // survivors of a sorted list are a prefix: count it
    int cnt = 0;
    for (int j0 = 0; j0 < a.KP; j0 += 32) {
      const unsigned long long c = in[g * a.KP + j0 + lane];
      const unsigned bal = __ballot_sync(0xffffffffu, c != 0ull && c >= thr);
      cnt += __popc(bal);
      if (bal != 0xffffffffu) break;  // uniform
    }
    if (lane == 0) s_cnt[g] = cnt;
  }
  __syncthreads();
  if (warp == 0) {
    const int c = lane < a.G ? s_cnt[lane] : 0;
    int incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int n = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += n;
    }
    s_off[lane] = incl - c;
    if (lane == 31) s_off[32] = incl;
  }
  __syncthreads();
  const int total = s_off[32];  // >= k
  for (int g = warp; g < a.G; g += kKnnThreads / 32) {
    const int cnt = s_cnt[g], off = s_off[g];
    for (int j = lane; j < cnt; j += 32) s[off + j] = in[g * a.KP + j];
  }
  int n = 32;
  while (n < total) n <<= 1;
  for (int i = total + t; i < n; i += kKnnThreads) s[i] = 0ull;
  __syncthreads();
  if (a.G > 1) bitonic_desc(s, n);  // one segment: already sorted
  const float scale = s_scale;
  for (int i = t; i < a.k; i += kKnnThreads) {
    const unsigned long long c = s[i];
    const float v = key_value((uint32_t)(c >> 32));
    a.yd[(int64_t)row * a.k + i] = a.q ? v * scale : v;
    a.yi[(int64_t)row * a.k + i] = (long long)(0xffffffffu - (uint32_t)(c & 0xffffffffu));
  }
}

// candidate list + group maxima (one per thread, four when k > 256)
size_t knn_smem1(int cap, int k) { return (size_t)cap * 8 + (size_t)kKnnThreads * (k > kKnnThreads ? 4 : 1) * 4; }

// Segment plan shared by the workspace query and the launch: about three CTAs per SM over all rows, segments of at
// least kKnnSegMin elements (the two sorts of pass 1 are paid per segment), at most 32 lists of KP for pass 2.
bool knn_plan(int N, int M, int k, int* KP, int* G, int* seg, int* cap) {
  if (N < 1 || M < 1 || k < 1 || k > M || k > kKnnMaxKP) return false;
  int kp = 32;
  while (kp < k) kp <<= 1;
  int gmax = kKnnMergeMax / kp < 32 ? kKnnMergeMax / kp : 32;
  const int by_len = (M + kKnnSegMin - 1) / kKnnSegMin;
  if (gmax > by_len) gmax = by_len;
  int sms = sm_count();
  if (sms < 1) sms = 148;
  int g = (3 * sms + N - 1) / N;
  if (g > gmax) g = gmax;
  if (g < 1) g = 1;
  const int sg = (M + g - 1) / g;
  g = (M + sg - 1) / sg;  // no empty segment
  *KP = kp;
  *G = g;
  *seg = sg;
  *cap = kp * 2 > 1024 ? kp * 2 : 1024;
  return true;
}

}  // namespace
}  // namespace avssl

using namespace avssl;

extern "C" size_t avssl_topk_rows_workspace_bytes(int N, int M, int k) {
  int KP, G, seg, cap;
  if (!knn_plan(N, M, k, &KP, &G, &seg, &cap)) return 0;
  return (size_t)N * G * KP * sizeof(unsigned long long);
}

extern "C" int avssl_topk_rows(const float* dist, int64_t ld, int N, int M, int k, const float* q_scale_rows, int D,
                               float* yd_out, int64_t* yi_out, void* workspace, size_t workspace_bytes, void* stream) {
  AVSSL_REQUIRE(dist && yd_out && yi_out && workspace, AVSSL_ERR_INVALID_ARGUMENT, "topk_rows: null pointer");
  AVSSL_REQUIRE(N >= 1 && N <= 65535 && M >= 1 && ld >= M && k >= 1 && k <= M, AVSSL_ERR_INVALID_ARGUMENT,
                "topk_rows: bad sizes N=%d M=%d k=%d", N, M, k);
  AVSSL_REQUIRE(!q_scale_rows || D > 0, AVSSL_ERR_INVALID_ARGUMENT, "topk_rows: D must be positive with q_scale_rows");
  KnnArgs a;
  AVSSL_REQUIRE(knn_plan(N, M, k, &a.KP, &a.G, &a.seg, &a.cap), AVSSL_ERR_UNSUPPORTED,
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
    AVSSL_CUDA_OK(cudaFuncSetAttribute(knn_merge_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)(kKnnMergeMax * sizeof(unsigned long long))));
  }
  knn_segment_topk_kernel<<<dim3(a.G, N), kKnnThreads, knn_smem1(a.cap, a.k), s>>>(a);
  AVSSL_LAUNCH_OK("knn_segment_topk_kernel");
  int P2 = 32;
  while (P2 < a.G * a.KP) P2 <<= 1;  // the survivors are padded to a power of two for the sort
  knn_merge_topk_kernel<<<N, kKnnThreads, (size_t)P2 * 8, s>>>(a);
  AVSSL_LAUNCH_OK("knn_merge_topk_kernel");
  return AVSSL_OK;
}
