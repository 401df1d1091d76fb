// K10 / K11 — SwAV: single-launch Sinkhorn-Knopp and the fused soft-target
// cross-entropy (forward + backward).
//   sinkhorn            models/contrastive.py:872-887 (+ exp(out/eps).t(), :665-666)
//   swap-prediction CE  models/contrastive.py:672-679, KLDivLoss :912-916
#include <cooperative_groups.h>

#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "sm100_ptx.cuh"

namespace avssl {

// ------------------------------------------------------------------ Sinkhorn-Knopp
// The reference rewrites the whole [P, B] matrix ~2 times per iteration (~25 launches
// per call).  Sinkhorn only ever rescales rows and columns, so the result is
//     code[b][k] = alpha_k * Q0[b][k] * beta_b,   Q0 = exp(score / eps)
// and the iterations only update the two scaling vectors:
//     row step: alpha_k = (1/P) / sum_b Q0[b][k] beta_b
//     col step: beta_b  = (1/B) / sum_k alpha_k Q0[b][k]
//     final   : beta_b  =   1   / sum_k alpha_k Q0[b][k]
// (the initial Q /= sum(Q) is a uniform scale that the first row step removes).
// One cooperative launch: every CTA keeps its samples' Q0 rows resident in shared
// memory for all passes; column sums are CTA-local (warp-shuffle reductions), row sums
// are combined across CTAs through a [grid][P] scratch and two grid barriers per
// iteration, in a fixed order (deterministic).
struct SinkArgs {
  const float* scores;  // [Btot, P]
  int Btot, P;
  float inv_eps;
  int iters, keep_last;
  float* out;           // [keep_last, P]
  float* g_part;        // [grid][P]
  float* g_alpha;       // [P]
  unsigned* bar;        // [2] count, generation (zero-initialised once)
  int spc;              // samples per CTA
  int slab_in_smem;
};

__device__ __forceinline__ void grid_barrier(unsigned* bar, unsigned nblocks) {
  __syncthreads();
  if (threadIdx.x == 0) {
    volatile unsigned* gen = bar + 1;
    const unsigned g = *gen;
    __threadfence();
    if (atomicAdd(bar, 1u) == nblocks - 1) {
      bar[0] = 0u;
      __threadfence();
      atomicAdd(bar + 1, 1u);
    } else {
      while (*gen == g) {
      }
    }
    __threadfence();
  }
  __syncthreads();
}

__global__ void __launch_bounds__(256, 1) sinkhorn_kernel(const SinkArgs a) {
  extern __shared__ __align__(16) float sm[];
  float* alpha = sm;                 // [P]
  float* beta = alpha + a.P;         // [spc]
  float* slab = beta + ((a.spc + 3) & ~3);  // [spc][P] when it fits
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
  const int b0 = blockIdx.x * a.spc;
  const int nb = max(0, min(a.spc, a.Btot - b0));
  const int P = a.P;

  // Q0 rows of this CTA
  if (a.slab_in_smem) {
    for (int b = 0; b < nb; ++b) {
      const float* src = a.scores + (size_t)(b0 + b) * P;
      for (int k = tid; k < P; k += blockDim.x) slab[b * P + k] = expf(src[k] * a.inv_eps);
    }
  }
  for (int b = tid; b < nb; b += blockDim.x) beta[b] = 1.f;
  for (int k = tid; k < P; k += blockDim.x) alpha[k] = 1.f;
  __syncthreads();
  auto q0 = [&](int b, int k) -> float {
    return a.slab_in_smem ? slab[b * P + k] : expf(a.scores[(size_t)(b0 + b) * P + k] * a.inv_eps);
  };

  const float r = 1.f / (float)P, c = 1.f / (float)a.Btot;
  const int kp = (P + gridDim.x - 1) / gridDim.x;
  for (int it = 0; it < a.iters; ++it) {
    // row step, part 1: partial row sums over this CTA's samples
    for (int k = tid; k < P; k += blockDim.x) {
      float s = 0.f;
      for (int b = 0; b < nb; ++b) s = fmaf(q0(b, k), beta[b], s);
      a.g_part[(size_t)blockIdx.x * P + k] = s;
    }
    grid_barrier(a.bar, gridDim.x);
    // row step, part 2: this CTA finishes a slice of k (one warp per k, lanes over CTAs)
    for (int kk = warp; kk < kp; kk += nw) {
      const int k = blockIdx.x * kp + kk;
      if (k < P) {
        float s = 0.f;
        for (int g = lane; g < (int)gridDim.x; g += 32) s += __ldcg(a.g_part + (size_t)g * P + k);
        s = warp_sum(s);
        if (lane == 0) a.g_alpha[k] = r / s;
      }
    }
    grid_barrier(a.bar, gridDim.x);
    for (int k = tid; k < P; k += blockDim.x) alpha[k] = __ldcg(a.g_alpha + k);
    __syncthreads();
    if (it < a.iters - 1) {
      // col step (local): one warp per sample
      for (int b = warp; b < nb; b += nw) {
        float s = 0.f;
        for (int k = lane; k < P; k += 32) s = fmaf(alpha[k], q0(b, k), s);
        s = warp_sum(s);
        if (lane == 0) beta[b] = c / s;
      }
      __syncthreads();
    }
  }
  // final column normalisation + output of the kept rows
  const int first_keep = a.Btot - a.keep_last;
  for (int b = warp; b < nb; b += nw) {
    const int gb = b0 + b;
    if (gb < first_keep) continue;
    float s = 0.f;
    for (int k = lane; k < P; k += 32) s = fmaf(alpha[k], q0(b, k), s);
    s = warp_sum(s);
    const float bb = 1.f / s;
    float* dst = a.out + (size_t)(gb - first_keep) * P;
    for (int k = lane; k < P; k += 32) dst[k] = alpha[k] * q0(b, k) * bb;
  }
}

// Single-cluster variant: when the whole Q0 matrix fits the shared memory of ONE thread-block
// cluster (16 CTAs x ~216 KB: Btot x P <= ~800k elements, e.g. 256 x 3000), the iterations need
// no global memory and no grid barrier at all -- row sums are combined through distributed shared
// memory in a fixed order (deterministic) behind hardware cluster barriers.  CTA c owns `spc`
// samples; it finishes the prototype slice [c * kslice, (c + 1) * kslice) of every row step and
// broadcasts the new alpha values into all CTAs of the cluster.
constexpr int kSinkClusterThreads = 1024;

__global__ void __launch_bounds__(kSinkClusterThreads, 1) sinkhorn_cluster_kernel(const SinkArgs a) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  extern __shared__ __align__(16) float sm[];
  const int P = a.P;
  const int Pp = (P + 3) & ~3;
  float* alpha = sm;                          // [Pp]
  float* part = alpha + Pp;                   // [Pp]  this CTA's partial row sums
  float* beta = part + Pp;                    // [spc]
  float* slab = beta + ((a.spc + 3) & ~3);    // [spc][P]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = kSinkClusterThreads >> 5;
  const unsigned nc = cluster.num_blocks(), cr = cluster.block_rank();
  const int b0 = (int)cr * a.spc;
  const int nb = max(0, min(a.spc, a.Btot - b0));

  // Q0 rows of this CTA: one contiguous block of nb * P scores.  Only 16 SMs pull the whole matrix,
  // so every thread keeps 4 x 128-bit loads in flight.
  if ((P & 3) == 0 && (reinterpret_cast<uintptr_t>(a.scores) & 15u) == 0) {
    const float4* src4 = reinterpret_cast<const float4*>(a.scores + (size_t)b0 * P);
    float4* slab4 = reinterpret_cast<float4*>(slab);
    const int n4 = nb * (P >> 2);
    for (int i0 = 0; i0 < n4; i0 += 4 * kSinkClusterThreads) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * kSinkClusterThreads + tid;
        v[u] = i < n4 ? ldg_stream(src4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * kSinkClusterThreads + tid;
        if (i < n4)
          slab4[i] = make_float4(expf(v[u].x * a.inv_eps), expf(v[u].y * a.inv_eps), expf(v[u].z * a.inv_eps),
                                 expf(v[u].w * a.inv_eps));
      }
    }
  } else {
    for (int b = 0; b < nb; ++b) {
      const float* src = a.scores + (size_t)(b0 + b) * P;
      for (int k = tid; k < P; k += kSinkClusterThreads) slab[b * P + k] = expf(__ldg(src + k) * a.inv_eps);
    }
  }
  for (int b = tid; b < nb; b += kSinkClusterThreads) beta[b] = 1.f;
  __syncthreads();

  const float r = 1.f / (float)P, c = 1.f / (float)a.Btot;
  const int kslice = (P + (int)nc - 1) / (int)nc;
  const int k_lo = (int)cr * kslice, k_hi = min(P, k_lo + kslice);
  for (int it = 0; it < a.iters; ++it) {
    // row step, part 1: partial sums over this CTA's samples
    for (int k = tid; k < P; k += kSinkClusterThreads) {
      float s0 = 0.f, s1 = 0.f;
      int b = 0;
#pragma unroll 4
      for (; b + 1 < nb; b += 2) {
        s0 = fmaf(slab[b * P + k], beta[b], s0);
        s1 = fmaf(slab[(b + 1) * P + k], beta[b + 1], s1);
      }
      if (b < nb) s0 = fmaf(slab[b * P + k], beta[b], s0);
      part[k] = s0 + s1;
    }
    cluster.sync();
    // row step, part 2: finish this CTA's slice over all CTAs (rank order), broadcast alpha
    for (int k = k_lo + tid; k < k_hi; k += kSinkClusterThreads) {
      float s = 0.f;
      for (unsigned g = 0; g < nc; ++g) s += cluster.map_shared_rank(part, g)[k];
      const float al = r / s;
      for (unsigned g = 0; g < nc; ++g) cluster.map_shared_rank(alpha, g)[k] = al;
    }
    cluster.sync();
    if (it < a.iters - 1) {
      // col step (local): one warp per sample
      // one warp per sample, two independent accumulators per lane
      for (int b = warp; b < nb; b += nw) {
        float s0 = 0.f, s1 = 0.f;
        int k = lane;
#pragma unroll 4
        for (; k + 32 < P; k += 64) {
          s0 = fmaf(alpha[k], slab[b * P + k], s0);
          s1 = fmaf(alpha[k + 32], slab[b * P + k + 32], s1);
        }
        if (k < P) s0 = fmaf(alpha[k], slab[b * P + k], s0);
        const float sum = warp_sum(s0 + s1);
        if (lane == 0) beta[b] = c / sum;
      }
      __syncthreads();
    }
  }
  if (a.iters == 0) {
    for (int k = tid; k < P; k += kSinkClusterThreads) alpha[k] = 1.f;
    __syncthreads();
  }
  // final column normalisation + output of the kept rows
  const int first_keep = a.Btot - a.keep_last;
  for (int b = warp; b < nb; b += nw) {
    const int gb = b0 + b;
    if (gb < first_keep) continue;
    float s = 0.f;
    for (int k = lane; k < P; k += 32) s = fmaf(alpha[k], slab[b * P + k], s);
    s = warp_sum(s);
    const float bb = 1.f / s;
    float* dst = a.out + (size_t)(gb - first_keep) * P;
    for (int k = lane; k < P; k += 32) __stcs(dst + k, alpha[k] * slab[b * P + k] * bb);
  }
}

// ------------------------------------------------------- soft-target cross-entropy
// One CTA per score row (crop v, sample r).  For every code set `a` with weight
// w[a][v] != 0:   loss += w * ( lse * sum_k code - sum_k code * s/T )
//                 dscore[k] += w * ( softmax_k * sum_k code - code_k ) / T
// (log(softmax(.)) of the reference == s/T - lse.)
constexpr int kMaxAssign = 4, kMaxCrops = 16;
struct SwavCeArgs {
  const float* scores;  // [n_crops*bs, P]
  const float* codes;   // [n_assign, bs, P]
  int n_crops, n_assign, bs, P;
  float inv_T;
  float w[kMaxAssign * kMaxCrops];
  float* loss_out;
  float* dscores;       // may be null
  float* row_loss;      // [n_crops*bs]
  unsigned* counter;
};

__global__ void __launch_bounds__(256) swav_ce_kernel(const SwavCeArgs a) {
  __shared__ float s_red[32];
  __shared__ unsigned s_last;
  const int row = blockIdx.x;
  const int v = row / a.bs, r = row % a.bs;
  const int P = a.P, tid = threadIdx.x;
  const float* s = a.scores + (size_t)row * P;

  float mx = -INFINITY;
  for (int k = tid; k < P; k += blockDim.x) mx = fmaxf(mx, s[k] * a.inv_T);
  mx = warp_max(mx);
  __syncthreads();
  if ((tid & 31) == 0) s_red[tid >> 5] = mx;
  __syncthreads();
  mx = s_red[0];
  for (int w = 1; w < (int)(blockDim.x >> 5); ++w) mx = fmaxf(mx, s_red[w]);
  float se = 0.f;
  for (int k = tid; k < P; k += blockDim.x) se += expf(s[k] * a.inv_T - mx);
  se = block_sum(se, s_red);
  const float lse = mx + logf(se);

  float loss = 0.f, wsum = 0.f;  // wsum = sum_a w * sum_k code
  for (int as = 0; as < a.n_assign; ++as) {
    const float w = a.w[as * a.n_crops + v];
    if (w == 0.f) continue;
    const float* code = a.codes + ((size_t)as * a.bs + r) * P;
    float dot = 0.f, sq = 0.f;
    for (int k = tid; k < P; k += blockDim.x) {
      const float cq = code[k];
      dot = fmaf(cq, s[k] * a.inv_T, dot);
      sq += cq;
    }
    dot = block_sum(dot, s_red);
    sq = block_sum(sq, s_red);
    loss += w * (lse * sq - dot);
    wsum += w * sq;
  }
  if (a.dscores) {
    float* d = a.dscores + (size_t)row * P;
    for (int k = tid; k < P; k += blockDim.x) {
      float g = expf(s[k] * a.inv_T - lse) * wsum;
      for (int as = 0; as < a.n_assign; ++as) {
        const float w = a.w[as * a.n_crops + v];
        if (w != 0.f) g -= w * a.codes[((size_t)as * a.bs + r) * P + k];
      }
      d[k] = g * a.inv_T;
    }
  }
  if (tid == 0) a.row_loss[row] = loss;
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    s_last = (atomicAdd(a.counter, 1u) == gridDim.x - 1) ? 1u : 0u;
  }
  __syncthreads();
  if (s_last) {
    __threadfence();
    float tot = 0.f;
    const int n = a.n_crops * a.bs;
    for (int i = tid; i < n; i += blockDim.x) tot += reinterpret_cast<volatile float*>(a.row_loss)[i];
    tot = block_sum(tot, s_red);
    if (tid == 0) {
      *a.loss_out = tot;
      *a.counter = 0u;
    }
  }
}

// Register-resident variant (P % 4 == 0, P <= 256 * 4 * kCeVec): the score row and the code rows
// are read from global memory exactly once as 128-bit loads, exp() is evaluated once per element
// and reused by the gradient, and the per-code reductions share one block-wide reduction.
// Algorithmic traffic: read scores, read codes (L2-resident: every code row serves n_crops - 1
// score rows), write dscores.
constexpr int kCeVec = 3;  // float4 per thread: rows up to 3072 columns

template <int kN>
__device__ __forceinline__ void block_sum_n(float (&v)[kN], float* scratch /* >= 32 * kN */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int i = 0; i < kN; ++i) v[i] = warp_sum(v[i]);
  __syncthreads();
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < kN; ++i) scratch[i * 32 + warp] = v[i];
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < kN; ++i) {
    float r = lane < nw ? scratch[i * 32 + lane] : 0.f;
    v[i] = warp_sum(r);
  }
}

template <int kA>  // code sets held in registers (n_assign <= kA)
__global__ void __launch_bounds__(256) swav_ce_reg_kernel(const SwavCeArgs a) {
  __shared__ float s_red[32 * (1 + 2 * kA)];
  __shared__ unsigned s_last;
  const int row = blockIdx.x;
  const int v = row / a.bs, r = row % a.bs;
  const int P4 = a.P >> 2, tid = threadIdx.x;
  const float4* s4 = reinterpret_cast<const float4*>(a.scores + (size_t)row * a.P);

  float4 x[kCeVec];
  float mx = -INFINITY;
#pragma unroll
  for (int j = 0; j < kCeVec; ++j) {
    const int c = tid + j * 256;
    if (c < P4) {
      x[j] = ldg_stream(s4 + c);
      x[j].x *= a.inv_T; x[j].y *= a.inv_T; x[j].z *= a.inv_T; x[j].w *= a.inv_T;
      mx = fmaxf(mx, fmaxf(fmaxf(x[j].x, x[j].y), fmaxf(x[j].z, x[j].w)));
    } else {
      x[j] = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    }
  }
  // code rows in flight while the max is reduced
  float wa[kA];
  float4 cd[kA][kCeVec];
#pragma unroll
  for (int as = 0; as < kA; ++as) {
    wa[as] = as < a.n_assign ? a.w[as * a.n_crops + v] : 0.f;
    const float4* c4 = reinterpret_cast<const float4*>(a.codes + ((size_t)as * a.bs + r) * a.P);
#pragma unroll
    for (int j = 0; j < kCeVec; ++j) {
      const int c = tid + j * 256;
      cd[as][j] = (wa[as] != 0.f && c < P4) ? __ldg(c4 + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  mx = warp_max(mx);
  if ((tid & 31) == 0) s_red[tid >> 5] = mx;
  __syncthreads();
  mx = s_red[0];
#pragma unroll
  for (int w = 1; w < 8; ++w) mx = fmaxf(mx, s_red[w]);

  // one reduction for: sum exp, and per code set (dot with s/T, sum of the code)
  float red[1 + 2 * kA];
#pragma unroll
  for (int i = 0; i < 1 + 2 * kA; ++i) red[i] = 0.f;
  float4 e[kCeVec];
#pragma unroll
  for (int j = 0; j < kCeVec; ++j) {
    e[j] = make_float4(__expf(x[j].x - mx), __expf(x[j].y - mx), __expf(x[j].z - mx), __expf(x[j].w - mx));
    red[0] += (e[j].x + e[j].y) + (e[j].z + e[j].w);
    const bool in = tid + j * 256 < P4;
#pragma unroll
    for (int as = 0; as < kA; ++as) {
      const float4 c = cd[as][j];
      if (in) {
        red[1 + 2 * as] = fmaf(c.x, x[j].x, fmaf(c.y, x[j].y, fmaf(c.z, x[j].z, fmaf(c.w, x[j].w, red[1 + 2 * as]))));
        red[2 + 2 * as] += (c.x + c.y) + (c.z + c.w);
      }
    }
  }
  block_sum_n<1 + 2 * kA>(red, s_red);
  const float lse = mx + logf(red[0]);
  float loss = 0.f, wsum = 0.f;  // wsum = sum_a w * sum_k code
#pragma unroll
  for (int as = 0; as < kA; ++as) {
    loss += wa[as] * (lse * red[2 + 2 * as] - red[1 + 2 * as]);
    wsum += wa[as] * red[2 + 2 * as];
  }
  if (a.dscores) {
    float4* d4 = reinterpret_cast<float4*>(a.dscores + (size_t)row * a.P);
    const float pscale = wsum / red[0];  // softmax_k * wsum = e_k * wsum / sum(e)
#pragma unroll
    for (int j = 0; j < kCeVec; ++j) {
      const int c = tid + j * 256;
      if (c < P4) {
        float4 g = make_float4(e[j].x * pscale, e[j].y * pscale, e[j].z * pscale, e[j].w * pscale);
#pragma unroll
        for (int as = 0; as < kA; ++as) {
          g.x = fmaf(-wa[as], cd[as][j].x, g.x);
          g.y = fmaf(-wa[as], cd[as][j].y, g.y);
          g.z = fmaf(-wa[as], cd[as][j].z, g.z);
          g.w = fmaf(-wa[as], cd[as][j].w, g.w);
        }
        st_stream(d4 + c, make_float4(g.x * a.inv_T, g.y * a.inv_T, g.z * a.inv_T, g.w * a.inv_T));
      }
    }
  }
  if (tid == 0) {
    a.row_loss[row] = loss;
    __threadfence();
    s_last = (atomicAdd(a.counter, 1u) == gridDim.x - 1) ? 1u : 0u;
  }
  __syncthreads();
  if (s_last) {
    __threadfence();
    float tot = 0.f;
    const int n = a.n_crops * a.bs;
    for (int i = tid; i < n; i += blockDim.x) tot += __ldcg(a.row_loss + i);
    tot = block_sum(tot, s_red);
    if (tid == 0) {
      *a.loss_out = tot;
      *a.counter = 0u;
    }
  }
}


// Sample-major variant (the hot one at cfg5): one CTA per (sample r, group of crops).  The score rows of ONE sample
// across the crops all meet the same code rows codes[as][r], so the CTA lands those once (bulk copy into shared
// memory, then registers, their sums reduced once) and streams the sample's score rows through a two-stage bulk-copy
// pipeline: L2 -> SM traffic is the algorithmic 1 + n_assign / crops_per_cta rows per score row.
// ncu on the first version (per-thread arithmetic of swav_ce_reg_kernel, profiles/r2_k11_ncu.md) showed the kernel
// ISSUE-bound, not memory-bound: ~800 instructions per warp and row, 22 per element.  Hence the arithmetic diet:
// scores go to the log2 domain with one multiply (exp = one FADD + MUFU.EX2), the code sums are taken once per CTA
// instead of once per row, zero weights need no select (they multiply every use), and 1/T is folded into the two
// factors of the gradient.  Same mathematics as the reference's -(q * log_softmax(s / T)).sum(); rounding differs
// from swav_ce_reg_kernel in the last bits (the test bounds it).
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int kA>
__global__ void __launch_bounds__(256) swav_ce_sample_kernel(const SwavCeArgs a, int crops_per_cta) {
  extern __shared__ __align__(16) uint8_t pipe_smem[];
  __shared__ float s_red[32 * (1 + kA)];
  __shared__ unsigned s_last;
  __shared__ __align__(8) uint64_t full[2];
  __shared__ __align__(8) uint64_t codes_full;
  constexpr float kLog2e = 1.4426950408889634f, kLn2 = 0.6931471805599453f;
  const int tid = threadIdx.x;
  const int P4 = a.P >> 2;
  const int n_rows = a.n_crops * a.bs;
  const uint32_t row_bytes = (uint32_t)a.P * 4u;
  const int r = blockIdx.x % a.bs;
  const int v0 = (blockIdx.x / a.bs) * crops_per_cta;
  const int v1 = min(v0 + crops_per_cta, a.n_crops);
  auto code_ptr = [&](int as) { return reinterpret_cast<float4*>(pipe_smem + (size_t)as * row_bytes); };
  auto stage_ptr = [&](int st) { return reinterpret_cast<float4*>(pipe_smem + (size_t)(kA + st) * row_bytes); };
  if (tid == 0) {
    ptx::mbar_init(&full[0], 1);
    ptx::mbar_init(&full[1], 1);
    ptx::mbar_init(&codes_full, 1);
    ptx::mbar_fence_init();
    // the first score row before the code rows: the max reduction only needs the scores
    ptx::mbar_arrive_expect_tx(&full[0], row_bytes);
    ptx::bulk_load(stage_ptr(0), a.scores + ((size_t)v0 * a.bs + r) * a.P, row_bytes, &full[0]);
    ptx::mbar_arrive_expect_tx(&codes_full, row_bytes * (uint32_t)min(a.n_assign, kA));
    for (int as = 0; as < kA && as < a.n_assign; ++as)
      ptx::bulk_load(code_ptr(as), a.codes + ((size_t)as * a.bs + r) * a.P, row_bytes, &codes_full);
  }
  __syncthreads();

  const float to_log2 = a.inv_T * kLog2e;
  float4 cd[kA][kCeVec];  // this sample's code rows, resident in registers across the crops
  float csum[kA];         // ... and their sums
  bool have_codes = false;
  int it = 0;
  for (int v = v0; v < v1; ++v, ++it) {
    const int st = it & 1;
    const int row = v * a.bs + r;
    if (tid == 0 && v + 1 < v1) {  // stage st^1 was released by the barriers of the previous iteration
      ptx::mbar_arrive_expect_tx(&full[st ^ 1], row_bytes);
      ptx::bulk_load(stage_ptr(st ^ 1), a.scores + ((size_t)(v + 1) * a.bs + r) * a.P, row_bytes, &full[st ^ 1]);
    }
    ptx::mbar_wait(&full[st], (it >> 1) & 1);

    const float4* xs = stage_ptr(st);
    float4 x[kCeVec];  // s / T in log2 units
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < kCeVec; ++j) {
      const int c = tid + j * 256;
      if (c < P4) {
        x[j] = xs[c];
        x[j].x *= to_log2; x[j].y *= to_log2; x[j].z *= to_log2; x[j].w *= to_log2;
        mx = fmaxf(mx, fmaxf(fmaxf(x[j].x, x[j].y), fmaxf(x[j].z, x[j].w)));
      } else {
        x[j] = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
      }
    }
    if (!have_codes) {  // uniform: first row of the CTA
      ptx::mbar_wait(&codes_full, 0);
#pragma unroll
      for (int as = 0; as < kA; ++as) {
        const float4* c4 = code_ptr(as);
        csum[as] = 0.f;
#pragma unroll
        for (int j = 0; j < kCeVec; ++j) {
          const int c = tid + j * 256;
          cd[as][j] = (as < a.n_assign && c < P4) ? c4[c] : make_float4(0.f, 0.f, 0.f, 0.f);
          csum[as] += (cd[as][j].x + cd[as][j].y) + (cd[as][j].z + cd[as][j].w);
        }
      }
      block_sum_n<kA>(csum, s_red);
      have_codes = true;
    }
    mx = warp_max(mx);
    __syncthreads();  // s_red of the previous reduction is dead
    if ((tid & 31) == 0) s_red[tid >> 5] = mx;
    __syncthreads();  // ... and every thread has read stage st: it may be refilled by the next iteration's issue
    mx = s_red[0];
#pragma unroll
    for (int w = 1; w < 8; ++w) mx = fmaxf(mx, s_red[w]);

    // one reduction for: sum of 2^(x - mx), and per code set the dot product with x
    float red[1 + kA];
#pragma unroll
    for (int i = 0; i < 1 + kA; ++i) red[i] = 0.f;
    float4 e[kCeVec];
#pragma unroll
    for (int j = 0; j < kCeVec; ++j) {
      e[j] = make_float4(ex2_approx(x[j].x - mx), ex2_approx(x[j].y - mx), ex2_approx(x[j].z - mx), ex2_approx(x[j].w - mx));
      red[0] += (e[j].x + e[j].y) + (e[j].z + e[j].w);  // padding columns: 2^(-inf) = 0
      if (tid + j * 256 < P4) {
#pragma unroll
        for (int as = 0; as < kA; ++as) {
          const float4 c = cd[as][j];
          red[1 + as] = fmaf(c.x, x[j].x, fmaf(c.y, x[j].y, fmaf(c.z, x[j].z, fmaf(c.w, x[j].w, red[1 + as]))));
        }
      }
    }
    block_sum_n<1 + kA>(red, s_red);
    const float lse = (mx + log2f(red[0])) * kLn2;
    float loss = 0.f, wsum = 0.f, wn[kA];  // wsum = sum_a w * sum_k code
#pragma unroll
    for (int as = 0; as < kA; ++as) {
      const float w = as < a.n_assign ? a.w[as * a.n_crops + v] : 0.f;
      loss += w * (lse * csum[as] - red[1 + as] * kLn2);
      wsum += w * csum[as];
      wn[as] = -w * a.inv_T;
    }
    if (a.dscores) {
      float4* d4 = reinterpret_cast<float4*>(a.dscores + (size_t)row * a.P);
      const float ps = wsum / red[0] * a.inv_T;  // softmax_k * wsum / T = e_k * ps
#pragma unroll
      for (int j = 0; j < kCeVec; ++j) {
        const int c = tid + j * 256;
        if (c < P4) {
          float4 g = make_float4(e[j].x * ps, e[j].y * ps, e[j].z * ps, e[j].w * ps);
#pragma unroll
          for (int as = 0; as < kA; ++as) {
            g.x = fmaf(wn[as], cd[as][j].x, g.x);
            g.y = fmaf(wn[as], cd[as][j].y, g.y);
            g.z = fmaf(wn[as], cd[as][j].z, g.z);
            g.w = fmaf(wn[as], cd[as][j].w, g.w);
          }
          st_stream(d4 + c, g);
        }
      }
    }
    if (tid == 0) a.row_loss[row] = loss;
  }
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    s_last = (atomicAdd(a.counter, 1u) == gridDim.x - 1) ? 1u : 0u;
  }
  __syncthreads();
  if (s_last) {
    __threadfence();
    float tot = 0.f;
    for (int i = tid; i < n_rows; i += blockDim.x) tot += __ldcg(a.row_loss + i);
    tot = block_sum(tot, s_red);
    if (tid == 0) {
      *a.loss_out = tot;
      *a.counter = 0u;
    }
  }
}

}  // namespace avssl

using namespace avssl;

static int sink_grid(int Btot) {
  int sms = sm_count();
  if (sms <= 0) return -1;
  return Btot < sms ? Btot : sms;
}

// Launches the single-cluster kernel when the problem fits one cluster of 16 (or 8) CTAs and the
// device can schedule such a cluster; returns false to fall back to the cooperative grid kernel.
static bool sinkhorn_try_cluster(SinkArgs& a, cudaStream_t s) {
  static int max_cluster = 0;  // largest schedulable cluster size with the full shared-memory carve-out
  static unsigned long long configured = 0ull;  // device ordinals whose function attributes are set
  constexpr size_t kSmemMax = 227 * 1024;
  if (first_use_on_device(configured)) {
    max_cluster = 0;
    if (cudaFuncSetAttribute(sinkhorn_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemMax) == cudaSuccess &&
        cudaFuncSetAttribute(sinkhorn_cluster_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess) {
      for (int cs = 16; cs >= 8 && !max_cluster; cs >>= 1) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(cs);
        cfg.blockDim = dim3(kSinkClusterThreads);
        cfg.dynamicSmemBytes = kSmemMax;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = cs;
        at[0].val.clusterDim.y = at[0].val.clusterDim.z = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        int n = 0;
        if (cudaOccupancyMaxActiveClusters(&n, sinkhorn_cluster_kernel, &cfg) == cudaSuccess && n >= 1) max_cluster = cs;
      }
    }
    cudaGetLastError();
  }
  if (max_cluster == 0) return false;
  const int cs = max_cluster;
  const int spc = (a.Btot + cs - 1) / cs;
  const size_t Pp = ((size_t)a.P + 3) & ~(size_t)3;
  const size_t smem = sizeof(float) * (2 * Pp + ((spc + 3) & ~3) + (size_t)spc * a.P);
  if (smem > kSmemMax) return false;
  a.spc = spc;
  a.slab_in_smem = 1;
  a.g_part = a.g_alpha = nullptr;
  a.bar = nullptr;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(cs);
  cfg.blockDim = dim3(kSinkClusterThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = cs;
  at[0].val.clusterDim.y = at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  if (cudaLaunchKernelEx(&cfg, sinkhorn_cluster_kernel, a) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return true;
}

extern "C" size_t avssl_sinkhorn_workspace_bytes(int Btot, int P) {
  if (Btot <= 0 || P <= 0) return 0;
  int g = sink_grid(Btot);
  if (g <= 0) {
    cudaGetLastError();
    g = Btot < 148 ? Btot : 148;
  }
  return 256 + 4 * (size_t)P + 256 + 4 * (size_t)g * P + 256;
}

extern "C" int avssl_sinkhorn(const float* scores, int Btot, int P, float eps, int iters, int keep_last,
                              float* codes_out, void* workspace, size_t workspace_bytes, void* stream) {
  AVSSL_REQUIRE(scores && codes_out && workspace, AVSSL_ERR_INVALID_ARGUMENT, "sinkhorn: null pointer");
  AVSSL_REQUIRE(Btot > 0 && P > 0 && eps > 0.f && iters >= 0 && keep_last > 0 && keep_last <= Btot,
                AVSSL_ERR_INVALID_ARGUMENT, "sinkhorn: bad arguments Btot=%d P=%d keep=%d", Btot, P, keep_last);
  const int grid = sink_grid(Btot);
  AVSSL_REQUIRE(grid > 0, AVSSL_ERR_CUDA, "sinkhorn: no CUDA device (there is no CPU fallback)");
  AVSSL_REQUIRE(workspace_bytes >= avssl_sinkhorn_workspace_bytes(Btot, P), AVSSL_ERR_WORKSPACE, "sinkhorn: workspace too small");
  SinkArgs a;
  a.scores = scores;
  a.Btot = Btot;
  a.P = P;
  a.inv_eps = 1.f / eps;
  a.iters = iters;
  a.keep_last = keep_last;
  a.out = codes_out;
  if (sinkhorn_try_cluster(a, static_cast<cudaStream_t>(stream))) return AVSSL_OK;
  char* w = static_cast<char*>(workspace);
  a.bar = reinterpret_cast<unsigned*>(w);
  a.g_alpha = reinterpret_cast<float*>(w + 256);
  a.g_part = reinterpret_cast<float*>(w + 256 + ((4 * (size_t)P + 255) / 256) * 256);
  a.spc = (Btot + grid - 1) / grid;
  const int grid_used = (Btot + a.spc - 1) / a.spc;
  const size_t fixed = sizeof(float) * ((size_t)P + ((a.spc + 3) & ~3));
  const size_t slab = sizeof(float) * (size_t)a.spc * P;
  a.slab_in_smem = (fixed + slab <= 200 * 1024) ? 1 : 0;
  const size_t smem = fixed + (a.slab_in_smem ? slab : 0);
  AVSSL_REQUIRE(fixed <= 200 * 1024, AVSSL_ERR_UNSUPPORTED, "sinkhorn: P=%d too large", P);
  AVSSL_CUDA_OK(cudaFuncSetAttribute(sinkhorn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(200 * 1024)));
  void* args[] = {&a};
  AVSSL_CUDA_OK(cudaLaunchCooperativeKernel((void*)sinkhorn_kernel, dim3(grid_used), dim3(256), args, smem,
                                            static_cast<cudaStream_t>(stream)));
  return AVSSL_OK;
}

extern "C" size_t avssl_swav_ce_workspace_bytes(int n_rows) { return 256 + 4 * (size_t)(n_rows > 0 ? n_rows : 0); }

extern "C" int avssl_swav_ce_fwd_bwd(const float* scores, const float* codes, int n_crops, int n_assign, int bs, int P,
                                     float T, const float* pair_w_host, float* loss_out, float* dscores_out,
                                     void* workspace, size_t workspace_bytes, void* stream) {
  AVSSL_REQUIRE(scores && codes && pair_w_host && loss_out && workspace, AVSSL_ERR_INVALID_ARGUMENT, "swav_ce: null pointer");
  AVSSL_REQUIRE(n_crops > 0 && n_crops <= kMaxCrops && n_assign > 0 && n_assign <= kMaxAssign && bs > 0 && P > 0 && T > 0.f,
                AVSSL_ERR_INVALID_ARGUMENT, "swav_ce: bad sizes (n_crops <= %d, n_assign <= %d)", kMaxCrops, kMaxAssign);
  AVSSL_REQUIRE(workspace_bytes >= avssl_swav_ce_workspace_bytes(n_crops * bs), AVSSL_ERR_WORKSPACE, "swav_ce: workspace too small");
  SwavCeArgs a;
  a.scores = scores;
  a.codes = codes;
  a.n_crops = n_crops;
  a.n_assign = n_assign;
  a.bs = bs;
  a.P = P;
  a.inv_T = 1.f / T;
  for (int i = 0; i < kMaxAssign * kMaxCrops; ++i) a.w[i] = i < n_assign * n_crops ? pair_w_host[i] : 0.f;
  a.loss_out = loss_out;
  a.dscores = dscores_out;
  a.counter = static_cast<unsigned*>(workspace);
  a.row_loss = reinterpret_cast<float*>(static_cast<char*>(workspace) + 256);
  const bool reg_path = P % 4 == 0 && P <= 256 * 4 * kCeVec &&
                        ((reinterpret_cast<uintptr_t>(scores) | reinterpret_cast<uintptr_t>(codes) |
                          reinterpret_cast<uintptr_t>(dscores_out)) & 15u) == 0;
  // more than two code sets (never produced by the reference: two global crops) take the generic kernel
  // AVSSL_SWAV_CE_KERNEL=reg: developer knob for A/B measurements (tools/next_bench.py); default: sample-major when large
  const char* knob = getenv("AVSSL_SWAV_CE_KERNEL");
  const bool big = reg_path && n_assign <= 2 && n_crops * bs >= 2 * sm_count();
  const size_t sample_smem = (size_t)(2 + 2) * P * 4;  // two code rows + two score stages
  if (big && n_crops >= 2 && sample_smem <= 200 * 1024 && !(knob && !strcmp(knob, "reg"))) {
    // code rows resident per CTA, the sample's score rows streamed; enough CTAs for ~3 per SM
    static unsigned long long configured = 0ull;
    if (first_use_on_device(configured))
      AVSSL_CUDA_OK(cudaFuncSetAttribute(swav_ce_sample_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    int groups = (3 * sm_count() + bs - 1) / bs;
    if (const char* gk = getenv("AVSSL_SWAV_CE_GROUPS")) groups = atoi(gk);  // developer knob (tools/k11_groups.py)
    if (groups < 1) groups = 1;
    if (groups > n_crops) groups = n_crops;
    const int per = (n_crops + groups - 1) / groups;
    groups = (n_crops + per - 1) / per;
    swav_ce_sample_kernel<2><<<bs * groups, 256, sample_smem, static_cast<cudaStream_t>(stream)>>>(a, per);
  } else if (reg_path && n_assign <= 2)
    swav_ce_reg_kernel<2><<<n_crops * bs, 256, 0, static_cast<cudaStream_t>(stream)>>>(a);
  else
    swav_ce_kernel<<<n_crops * bs, 256, 0, static_cast<cudaStream_t>(stream)>>>(a);
  AVSSL_LAUNCH_OK("swav_ce_kernel");
  return AVSSL_OK;
}
