// Merge step of the split MoCo InfoNCE kernels (K2+K3): one CTA merges the per-split
// partials (m, l, acc) of ONE query row, adds the positive logits and emits loss terms,
// lse, q, df (the gradient through the l2-normalisation) and logits column 0.
//
// Shared by the stand-alone combine kernel (CUDA-core path, infonce_simt.cu) and the
// tcgen05 kernel, which runs it after a grid-wide barrier inside the same launch
// (infonce_tc.cu).  Deterministic: partials are merged in an order fixed by the launch
// geometry, never with floating-point atomics.
//
// Replaces models/contrastive.py:462 (Normalize), :490-500 (positive logits, cat, /T) and
// models/losses.py:20-25 (cross-entropy against class 0, mean over rows).
#pragma once
#include "infonce.cuh"
#include "tc_trace.cuh"

namespace avssl {

constexpr int kCombineCols = 128;  // columns owned by the first 128 threads (x2 for D > 128)
constexpr int kMaxSplits = 1024;

template <int kThreads>
struct CombineSmem {
  float w[kMaxSplits];
  float red[32];
  float bcast[4];
  __align__(16) float acc[kThreads / 32][2 * kCombineCols];
  unsigned is_last;
  int peer_slot;
  long long enq_ptr;
};

__device__ __forceinline__ float fast_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float fast_lg2(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// All kThreads threads of the CTA call this with the same row i.  MAXC = ceil(D / 128).
// key0_row: row i of the first key tensor (p.keys[0] + i * D, or a row of the peer exchange
// buffer -- read through L2, it may have been written by another GPU).
//
// Latency matters more than bandwidth here (the partials sit in L2, ~75 KB per row at 148
// splits), and the code runs once per launch with a cold instruction cache, so it is kept
// small: every global load of the row -- part_m, part_l, the first kWide*G rows of part_acc,
// f and the first key -- is issued before the first dependent instruction (one L2 round
// trip), rare cases (more splits than kWide*G, D > 128, several keys) go through compact
// non-unrolled loops, and transcendental functions use the hardware approximations
// (relative error ~2^-22, far inside the fp32 tolerance of the loss).
template <int kThreads, int MAXC>
__device__ __forceinline__ void infonce_combine_row(const InfoNceParams& p, int i, CombineSmem<kThreads>& sm,
                                                    const float* __restrict__ key0_row) {
  constexpr int kGroups = kThreads / 32;
  constexpr int kWide = 11;  // split rows in flight per lane (11 x 14 warps covers 148 splits)
  constexpr int kMl = (kMaxSplits + kThreads - 1) / kThreads;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int col = tid % kCombineCols;
  const int D = p.D, B = p.B, S = p.n_splits;
  const float* f = p.feat_q + (size_t)i * D;
  const size_t stride = (size_t)B * D;
  const bool owner = tid < kCombineCols;  // these 128 threads own the columns in the second half

  __syncthreads();  // previous row's shared state is dead
  if (tid == 0) TC_TRACE(10, 0);

  // ---- every load up front
  float ms[kMl], ls[kMl];
#pragma unroll
  for (int k = 0; k < kMl; ++k) {
    const int s = tid + k * kThreads;
    ms[k] = s < S ? __ldcg(p.part_m + (size_t)s * B + i) : -INFINITY;
    ls[k] = s < S ? __ldcg(p.part_l + (size_t)s * B + i) : 0.f;
  }
  float4 v0[kWide];  // first chunk of this warp's split rows, columns 4*lane..4*lane+3
  {
    const bool in = lane * 4 < D;
    const float* src = p.part_acc + (size_t)i * D + lane * 4 + (size_t)warp * stride;
#pragma unroll
    for (int k = 0; k < kWide; ++k) {
      v0[k] = (in && warp + k * kGroups < S) ? __ldcg(reinterpret_cast<const float4*>(src)) : make_float4(0.f, 0.f, 0.f, 0.f);
      src += (size_t)kGroups * stride;
    }
  }
  float fv[MAXC], kv0[MAXC];  // f and the first key (usually the only one)
#pragma unroll
  for (int u = 0; u < MAXC; ++u) {
    const int c = col + u * kCombineCols;
    fv[u] = (owner && c < D) ? f[c] : 0.f;
    kv0[u] = (owner && c < D) ? __ldcg(key0_row + c) : 0.f;
  }
  if (warp == kGroups - 1) {  // the last warp holds the fewest partial rows
    const float nrm = warp_row_norm(f, D, lane);
    if (lane == 0) sm.bcast[0] = nrm;
  } else if (warp == kGroups - 2 && p.keys_raw) {  // Normalize of the raw key row, same arithmetic as l2norm_fwd_kernel
    const float knrm = warp_row_norm(key0_row, D, lane);
    if (lane == 0) sm.bcast[1] = knrm;
  }

  // ---- global max over the splits, weights, L
  if (tid == 0) TC_TRACE(10, 1);
  float mloc = ms[0];
#pragma unroll
  for (int k = 1; k < kMl; ++k) mloc = fmaxf(mloc, ms[k]);
  mloc = warp_max(mloc);
  if (lane == 0) sm.red[warp] = mloc;
  __syncthreads();
  if (tid == 0) TC_TRACE(10, 2);
  float M = sm.red[0];
#pragma unroll
  for (int w = 1; w < kGroups; ++w) M = fmaxf(M, sm.red[w]);
  const float nrm = sm.bcast[0];
  if (p.keys_raw) {
    const float knrm = sm.bcast[1];
#pragma unroll
    for (int u = 0; u < MAXC; ++u) kv0[u] = kv0[u] / knrm;
  }
  float lloc = 0.f;
#pragma unroll
  for (int k = 0; k < kMl; ++k) {
    const int s = tid + k * kThreads;
    if (s < S) {
      const float w = fast_ex2(ms[k] - M);
      sm.w[s] = w;
      lloc = fmaf(w, ls[k], lloc);
    }
  }
  const float L = block_sum(lloc, sm.red);  // (order fixed by the launch geometry; syncs publish sm.w)
  if (tid == 0) TC_TRACE(10, 3);

  // ---- merge the accumulators: warp g takes splits g, g+G, ...; lane l owns columns 4l..4l+3
  {
    float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
#pragma unroll
    for (int k = 0; k < kWide; ++k) {
      const int s = warp + k * kGroups;
      const float w = s < S ? sm.w[s] : 0.f;
      float4& t = (k & 1) ? a1 : a0;
      t.x = fmaf(w, v0[k].x, t.x);
      t.y = fmaf(w, v0[k].y, t.y);
      t.z = fmaf(w, v0[k].z, t.z);
      t.w = fmaf(w, v0[k].w, t.w);
    }
    // more splits than kWide * G (never on a 148-SM part): plain loop
#pragma unroll 1
    for (int s = warp + kWide * kGroups; s < S; s += kGroups) {
      if (lane * 4 < D) {
        const float4 v = __ldcg(reinterpret_cast<const float4*>(p.part_acc + (size_t)s * stride + (size_t)i * D + lane * 4));
        const float w = sm.w[s];
        a0.x = fmaf(w, v.x, a0.x);
        a0.y = fmaf(w, v.y, a0.y);
        a0.z = fmaf(w, v.z, a0.z);
        a0.w = fmaf(w, v.w, a0.w);
      }
    }
    *reinterpret_cast<float4*>(&sm.acc[warp][lane * 4]) = make_float4(a0.x + a1.x, a0.y + a1.y, a0.z + a1.z, a0.w + a1.w);
    if (MAXC > 1) {  // columns 128..255 (CUDA-core path only)
      const int c = lane * 4 + kCombineCols;
      float4 b0 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c < D) {
#pragma unroll 1
        for (int s = warp; s < S; s += kGroups) {
          const float4 v = __ldcg(reinterpret_cast<const float4*>(p.part_acc + (size_t)s * stride + (size_t)i * D + c));
          const float w = sm.w[s];
          b0.x = fmaf(w, v.x, b0.x);
          b0.y = fmaf(w, v.y, b0.y);
          b0.z = fmaf(w, v.z, b0.z);
          b0.w = fmaf(w, v.w, b0.w);
        }
      }
      *reinterpret_cast<float4*>(&sm.acc[warp][c]) = b0;
    }
  }
  __syncthreads();
  if (tid == 0) TC_TRACE(10, 4);

  float q[MAXC], acc[MAXC], dq[MAXC];
#pragma unroll
  for (int u = 0; u < MAXC; ++u) {
    const int c = col + u * kCombineCols;
    q[u] = acc[u] = dq[u] = 0.f;
    if (owner && c < D) {
      q[u] = fv[u] / nrm;
      float a = sm.acc[0][c];
#pragma unroll
      for (int g = 1; g < kGroups; ++g) a += sm.acc[g][c];
      acc[u] = a;
      p.q_out[(size_t)i * D + c] = q[u];
    }
  }

  const int n_rows = p.n_keys * B;
  const float gscale = p.inv_T / (float)n_rows;
#pragma unroll 1
  for (int k = 0; k < p.n_keys; ++k) {
    float kv[MAXC], dot = 0.f;
#pragma unroll
    for (int u = 0; u < MAXC; ++u) {
      const int c = col + u * kCombineCols;
      kv[u] = k == 0 ? kv0[u] : ((owner && c < D) ? p.keys[k][(size_t)i * D + c] : 0.f);
      dot = fmaf(q[u], kv[u], dot);
    }
    dot = block_sum(dot, sm.red);
    const float s0 = dot * p.inv_T;  // positive logit (column 0)
    const float s0_2 = s0 * kLog2e;
    const float Mk = fmaxf(M, s0_2);
    const float e0 = fast_ex2(s0_2 - Mk);
    const float wq = fast_ex2(M - Mk);
    const float Z = fmaf(L, wq, e0);
    const float lse = (Mk + fast_lg2(Z)) * kLn2;
    const float rz = __fdividef(1.f, Z);
    const float p0 = e0 * rz;
    const float pq = wq * rz;  // scales acc to sum_j p_kij queue_j
#pragma unroll
    for (int u = 0; u < MAXC; ++u) dq[u] += (pq * acc[u] + p0 * kv[u] - kv[u]) * gscale;
    if (tid == 0) {
      p.row_loss[(size_t)k * B + i] = lse - s0;
      if (p.row_lse_out) p.row_lse_out[(size_t)k * B + i] = lse;
      if (p.logits_out) p.logits_out[((size_t)k * B + i) * (size_t)(p.K + 1)] = s0;
    }
  }
  if (tid == 0) TC_TRACE(10, 5);
  // gradient through the normalisation: df = (dq - (dq.q) q) / ||f||
  float dd = 0.f;
#pragma unroll
  for (int u = 0; u < MAXC; ++u) dd = fmaf(dq[u], q[u], dd);
  dd = block_sum(dd, sm.red);
  const float rn = 1.f / nrm;
#pragma unroll
  for (int u = 0; u < MAXC; ++u) {
    const int c = col + u * kCombineCols;
    if (owner && c < D) p.dfeat_out[(size_t)i * D + c] = (dq[u] - dd * q[u]) * rn;
  }
  if (tid == 0) TC_TRACE(10, 6);
}

// Mean over all logits rows in row order, by the CTA that arrives last at p.counter[0]
// (`n_arrivals` CTAs call this once each, after their rows are merged).  Returns true in the
// last CTA (all threads), after loss_out is written and the counter is reset.
template <int kThreads>
__device__ __forceinline__ bool infonce_finish(const InfoNceParams& p, unsigned n_arrivals, CombineSmem<kThreads>& sm) {
  const int tid = threadIdx.x;
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    sm.is_last = (atomicAdd(p.counter, 1u) == n_arrivals - 1) ? 1u : 0u;
  }
  __syncthreads();
  if (!sm.is_last) return false;
  __threadfence();
  const int n_rows = p.n_keys * p.B;
  float tot = 0.f;
  for (int r = tid; r < n_rows; r += kThreads) tot += __ldcg(p.row_loss + r);
  tot = block_sum(tot, sm.red);
  if (tid == 0) {
    *p.loss_out = tot / (float)n_rows;
    *p.counter = 0u;
  }
  return true;
}

}  // namespace avssl
