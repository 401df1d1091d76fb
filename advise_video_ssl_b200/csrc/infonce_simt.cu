// K2+K3, CUDA-core implementation: exact-fp32 split kernel + combine kernel.
//
// Replaces models/contrastive.py:462,486-500 and models/losses.py:20-25.
// One pass over the queue: every CTA streams a contiguous slice of queue rows in
// 64-row tiles (cp.async double buffer), computes S = q.tile^T with fp32 FMAs,
// keeps a running (max, sum, sum_j p_ij queue_j) per query row (online softmax,
// log2 domain), and writes one partial per (split, row).  The combine kernel merges
// the partials, adds the positive logits and emits loss, lse, df (the gradient
// through the l2-normalisation), q and logits column 0.
//
// This kernel is the fp32 reference implementation on the device (any D % 4 == 0,
// D <= 256, any B, any K) and the fallback for shapes the tcgen05 kernel does not
// take.  It is FFMA-bound (2.1 GFLOP at cfg1), not HBM-bound.
#include "infonce.cuh"
#include "simt_tile.cuh"

namespace avssl {

template <int DP>
__global__ void __launch_bounds__(kSimtThreads, 1) infonce_simt_kernel(const InfoNceParams p) {
  constexpr int KS = DP + 4;
  constexpr int CC = DP / 64;
  extern __shared__ __align__(16) float smem[];
  float* qs = smem;                       // [64][KS]
  float* ks0 = qs + kTileI * KS;          // [2][64][KS]
  float* ps = ks0 + 2 * kTileJ * KS;      // [64][kPsStride]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int split = blockIdx.x;
  const int i_base = blockIdx.y * kTileI;
  const int j_begin = split * p.rows_per_split;
  const int j_end = min(p.K, j_begin + p.rows_per_split);
  const int n_tiles = (j_end - j_begin + kTileJ - 1) / kTileJ;
  const int D = p.D;

  if (n_tiles > 0) {
    load_tile<DP>(ks0, p.queue, D, j_begin, j_end);
    cp_async_commit();
  }

  // prologue: q = f / ||f|| for this CTA's 64 rows (one warp per row)
  for (int r = warp; r < kTileI; r += kSimtThreads / 32) {
    const int i = i_base + r;
    if (i < p.B) {
      const float* f = p.feat_q + (size_t)i * D;
      const float nrm = warp_row_norm(f, D, lane);
      for (int c = lane; c < DP; c += 32) qs[r * KS + c] = (c < D) ? f[c] / nrm : 0.f;
    } else {
      for (int c = lane; c < DP; c += 32) qs[r * KS + c] = 0.f;
    }
  }

  const int ty = tid >> 4, tx = tid & 15;
  const float scale2 = p.inv_T * kLog2e;
  float m_run[4], l_run[4];
  float4 acc[4][CC];
#pragma unroll
  for (int ii = 0; ii < 4; ++ii) {
    m_run[ii] = -INFINITY;
    l_run[ii] = 0.f;
#pragma unroll
    for (int cc = 0; cc < CC; ++cc) acc[ii][cc] = make_float4(0.f, 0.f, 0.f, 0.f);
  }

  for (int t = 0; t < n_tiles; ++t) {
    float* ks = ks0 + (t & 1) * kTileJ * KS;
    if (t + 1 < n_tiles) {
      load_tile<DP>(ks0 + ((t + 1) & 1) * kTileJ * KS, p.queue, D, j_begin + (t + 1) * kTileJ, j_end);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();

    // ---- S = q . tile^T : thread owns rows ty*4+ii, columns tx+16*jj
    float s[4][4];
    simt_s_tile<DP>(qs, ks, ty, tx, s);

    const int jt0 = j_begin + t * kTileJ;
    // optional logits materialisation (models/contrastive.py:498: logits / T)
    if (p.logits_out) {
#pragma unroll
      for (int ii = 0; ii < 4; ++ii) {
        const int i = i_base + ty * 4 + ii;
        if (i < p.B) {
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            const int j = jt0 + tx + 16 * jj;
            if (j < j_end) {
              const float v = s[ii][jj] * p.inv_T;
              for (int k = 0; k < p.n_keys; ++k)
                p.logits_out[((size_t)k * p.B + i) * (size_t)(p.K + 1) + 1 + j] = v;
            }
          }
        }
      }
    }

    // ---- online softmax (log2 domain); 16 lanes share a row group
    float alpha[4];
#pragma unroll
    for (int ii = 0; ii < 4; ++ii) {
      float tmax = -INFINITY;
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const bool valid = (jt0 + tx + 16 * jj) < j_end;
        s[ii][jj] = valid ? s[ii][jj] * scale2 : -INFINITY;
        tmax = fmaxf(tmax, s[ii][jj]);
      }
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, o));
      const float m_new = fmaxf(m_run[ii], tmax);
      alpha[ii] = exp2f(m_run[ii] - m_new);  // 0 on the first tile (m_run = -inf)
      float psum = 0.f;
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const float pv = exp2f(s[ii][jj] - m_new);
        psum += pv;
        ps[(ty * 4 + ii) * kPsStride + tx + 16 * jj] = pv;
      }
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) psum += __shfl_xor_sync(0xffffffffu, psum, o);
      l_run[ii] = l_run[ii] * alpha[ii] + psum;
      m_run[ii] = m_new;
    }
    __syncthreads();

    // ---- acc = acc*alpha + P . tile : thread owns rows ty*4+ii, columns tx*4 + 64*cc
#pragma unroll
    for (int ii = 0; ii < 4; ++ii)
#pragma unroll
      for (int cc = 0; cc < CC; ++cc) {
        acc[ii][cc].x *= alpha[ii];
        acc[ii][cc].y *= alpha[ii];
        acc[ii][cc].z *= alpha[ii];
        acc[ii][cc].w *= alpha[ii];
      }
    simt_pv_tile<DP>(ps, ks, ty, tx, acc);
    __syncthreads();
  }

  // ---- partials
#pragma unroll
  for (int ii = 0; ii < 4; ++ii) {
    const int i = i_base + ty * 4 + ii;
    if (i >= p.B) continue;
    const size_t row = (size_t)split * p.B + i;
    if (tx == 0) {
      p.part_m[row] = m_run[ii];
      p.part_l[row] = l_run[ii];
    }
#pragma unroll
    for (int cc = 0; cc < CC; ++cc) {
      const int c = tx * 4 + 64 * cc;
      if (c < D) *reinterpret_cast<float4*>(p.part_acc + row * D + c) = acc[ii][cc];
    }
  }
}

// --------------------------------------------------------------------- combine
// One CTA per query row.  Deterministic: partials are merged in a fixed order and the
// mean over rows is taken by the last CTA in row order.  512 threads = 16 groups of 32
// lanes; a group reads one split-partial row as float4 per lane, 4 rows in flight per
// group, so ~32 KB of independent loads are outstanding per CTA (the partials sit in L2:
// ~5 MB at 147 splits).
constexpr int kCombineThreads = 512;
constexpr int kCombineCols = 128;
constexpr int kCombineGroups = kCombineThreads / 32;
constexpr int kMaxSplits = 1024;

__global__ void __launch_bounds__(kCombineThreads) infonce_combine_kernel(const InfoNceParams p) {
  __shared__ float s_w[kMaxSplits];
  __shared__ float s_red[32];
  __shared__ float s_bcast[4];
  __shared__ __align__(16) float s_acc[kCombineGroups][2 * kCombineCols];
  __shared__ unsigned s_is_last;
  const int i = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int col = tid % kCombineCols;
  const int grp = tid >> 5;  // split group (warp) for the accumulator merge
  const int D = p.D, B = p.B, S = p.n_splits;
  constexpr int MAXC = 2;  // D <= 256
  const float* f = p.feat_q + (size_t)i * D;

  if (warp == 0) {
    const float nrm = warp_row_norm(f, D, lane);
    if (lane == 0) s_bcast[0] = nrm;
  }
  // global max over the splits
  float mloc = -INFINITY;
  for (int s = tid; s < S; s += kCombineThreads) mloc = fmaxf(mloc, p.part_m[(size_t)s * B + i]);
  mloc = warp_max(mloc);
  if (lane == 0) s_red[warp] = mloc;
  __syncthreads();
  float M = s_red[0];
  for (int w = 1; w < kCombineThreads / 32; ++w) M = fmaxf(M, s_red[w]);
  const float nrm = s_bcast[0];
  __syncthreads();
  float lloc = 0.f;
  for (int s = tid; s < S; s += kCombineThreads) {
    const float w = exp2f(p.part_m[(size_t)s * B + i] - M);
    s_w[s] = w;
    lloc += w * p.part_l[(size_t)s * B + i];
  }
  const float L = block_sum(lloc, s_red);  // (order fixed by the launch geometry)
  __syncthreads();

  // merge the accumulators: warp g takes splits g, g+G, ...; lane l owns columns 4l..4l+3
  // (+128 for D > 128), four split rows in flight
  {
    const size_t stride = (size_t)B * D;
#pragma unroll
    for (int u = 0; u < MAXC; ++u) {
      const int c = lane * 4 + u * kCombineCols;
      float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0, a2 = a0, a3 = a0;
      if (c < D) {
        const float* base = p.part_acc + (size_t)i * D + c;
        auto ld = [&](int s) { return __ldcg(reinterpret_cast<const float4*>(base + (size_t)s * stride)); };
        auto fma4 = [](float w, const float4& v, float4& a) {
          a.x = fmaf(w, v.x, a.x);
          a.y = fmaf(w, v.y, a.y);
          a.z = fmaf(w, v.z, a.z);
          a.w = fmaf(w, v.w, a.w);
        };
        int s = grp;
        for (; s + 3 * kCombineGroups < S; s += 4 * kCombineGroups) {
          const float4 v0 = ld(s), v1 = ld(s + kCombineGroups), v2 = ld(s + 2 * kCombineGroups),
                       v3 = ld(s + 3 * kCombineGroups);
          fma4(s_w[s], v0, a0);
          fma4(s_w[s + kCombineGroups], v1, a1);
          fma4(s_w[s + 2 * kCombineGroups], v2, a2);
          fma4(s_w[s + 3 * kCombineGroups], v3, a3);
        }
        for (; s < S; s += kCombineGroups) fma4(s_w[s], ld(s), a0);
      }
      float4 r;
      r.x = (a0.x + a1.x) + (a2.x + a3.x);
      r.y = (a0.y + a1.y) + (a2.y + a3.y);
      r.z = (a0.z + a1.z) + (a2.z + a3.z);
      r.w = (a0.w + a1.w) + (a2.w + a3.w);
      *reinterpret_cast<float4*>(&s_acc[grp][c]) = r;
    }
  }
  __syncthreads();

  const bool owner = tid < kCombineCols;  // these 128 threads own the columns from here on
  float q[MAXC], acc[MAXC], dq[MAXC];
#pragma unroll
  for (int u = 0; u < MAXC; ++u) {
    const int c = col + u * kCombineCols;
    q[u] = acc[u] = dq[u] = 0.f;
    if (owner && c < D) {
      q[u] = f[c] / nrm;
      float a = s_acc[0][c];
#pragma unroll
      for (int g = 1; g < kCombineGroups; ++g) a += s_acc[g][c];
      acc[u] = a;
      p.q_out[(size_t)i * D + c] = q[u];
    }
  }

  const int n_rows = p.n_keys * B;
  const float gscale = p.inv_T / (float)n_rows;
  for (int k = 0; k < p.n_keys; ++k) {
    const float* key = p.keys[k] + (size_t)i * D;
    float kv[MAXC], dot = 0.f;
#pragma unroll
    for (int u = 0; u < MAXC; ++u) {
      const int c = col + u * kCombineCols;
      kv[u] = (owner && c < D) ? key[c] : 0.f;
      dot = fmaf(q[u], kv[u], dot);
    }
    dot = block_sum(dot, s_red);
    const float s0 = dot * p.inv_T;         // positive logit (column 0)
    const float s0_2 = s0 * kLog2e;
    const float Mk = fmaxf(M, s0_2);
    const float e0 = exp2f(s0_2 - Mk);
    const float wq = exp2f(M - Mk);
    const float Z = e0 + L * wq;
    const float lse = (Mk + log2f(Z)) * kLn2;
    const float p0 = e0 / Z;
    const float pq = wq / Z;  // scales acc to sum_j p_kij queue_j
#pragma unroll
    for (int u = 0; u < MAXC; ++u) dq[u] += (pq * acc[u] + p0 * kv[u] - kv[u]) * gscale;
    if (tid == 0) {
      p.row_loss[(size_t)k * B + i] = lse - s0;
      if (p.row_lse_out) p.row_lse_out[(size_t)k * B + i] = lse;
      if (p.logits_out) p.logits_out[((size_t)k * B + i) * (size_t)(p.K + 1)] = s0;
    }
  }
  // gradient through the normalisation: df = (dq - (dq.q) q) / ||f||
  float dd = 0.f;
#pragma unroll
  for (int u = 0; u < MAXC; ++u) dd = fmaf(dq[u], q[u], dd);
  dd = block_sum(dd, s_red);
#pragma unroll
  for (int u = 0; u < MAXC; ++u) {
    const int c = col + u * kCombineCols;
    if (owner && c < D) p.dfeat_out[(size_t)i * D + c] = (dq[u] - dd * q[u]) / nrm;
  }

  // mean over all logits rows, by the last CTA, in row order
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    s_is_last = (atomicAdd(p.counter, 1u) == gridDim.x - 1) ? 1u : 0u;
  }
  __syncthreads();
  if (s_is_last) {
    __threadfence();
    float tot = 0.f;
    for (int r = tid; r < n_rows; r += kCombineThreads) tot += reinterpret_cast<volatile float*>(p.row_loss)[r];
    tot = block_sum(tot, s_red);
    if (tid == 0) {
      *p.loss_out = tot / (float)n_rows;
      *p.counter = 0u;
    }
  }
}

template <int DP>
static int launch_simt_dp(const InfoNceParams& p, cudaStream_t s) {
  const size_t smem = sizeof(float) * ((size_t)kTileI * (DP + 4) + 2 * kTileJ * (DP + 4) + kTileI * kPsStride);
  static bool configured = false;
  if (!configured) {
    AVSSL_CUDA_OK(cudaFuncSetAttribute(infonce_simt_kernel<DP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  dim3 grid(p.n_splits, (p.B + kTileI - 1) / kTileI);
  infonce_simt_kernel<DP><<<grid, kSimtThreads, smem, s>>>(p);
  AVSSL_LAUNCH_OK("infonce_simt_kernel");
  return AVSSL_OK;
}

int launch_infonce_simt(const InfoNceParams& p, cudaStream_t s) {
  AVSSL_REQUIRE(p.D % 4 == 0 && p.D >= 4 && p.D <= 256, AVSSL_ERR_UNSUPPORTED,
                "moco_infonce: D=%d unsupported (need D %% 4 == 0 and D <= 256)", p.D);
  if (p.D <= 64) return launch_simt_dp<64>(p, s);
  if (p.D <= 128) return launch_simt_dp<128>(p, s);
  return launch_simt_dp<256>(p, s);
}

int launch_infonce_combine(const InfoNceParams& p, cudaStream_t s) {
  AVSSL_REQUIRE(p.n_splits <= kMaxSplits, AVSSL_ERR_INVALID_ARGUMENT, "moco_infonce: too many splits");
  infonce_combine_kernel<<<p.B, kCombineThreads, 0, s>>>(p);
  AVSSL_LAUNCH_OK("infonce_combine_kernel");
  return AVSSL_OK;
}

}  // namespace avssl
