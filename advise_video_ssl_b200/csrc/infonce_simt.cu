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
#include "infonce_combine.cuh"
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
// One CTA per query row (infonce_combine.cuh); the mean over rows is taken by the last
// CTA to finish, in row order.
constexpr int kCombineThreads = 512;

__global__ void __launch_bounds__(kCombineThreads) infonce_combine_kernel(const InfoNceParams p) {
  __shared__ CombineSmem<kCombineThreads> sm;
  infonce_combine_row<kCombineThreads, 2>(p, blockIdx.x, sm, p.keys[0] + (size_t)blockIdx.x * p.D);
  infonce_finish<kCombineThreads>(p, gridDim.x, sm);
}

template <int DP>
static int launch_simt_dp(const InfoNceParams& p, cudaStream_t s) {
  const size_t smem = sizeof(float) * ((size_t)kTileI * (DP + 4) + 2 * kTileJ * (DP + 4) + kTileI * kPsStride);
  static unsigned long long configured = 0ull;  // device ordinals already set up
  if (first_use_on_device(configured)) {
    AVSSL_CUDA_OK(cudaFuncSetAttribute(infonce_simt_kernel<DP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
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
