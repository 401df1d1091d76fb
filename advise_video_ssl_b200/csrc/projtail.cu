// Projection tail: the LAST Linear of the projection MLP with the head's Normalize
// as its epilogue, and the matching backward.
//   MLPHead.projection[-1]   models/head_helper.py:52-58   y = x W^T + b      (x: [B, mlp_dim] after BN + ReLU)
//   Normalize                models/contrastive.py:923-934 q = y / ||y||      (:462 / :350 / :757 callers)
// The raw projection y never reaches HBM: forward emits the unit rows q and ||y||; backward takes dL/dq, applies the
// gradient of the normalisation where it reads it, and produces dx, dW, db in one launch.
//
// Shapes on the path: B = 64..512 rows, Kin = SSL.MLP_DIM (2048), Dout = CONTRASTIVE.DIM (128 / 256): 33 MFLOP and
// 1.5 MB at cfg2 -- latency-bound, exact fp32 on the CUDA cores (the tensor pipe would need a 3-term split for fp32
// parity and its 128-row tile would be half empty).
//   forward : split-K over a thread-block cluster of 8 CTAs per 8-row block; the partial sums are combined through
//             distributed shared memory in rank order (deterministic), CTA r of the cluster finishes row r.
//   backward: one CTA per 16-column slice of Kin (dW and dx are both independent along Kin, so no reduction crosses
//             CTAs); dy = (g - (g.q) q) / ||y|| is recomputed per CTA from g, q (B x Dout is small).
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace avssl {
namespace {

constexpr int kPtThreads = 256;
constexpr int kPtCluster = 8;        // CTAs per cluster = K-slices = rows per row block
constexpr int kPtRows = kPtCluster;  // CTA r of the cluster finishes row r of the block
constexpr int kPtKT = 64;            // k-columns staged per step
constexpr int kPtWs = kPtKT + 4;     // padded row stride of the staged W tile (floats): conflict-free 128-bit reads

struct ProjFwdArgs {
  const float* x;     // [B, Kin]
  const float* W;     // [Dout, Kin]
  const float* bias;  // [Dout] or null
  float* q;           // [B, Dout]
  float* norm;        // [B] or null
  int B, Kin, Dout;
  float eps;
  int normalize;  // 0: plain Linear (q = y)
};

// CPT = output columns per thread (1: Dout <= 128, 2: Dout <= 256)
template <int CPT>
__global__ void __cluster_dims__(kPtCluster, 1, 1) __launch_bounds__(kPtThreads, 1) linear_l2norm_fwd_kernel(const ProjFwdArgs a) {
  extern __shared__ __align__(16) float smem[];
  cg::cluster_group cluster = cg::this_cluster();
  const int DP = CPT * 128;                   // padded column count
  float* W_s = smem;                          // [DP][kPtWs]
  float* x_s = W_s + DP * kPtWs;              // [kPtRows][kPtKT]
  float* part = x_s + kPtRows * kPtKT;        // [kPtRows][DP]  this CTA's partial sums (read by the whole cluster)
  float* row_s = part + kPtRows * DP;         // [DP]           the finished row of this CTA

  const int t = threadIdx.x;
  const unsigned cr = cluster.block_rank();   // K-slice of this CTA, and the row it finishes
  const int row0 = blockIdx.y * kPtRows;
  const int Kin = a.Kin, Dout = a.Dout;
  // K-slice [ks, ke): multiples of 4 so that every staged 128-bit load stays inside the slice
  const int per = ((Kin + kPtCluster - 1) / kPtCluster + 3) & ~3;
  const int ks = min((int)cr * per, Kin), ke = min(ks + per, Kin);
  const int n_steps = (ke - ks + kPtKT - 1) / kPtKT;

  const int c0 = t & 127, rg = t >> 7;  // columns c0 (+128), rows rg*4 .. rg*4+3 of the block
  float acc[4][CPT];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int j = 0; j < CPT; ++j) acc[r][j] = 0.f;

  // register staging: W tile = DP rows x 16 float4 -> DP*16/256 = 8*CPT float4 per thread; x tile = 8 x 16 float4
  constexpr int WV = 8 * CPT;
  float4 wreg[WV];
  float4 xreg;
  auto load_tile = [&](int step) {
    const int k0 = ks + step * kPtKT;
#pragma unroll
    for (int i = 0; i < WV; ++i) {
      const int idx = t + i * kPtThreads;  // float4 index in the tile
      const int c = idx >> 4, k = k0 + ((idx & 15) << 2);
      wreg[i] = (c < Dout && k < ke) ? __ldg(reinterpret_cast<const float4*>(a.W + (int64_t)c * Kin + k)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (t < kPtRows * 16) {
      const int r = t >> 4, k = k0 + ((t & 15) << 2);
      xreg = (row0 + r < a.B && k < ke) ? __ldg(reinterpret_cast<const float4*>(a.x + (int64_t)(row0 + r) * Kin + k)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  if (n_steps > 0) load_tile(0);
  for (int step = 0; step < n_steps; ++step) {
    __syncthreads();  // the previous tile is no longer read
#pragma unroll
    for (int i = 0; i < WV; ++i) {
      const int idx = t + i * kPtThreads;
      *reinterpret_cast<float4*>(W_s + (idx >> 4) * kPtWs + ((idx & 15) << 2)) = wreg[i];
    }
    if (t < kPtRows * 16) *reinterpret_cast<float4*>(x_s + (t >> 4) * kPtKT + ((t & 15) << 2)) = xreg;
    __syncthreads();
    if (step + 1 < n_steps) load_tile(step + 1);  // in flight under the arithmetic below
#pragma unroll 4
    for (int kk = 0; kk < kPtKT; kk += 4) {
      float4 w[CPT];
#pragma unroll
      for (int j = 0; j < CPT; ++j) w[j] = *reinterpret_cast<const float4*>(W_s + (c0 + 128 * j) * kPtWs + kk);
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const float4 xv = *reinterpret_cast<const float4*>(x_s + (rg * 4 + r) * kPtKT + kk);  // warp-wide broadcast
#pragma unroll
        for (int j = 0; j < CPT; ++j) {
          acc[r][j] = fmaf(xv.x, w[j].x, acc[r][j]);
          acc[r][j] = fmaf(xv.y, w[j].y, acc[r][j]);
          acc[r][j] = fmaf(xv.z, w[j].z, acc[r][j]);
          acc[r][j] = fmaf(xv.w, w[j].w, acc[r][j]);
        }
      }
    }
  }
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int j = 0; j < CPT; ++j) part[(rg * 4 + r) * DP + c0 + 128 * j] = acc[r][j];
  cluster.sync();

  // CTA `cr` finishes row `cr`: K-slices added in rank order, then the bias (y = x W^T + b as nn.Linear states it)
  const int row = row0 + (int)cr;
  for (int c = t; c < DP; c += kPtThreads) {
    float v = 0.f;
#pragma unroll
    for (int g = 0; g < kPtCluster; ++g) v += cluster.map_shared_rank(part, g)[cr * DP + c];
    if (c < Dout && a.bias) v += a.bias[c];
    row_s[c] = v;
  }
  __syncthreads();
  if (t < 32 && row < a.B) {
    // Normalize with the arithmetic of l2norm_fwd_kernel (rowops.cu): same summation order, true division
    float den = 1.f, nrm = 0.f;
    if (a.normalize) {
      nrm = sqrtf(row_sumsq(row_s, Dout, t));
      den = fmaxf(nrm, a.eps);
    }
    for (int c = t; c < Dout; c += 32) a.q[(int64_t)row * Dout + c] = a.normalize ? row_s[c] / den : row_s[c];
    if (a.norm && t == 0) a.norm[row] = nrm;
  }
  cluster.sync();  // no CTA leaves while another one may still read its partial sums
}

// ------------------------------------------------------------------------------------------------ backward
constexpr int kPbKS = 16;    // Kin columns per CTA
constexpr int kPbRows = 64;  // rows per pass

struct ProjBwdArgs {
  const float* x;     // [B, Kin]
  const float* W;     // [Dout, Kin]
  const float* q;     // [B, Dout]   forward output
  const float* norm;  // [B]         ||y|| (unused when normalize == 0)
  const float* g;     // [B, Dout]   dL/dq
  float* dx;          // [B, Kin] or null
  float* dW;          // [Dout, Kin] or null
  float* db;          // [Dout] or null
  int B, Kin, Dout;
  float eps;
  int normalize;
};

template <int CPT>
__global__ void __launch_bounds__(kPtThreads, 1) linear_l2norm_bwd_kernel(const ProjBwdArgs a) {
  extern __shared__ __align__(16) float smem[];
  const int DP = CPT * 128;
  const int DYS = DP + 4;                // row stride of dy_s (floats)
  float* dy_s = smem;                    // [kPbRows][DYS]
  float* x_s = dy_s + kPbRows * DYS;     // [kPbRows][kPbKS]
  float* W_s = x_s + kPbRows * kPbKS;    // [DP][kPbKS]

  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int Kin = a.Kin, Dout = a.Dout;
  const int k0 = blockIdx.x * kPbKS;  // Kin % 4 == 0: every 4-column group is entirely inside or outside

  // W slice (zero-padded rows / columns) into registers: its latency overlaps the first pass's g / q loads below
  constexpr int WV = 2 * CPT;  // DP * 4 float4 / 256 threads
  float4 wreg[WV];
#pragma unroll
  for (int i = 0; i < WV; ++i) {
    const int idx = t + i * kPtThreads;
    const int c = idx >> 2, k = k0 + ((idx & 3) << 2);
    wreg[i] = (c < Dout && k < Kin) ? __ldg(reinterpret_cast<const float4*>(a.W + (int64_t)c * Kin + k)) : make_float4(0.f, 0.f, 0.f, 0.f);
  }

  const int c0 = t & 127, kg = t >> 7;   // dW: columns c0 (+128) of Dout, Kin columns kg*8 .. +7 of the slice
  const int kx = t & 15, bg = t >> 4;    // dx: Kin column kx of the slice, rows bg*4 .. +3 of the pass
  float accw[CPT][8];
#pragma unroll
  for (int j = 0; j < CPT; ++j)
#pragma unroll
    for (int i = 0; i < 8; ++i) accw[j][i] = 0.f;
  float accb = 0.f;  // CTA 0: db[t], rows added in order

  for (int b0 = 0; b0 < a.B; b0 += kPbRows) {
    // x slice of this pass: 64 rows x 4 float4 = one per thread (in flight under the dy arithmetic)
    float4 xreg;
    {
      const int r = t >> 2, k = k0 + ((t & 3) << 2);
      xreg = (b0 + r < a.B && k < Kin) ? __ldg(reinterpret_cast<const float4*>(a.x + (int64_t)(b0 + r) * Kin + k)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();  // the previous pass no longer reads dy_s / x_s
    // dy rows of this pass: one warp per row, eight rows per warp, the loads of four rows issued together;
    // arithmetic of l2norm_bwd_kernel (rowops.cu): same summation order, true division
    constexpr int RPB = 4, NV = 4 * CPT;
    for (int rb = 0; rb < kPbRows / 8; rb += RPB) {
      float gv[RPB][NV], qv[RPB][NV], nr[RPB];
#pragma unroll
      for (int u = 0; u < RPB; ++u) {
        const int row = b0 + warp + 8 * (rb + u);
        const bool ok = row < a.B;
        nr[u] = (ok && a.normalize) ? a.norm[row] : 1.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          const int c = lane + 32 * i;
          gv[u][i] = (ok && c < Dout) ? a.g[(int64_t)row * Dout + c] : 0.f;
          qv[u][i] = (ok && a.normalize && c < Dout) ? a.q[(int64_t)row * Dout + c] : 0.f;
        }
      }
#pragma unroll
      for (int u = 0; u < RPB; ++u) {
        float* out = dy_s + (warp + 8 * (rb + u)) * DYS;
        if (!a.normalize || b0 + warp + 8 * (rb + u) >= a.B) {
#pragma unroll
          for (int i = 0; i < NV; ++i) out[lane + 32 * i] = gv[u][i];  // plain Linear; rows past B hold zeros
        } else if (nr[u] > a.eps) {
          float dot = 0.f;
#pragma unroll
          for (int i = 0; i < NV; ++i) dot = fmaf(gv[u][i], qv[u][i], dot);
          dot = warp_sum(dot);
#pragma unroll
          for (int i = 0; i < NV; ++i) out[lane + 32 * i] = (lane + 32 * i < Dout) ? (gv[u][i] - dot * qv[u][i]) / nr[u] : 0.f;
        } else {
#pragma unroll
          for (int i = 0; i < NV; ++i) out[lane + 32 * i] = (lane + 32 * i < Dout) ? gv[u][i] / a.eps : 0.f;
        }
      }
    }
    if (b0 == 0) {
#pragma unroll
      for (int i = 0; i < WV; ++i) {
        const int idx = t + i * kPtThreads;
        *reinterpret_cast<float4*>(W_s + (idx >> 2) * kPbKS + ((idx & 3) << 2)) = wreg[i];
      }
    }
    *reinterpret_cast<float4*>(x_s + (t >> 2) * kPbKS + ((t & 3) << 2)) = xreg;
    __syncthreads();

    if (a.dW) {  // dW[c][k] += sum_b dy[b][c] x[b][k]
#pragma unroll 4
      for (int b = 0; b < kPbRows; ++b) {
        const float4 xa = *reinterpret_cast<const float4*>(x_s + b * kPbKS + kg * 8);
        const float4 xb = *reinterpret_cast<const float4*>(x_s + b * kPbKS + kg * 8 + 4);
#pragma unroll
        for (int j = 0; j < CPT; ++j) {
          const float d = dy_s[b * DYS + c0 + 128 * j];
          accw[j][0] = fmaf(d, xa.x, accw[j][0]);
          accw[j][1] = fmaf(d, xa.y, accw[j][1]);
          accw[j][2] = fmaf(d, xa.z, accw[j][2]);
          accw[j][3] = fmaf(d, xa.w, accw[j][3]);
          accw[j][4] = fmaf(d, xb.x, accw[j][4]);
          accw[j][5] = fmaf(d, xb.y, accw[j][5]);
          accw[j][6] = fmaf(d, xb.z, accw[j][6]);
          accw[j][7] = fmaf(d, xb.w, accw[j][7]);
        }
      }
    }
    if (a.dx) {  // dx[b][k] = sum_c dy[b][c] W[c][k]
      float accx[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 2
      for (int c = 0; c < DP; c += 4) {
        const float w0 = W_s[(c + 0) * kPbKS + kx], w1 = W_s[(c + 1) * kPbKS + kx];
        const float w2 = W_s[(c + 2) * kPbKS + kx], w3 = W_s[(c + 3) * kPbKS + kx];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const float4 d = *reinterpret_cast<const float4*>(dy_s + (bg * 4 + r) * DYS + c);
          accx[r] = fmaf(d.x, w0, accx[r]);
          accx[r] = fmaf(d.y, w1, accx[r]);
          accx[r] = fmaf(d.z, w2, accx[r]);
          accx[r] = fmaf(d.w, w3, accx[r]);
        }
      }
      if (k0 + kx < Kin) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const int row = b0 + bg * 4 + r;
          if (row < a.B) a.dx[(int64_t)row * Kin + k0 + kx] = accx[r];
        }
      }
    }
    if (a.db && blockIdx.x == 0 && t < Dout) {
      for (int b = 0; b < kPbRows; ++b) accb += dy_s[b * DYS + t];
    }
  }
  if (a.dW) {
#pragma unroll
    for (int j = 0; j < CPT; ++j) {
      const int c = c0 + 128 * j, k = k0 + kg * 8;
      if (c < Dout) {
        float* o = a.dW + (int64_t)c * Kin + k;
        if (k < Kin) *reinterpret_cast<float4*>(o) = make_float4(accw[j][0], accw[j][1], accw[j][2], accw[j][3]);
        if (k + 4 < Kin) *reinterpret_cast<float4*>(o + 4) = make_float4(accw[j][4], accw[j][5], accw[j][6], accw[j][7]);
      }
    }
  }
  if (a.db && blockIdx.x == 0 && t < Dout) a.db[t] = accb;
}

template <int CPT>
size_t fwd_smem() {
  const int DP = CPT * 128;
  return sizeof(float) * (size_t)(DP * kPtWs + kPtRows * kPtKT + kPtRows * DP + DP);
}
template <int CPT>
size_t bwd_smem() {
  const int DP = CPT * 128;
  return sizeof(float) * (size_t)(kPbRows * (DP + 4) + kPbRows * kPbKS + DP * kPbKS);
}

template <int CPT>
int launch_fwd(const ProjFwdArgs& a, cudaStream_t s) {
  static unsigned long long configured = 0;
  const size_t smem = fwd_smem<CPT>();
  if (first_use_on_device(configured))
    AVSSL_CUDA_OK(cudaFuncSetAttribute(linear_l2norm_fwd_kernel<CPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const dim3 grid(kPtCluster, (a.B + kPtRows - 1) / kPtRows);
  linear_l2norm_fwd_kernel<CPT><<<grid, kPtThreads, smem, s>>>(a);
  AVSSL_LAUNCH_OK("linear_l2norm_fwd_kernel");
  return AVSSL_OK;
}

template <int CPT>
int launch_bwd(const ProjBwdArgs& a, cudaStream_t s) {
  static unsigned long long configured = 0;
  const size_t smem = bwd_smem<CPT>();
  if (first_use_on_device(configured))
    AVSSL_CUDA_OK(cudaFuncSetAttribute(linear_l2norm_bwd_kernel<CPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  linear_l2norm_bwd_kernel<CPT><<<(a.Kin + kPbKS - 1) / kPbKS, kPtThreads, smem, s>>>(a);
  AVSSL_LAUNCH_OK("linear_l2norm_bwd_kernel");
  return AVSSL_OK;
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace
}  // namespace avssl

using namespace avssl;

extern "C" int avssl_linear_l2norm_max_dout() { return 256; }

extern "C" int avssl_linear_l2norm_fwd(const float* x, const float* W, const float* bias, int B, int Kin, int Dout, float eps,
                                       int normalize, float* q_out, float* norm_out, void* stream) {
  AVSSL_REQUIRE(x && W && q_out, AVSSL_ERR_INVALID_ARGUMENT, "linear_l2norm_fwd: null pointer");
  AVSSL_REQUIRE(B >= 0 && Kin > 0 && Dout > 0 && eps >= 0.f, AVSSL_ERR_INVALID_ARGUMENT, "linear_l2norm_fwd: bad sizes");
  AVSSL_REQUIRE(Dout <= avssl_linear_l2norm_max_dout() && Kin % 4 == 0 && aligned16(x) && aligned16(W), AVSSL_ERR_UNSUPPORTED,
                "linear_l2norm_fwd: needs Dout <= 256, Kin %% 4 == 0 and 16-byte aligned x / W (got Dout=%d Kin=%d)", Dout, Kin);
  AVSSL_REQUIRE((B + kPtRows - 1) / kPtRows <= 65535, AVSSL_ERR_INVALID_ARGUMENT, "linear_l2norm_fwd: too many rows");
  if (B == 0) return AVSSL_OK;
  ProjFwdArgs a{x, W, bias, q_out, norm_out, B, Kin, Dout, eps, normalize ? 1 : 0};
  return Dout <= 128 ? launch_fwd<1>(a, static_cast<cudaStream_t>(stream)) : launch_fwd<2>(a, static_cast<cudaStream_t>(stream));
}

extern "C" int avssl_linear_l2norm_bwd(const float* x, const float* W, const float* q, const float* norm, const float* grad_q,
                                       int B, int Kin, int Dout, float eps, int normalize, float* dx_out, float* dW_out,
                                       float* db_out, void* stream) {
  AVSSL_REQUIRE(x && W && grad_q && (!normalize || (q && norm)), AVSSL_ERR_INVALID_ARGUMENT, "linear_l2norm_bwd: null pointer");
  AVSSL_REQUIRE(B > 0 && Kin > 0 && Dout > 0 && eps >= 0.f, AVSSL_ERR_INVALID_ARGUMENT, "linear_l2norm_bwd: bad sizes");
  AVSSL_REQUIRE(Dout <= avssl_linear_l2norm_max_dout() && Kin % 4 == 0 && aligned16(x) && aligned16(W) &&
                    (!dx_out || aligned16(dx_out)) && (!dW_out || aligned16(dW_out)),
                AVSSL_ERR_UNSUPPORTED, "linear_l2norm_bwd: needs Dout <= 256, Kin %% 4 == 0 and 16-byte aligned tensors (got Dout=%d Kin=%d)",
                Dout, Kin);
  if (!dx_out && !dW_out && !db_out) return AVSSL_OK;
  ProjBwdArgs a{x, W, q, norm, grad_q, dx_out, dW_out, db_out, B, Kin, Dout, eps, normalize ? 1 : 0};
  return Dout <= 128 ? launch_bwd<1>(a, static_cast<cudaStream_t>(stream)) : launch_bwd<2>(a, static_cast<cudaStream_t>(stream));
}
