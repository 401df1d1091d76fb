// K2 / K7 / K9 — row-wise kernels: l2-normalise forward/backward, BYOL similarity
// loss forward+backward.  One warp per row, rows are [n, D] fp32 contiguous.
//   Normalize            models/contrastive.py:923-934   (eps = 0)
//   F.normalize          :850,867 (SwAV), :617-621 (prototype renorm)  (eps = 1e-12)
//   sim_loss + l2_norm   :243-249, :533, :572-582 (BYOL)
#include "common.cuh"

namespace avssl {

__global__ void __launch_bounds__(256)
l2norm_fwd_kernel(const float* __restrict__ x, int n, int D, float eps, float* __restrict__ y, float* __restrict__ norm_out) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= n) return;
  const float* xr = x + (int64_t)r * D;
  const float nrm = sqrtf(row_sumsq(xr, D, lane));
  const float den = fmaxf(nrm, eps);  // eps == 0 reproduces Normalize exactly
  for (int c = lane; c < D; c += 32) y[(int64_t)r * D + c] = xr[c] / den;
  if (norm_out && lane == 0) norm_out[r] = nrm;
}

// dx = (dy - (dy.y) y) / ||x||   when ||x|| > eps, else dy / eps
__global__ void __launch_bounds__(256)
l2norm_bwd_kernel(const float* __restrict__ y, const float* __restrict__ norm, const float* __restrict__ dy, int n,
                  int D, float eps, float* __restrict__ dx) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= n) return;
  const float* yr = y + (int64_t)r * D;
  const float* gr = dy + (int64_t)r * D;
  const float nrm = norm[r];
  if (nrm > eps) {
    float dot = 0.f;
    for (int c = lane; c < D; c += 32) dot = fmaf(gr[c], yr[c], dot);
    dot = warp_sum(dot);
    for (int c = lane; c < D; c += 32) dx[(int64_t)r * D + c] = (gr[c] - dot * yr[c]) / nrm;
  } else {
    for (int c = lane; c < D; c += 32) dx[(int64_t)r * D + c] = gr[c] / eps;
  }
}

// BYOL: loss = -mean_n( sum_c p_nc k_nc ) / T with p = pred/||pred|| (normalize=1)
// or p = pred (normalize=0).  Also emits d loss / d pred.  Deterministic mean: row
// sums go to `row_sim`, the last CTA reduces them in row order.
__global__ void __launch_bounds__(256)
byol_simloss_kernel(const float* __restrict__ pred, const float* __restrict__ key, int n, int D, float inv_T,
                    int normalize, float* __restrict__ loss_out, float* __restrict__ dpred, float* row_sim,
                    unsigned* counter) {
  __shared__ float s_red[32];
  __shared__ unsigned s_last;
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r < n) {
    const float* pr = pred + (int64_t)r * D;
    const float* kr = key + (int64_t)r * D;
    const float nrm = normalize ? sqrtf(row_sumsq(pr, D, lane)) : 1.f;
    float dot = 0.f;
    for (int c = lane; c < D; c += 32) dot = fmaf(pr[c] / nrm, kr[c], dot);
    dot = warp_sum(dot);  // = p . k
    if (lane == 0) row_sim[r] = dot * inv_T;
    if (dpred) {
      // dL/dp = -k / (T n);  through the normalisation: (dp - (dp.p) p) / ||pred||
      const float g = -inv_T / (float)n;
      for (int c = lane; c < D; c += 32) {
        const float p = pr[c] / nrm;
        const float dp = g * kr[c];
        dpred[(int64_t)r * D + c] = normalize ? (dp - (g * dot) * p) / nrm : dp;
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    s_last = (atomicAdd(counter, 1u) == gridDim.x - 1) ? 1u : 0u;
  }
  __syncthreads();
  if (s_last) {
    __threadfence();
    float tot = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) tot += reinterpret_cast<volatile float*>(row_sim)[i];
    tot = block_sum(tot, s_red);
    if (threadIdx.x == 0) {
      *loss_out = -(tot / (float)n);
      *counter = 0u;
    }
  }
}

// ContrastiveLoss (models/losses.py:15-25): CrossEntropy(logits, target = 0), mean.
// One CTA per row; deterministic mean by the last CTA.
__global__ void __launch_bounds__(256)
ce_target0_fwd_kernel(const float* __restrict__ logits, int n, int C, float* __restrict__ loss_out,
                      float* __restrict__ row_lse, float* row_loss, unsigned* counter) {
  __shared__ float s_red[32];
  __shared__ unsigned s_last;
  const int r = blockIdx.x;
  const float* row = logits + (int64_t)r * C;
  float mx = -INFINITY;
  for (int c = threadIdx.x; c < C; c += blockDim.x) mx = fmaxf(mx, row[c]);
  mx = warp_max(mx);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = mx;
  __syncthreads();
  mx = s_red[0];
  for (int w = 1; w < (int)(blockDim.x >> 5); ++w) mx = fmaxf(mx, s_red[w]);
  float se = 0.f;
  for (int c = threadIdx.x; c < C; c += blockDim.x) se += expf(row[c] - mx);
  se = block_sum(se, s_red);
  const float lse = mx + logf(se);
  if (threadIdx.x == 0) {
    row_lse[r] = lse;
    row_loss[r] = lse - row[0];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    s_last = (atomicAdd(counter, 1u) == gridDim.x - 1) ? 1u : 0u;
  }
  __syncthreads();
  if (s_last) {
    __threadfence();
    float tot = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) tot += reinterpret_cast<volatile float*>(row_loss)[i];
    tot = block_sum(tot, s_red);
    if (threadIdx.x == 0) {
      *loss_out = tot / (float)n;
      *counter = 0u;
    }
  }
}

// dlogits[r][c] = (softmax[r][c] - [c == 0]) * gscale,  gscale = grad_out / n
__global__ void __launch_bounds__(256)
ce_target0_bwd_kernel(const float* __restrict__ logits, const float* __restrict__ row_lse, int n, int C,
                      const float* __restrict__ grad_out, float* __restrict__ dlogits) {
  const int r = blockIdx.y;
  const float g = grad_out[0] / (float)n;
  const float lse = row_lse[r];
  const float* row = logits + (int64_t)r * C;
  float* out = dlogits + (int64_t)r * C;
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < C; c += gridDim.x * blockDim.x)
    out[c] = (expf(row[c] - lse) - (c == 0 ? 1.f : 0.f)) * g;
}

}  // namespace avssl

using namespace avssl;

extern "C" int avssl_l2norm_fwd(const float* x, int n, int D, float eps, float* y, float* norm_out, void* stream) {
  AVSSL_REQUIRE(x && y, AVSSL_ERR_INVALID_ARGUMENT, "l2norm_fwd: null pointer");
  AVSSL_REQUIRE(n >= 0 && D > 0 && eps >= 0.f, AVSSL_ERR_INVALID_ARGUMENT, "l2norm_fwd: bad sizes");
  if (n == 0) return AVSSL_OK;
  l2norm_fwd_kernel<<<(n + 7) / 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, n, D, eps, y, norm_out);
  AVSSL_LAUNCH_OK("l2norm_fwd_kernel");
  return AVSSL_OK;
}

extern "C" int avssl_l2norm_bwd(const float* y, const float* norm, const float* dy, int n, int D, float eps,
                                float* dx, void* stream) {
  AVSSL_REQUIRE(y && norm && dy && dx, AVSSL_ERR_INVALID_ARGUMENT, "l2norm_bwd: null pointer");
  AVSSL_REQUIRE(n >= 0 && D > 0 && eps >= 0.f, AVSSL_ERR_INVALID_ARGUMENT, "l2norm_bwd: bad sizes");
  if (n == 0) return AVSSL_OK;
  l2norm_bwd_kernel<<<(n + 7) / 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(y, norm, dy, n, D, eps, dx);
  AVSSL_LAUNCH_OK("l2norm_bwd_kernel");
  return AVSSL_OK;
}

extern "C" size_t avssl_byol_simloss_workspace_bytes(int n) { return 256 + sizeof(float) * (size_t)(n > 0 ? n : 0); }

extern "C" int avssl_byol_simloss_fwd_bwd(const float* pred, const float* key, int n, int D, float T, int normalize,
                                          float* loss_out, float* dpred_out, void* workspace, size_t workspace_bytes,
                                          void* stream) {
  AVSSL_REQUIRE(pred && key && loss_out && workspace, AVSSL_ERR_INVALID_ARGUMENT, "byol_simloss: null pointer");
  AVSSL_REQUIRE(n > 0 && D > 0 && T > 0.f, AVSSL_ERR_INVALID_ARGUMENT, "byol_simloss: bad sizes");
  AVSSL_REQUIRE(workspace_bytes >= avssl_byol_simloss_workspace_bytes(n), AVSSL_ERR_WORKSPACE,
                "byol_simloss: workspace too small");
  unsigned* counter = static_cast<unsigned*>(workspace);
  float* row_sim = reinterpret_cast<float*>(static_cast<char*>(workspace) + 256);
  byol_simloss_kernel<<<(n + 7) / 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      pred, key, n, D, 1.0f / T, normalize ? 1 : 0, loss_out, dpred_out, row_sim, counter);
  AVSSL_LAUNCH_OK("byol_simloss_kernel");
  return AVSSL_OK;
}

extern "C" size_t avssl_ce_target0_workspace_bytes(int n) { return 256 + sizeof(float) * (size_t)(n > 0 ? n : 0); }

extern "C" int avssl_ce_target0_fwd(const float* logits, int n, int C, float* loss_out, float* row_lse_out,
                                    void* workspace, size_t workspace_bytes, void* stream) {
  AVSSL_REQUIRE(logits && loss_out && row_lse_out && workspace, AVSSL_ERR_INVALID_ARGUMENT, "ce_target0_fwd: null pointer");
  AVSSL_REQUIRE(n > 0 && C > 0, AVSSL_ERR_INVALID_ARGUMENT, "ce_target0_fwd: bad sizes");
  AVSSL_REQUIRE(workspace_bytes >= avssl_ce_target0_workspace_bytes(n), AVSSL_ERR_WORKSPACE, "ce_target0_fwd: workspace too small");
  unsigned* counter = static_cast<unsigned*>(workspace);
  float* row_loss = reinterpret_cast<float*>(static_cast<char*>(workspace) + 256);
  ce_target0_fwd_kernel<<<n, 256, 0, static_cast<cudaStream_t>(stream)>>>(logits, n, C, loss_out, row_lse_out, row_loss, counter);
  AVSSL_LAUNCH_OK("ce_target0_fwd_kernel");
  return AVSSL_OK;
}

extern "C" int avssl_ce_target0_bwd(const float* logits, const float* row_lse, int n, int C, const float* grad_out_dev,
                                    float* dlogits, void* stream) {
  AVSSL_REQUIRE(logits && row_lse && grad_out_dev && dlogits, AVSSL_ERR_INVALID_ARGUMENT, "ce_target0_bwd: null pointer");
  AVSSL_REQUIRE(n > 0 && n <= 65535 && C > 0, AVSSL_ERR_INVALID_ARGUMENT, "ce_target0_bwd: bad sizes");
  int gx = (C + 1023) / 1024;
  if (gx > 64) gx = 64;
  ce_target0_bwd_kernel<<<dim3(gx, n), 256, 0, static_cast<cudaStream_t>(stream)>>>(logits, row_lse, n, C, grad_out_dev, dlogits);
  AVSSL_LAUNCH_OK("ce_target0_bwd_kernel");
  return AVSSL_OK;
}
