// K6 — SimCLR NT-Xent over THIS rank's rows against all gathered columns
// (replaces models/contrastive.py:770-792 and the gradient bookkeeping of
//  AllGatherWithGradient, utils/distributed.py:131-155).
//
//   out = [q_all ; q2_all]  (2N unit rows),  s_rc = out_r . out_c / T
//   Z_r  = sum_{c != r} e^{s_rc - 1/T}                       (pass 1, "rowsum")
//   loss = mean_r( log Z_r + 1/T - s_{r,r+} )                 (r+ = partner row)
//   G_r  = [ sum_{c != r} e^{s_rc-1/T} (1/Z_r + 1/Z_c) out_c - 2 out_{r+} ] / (2N T)   (pass 2)
//
// The reference materialises the 2N x 2N matrix (and four same-sized temporaries) on
// every rank; here each rank touches only its 2*B_local rows, never stores a tile of
// the matrix, and the 1/T shift (|s| <= 1/T for unit rows) keeps exp() in range
// where the reference's raw exp overflows for T < 0.0113.
// Both passes are split over column ranges; partials are reduced in a fixed order.
// CUDA-core kernels (fp32 exact) = AVSSL_IMPL_SIMT; the tcgen05 kernels are in ntxent_tc.cu.
#include <cuda_fp16.h>
#include <string.h>

#include "ntxent.cuh"
#include "peer.cuh"
#include "simt_tile.cuh"

namespace avssl {

constexpr float kLog2eN = 1.4426950408889634f;



template <int DP>
__device__ __forceinline__ void ntx_load_rows(float* qs, const NtxArgs& a, int i_base) {
  constexpr int KS = DP + 4;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int r = warp; r < kTileI; r += kSimtThreads / 32) {
    const int i = i_base + r;
    const float* src = (i < a.n_loc) ? a.out + (size_t)a.rows[i] * a.D : nullptr;
    for (int c = lane; c < DP; c += 32) qs[r * KS + c] = (src && c < a.D) ? src[c] : 0.f;
  }
}

template <int DP, bool kGrad>
__global__ void __launch_bounds__(kSimtThreads, 1) ntxent_pass_kernel(const NtxArgs a) {
  constexpr int KS = DP + 4;
  constexpr int CC = DP / 64;
  extern __shared__ __align__(16) float smem[];
  float* qs = smem;
  float* ks0 = qs + kTileI * KS;
  float* ps = ks0 + 2 * kTileJ * KS;  // only when kGrad
  __shared__ int s_rowid[kTileI];
  __shared__ float s_invz[kTileI];

  const int tid = threadIdx.x;
  const int split = blockIdx.x;
  const int i_base = blockIdx.y * kTileI;
  const int j_begin = split * a.cols_per_split;
  const int j_end = min(a.N2, j_begin + a.cols_per_split);
  const int n_tiles = (j_end - j_begin + kTileJ - 1) / kTileJ;

  if (n_tiles > 0) {
    load_tile<DP>(ks0, a.out, a.D, j_begin, j_end);
    cp_async_commit();
  }
  ntx_load_rows<DP>(qs, a, i_base);
  if (tid < kTileI) {
    const int i = i_base + tid;
    s_rowid[tid] = i < a.n_loc ? a.rows[i] : -1;
    if (kGrad) s_invz[tid] = i < a.n_loc ? 1.f / a.z_all[a.rows[i]] : 0.f;
  }

  const int ty = tid >> 4, tx = tid & 15;
  const float scale2 = a.inv_T * kLog2eN;
  float zsum[4] = {0.f, 0.f, 0.f, 0.f};
  float4 acc[4][CC];
#pragma unroll
  for (int ii = 0; ii < 4; ++ii)
#pragma unroll
    for (int cc = 0; cc < CC; ++cc) acc[ii][cc] = make_float4(0.f, 0.f, 0.f, 0.f);

  for (int t = 0; t < n_tiles; ++t) {
    float* ks = ks0 + (t & 1) * kTileJ * KS;
    if (t + 1 < n_tiles) {
      load_tile<DP>(ks0 + ((t + 1) & 1) * kTileJ * KS, a.out, a.D, j_begin + (t + 1) * kTileJ, j_end);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();

    float s[4][4];
    simt_s_tile<DP>(qs, ks, ty, tx, s);
    const int jt0 = j_begin + t * kTileJ;
#pragma unroll
    for (int ii = 0; ii < 4; ++ii) {
      const int rid = s_rowid[ty * 4 + ii];
      const float invz_r = kGrad ? s_invz[ty * 4 + ii] : 0.f;
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const int j = jt0 + tx + 16 * jj;
        const bool valid = (j < j_end) && (j != rid) && (rid >= 0);
        // e^{(s - 1)/T}: unit rows give s <= 1, so the exponent is <= 0
        const float e = valid ? exp2f((s[ii][jj] - 1.f) * scale2) : 0.f;
        if (kGrad) {
          const float w = valid ? e * (invz_r + 1.f / a.z_all[j]) : 0.f;
          ps[(ty * 4 + ii) * kPsStride + tx + 16 * jj] = w;
        } else {
          zsum[ii] += e;
        }
      }
    }
    if (kGrad) {
      __syncthreads();
      simt_pv_tile<DP>(ps, ks, ty, tx, acc);
    }
    __syncthreads();
  }

#pragma unroll
  for (int ii = 0; ii < 4; ++ii) {
    const int i = i_base + ty * 4 + ii;
    if (kGrad) {
      if (i < a.n_loc) {
#pragma unroll
        for (int cc = 0; cc < CC; ++cc) {
          const int c = tx * 4 + 64 * cc;
          if (c < a.D) *reinterpret_cast<float4*>(a.part_g + ((size_t)split * a.n_loc + i) * a.D + c) = acc[ii][cc];
        }
      }
    } else {
      float z = zsum[ii];
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) z += __shfl_xor_sync(0xffffffffu, z, o);
      if (tx == 0 && i < a.n_loc) a.part_z[(size_t)split * a.n_loc + i] = z;
    }
  }
}

__global__ void ntxent_sum_z_kernel(const float* __restrict__ part_z, int n_splits, int n_loc, float* __restrict__ z_loc) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_loc) return;
  // fixed summation order (deterministic); eight split rows in flight per thread
  float z = 0.f;
  for (int s0 = 0; s0 < n_splits; s0 += 8) {
    float v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = s0 + u < n_splits ? __ldcg(part_z + (size_t)(s0 + u) * n_loc + i) : 0.f;
#pragma unroll
    for (int u = 0; u < 8; ++u) z += v[u];
  }
  z_loc[i] = z;
}

// Finalisation of pass 2 in ONE launch (8 warps per CTA):
//   CTAs [0, comb_blocks):  one warp per local row -- merge the gradient partials in split order
//     (deterministic; 128-bit loads, 8 split rows in flight per lane), subtract the positive term,
//     apply 1/(2N T) and the world-size factor, then chain through the l2-normalisation.
//     D <= 256: lane l owns columns 4l..4l+3 and 128+4l..128+4l+3.
//   CTAs [comb_blocks, ...): about one per SM; every warp walks its share of the PAIRS (r, r+) of all 2N rows --
//     both rows share the same dot product: loss = mean_r( log Z_r + 1/T - out_r . out_{r+} / T ); the last of
//     these CTAs to finish folds the per-CTA sums in CTA order.
constexpr int kFinWarps = 8;

__global__ void __launch_bounds__(kFinWarps * 32)
ntxent_finish_kernel(const NtxArgs a, const float* __restrict__ norm_loc, float gscale, float* __restrict__ dfeat,
                     int comb_blocks, float* loss_out, float* row_term, unsigned* counter) {
  __shared__ float s_red[32];
  __shared__ unsigned s_last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int D = a.D, half = a.N2 / 2;
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);

  if ((int)blockIdx.x < comb_blocks) {
    const int i = blockIdx.x * kFinWarps + warp;
    if (i >= a.n_loc) return;
    const int r = a.rows[i];
    const int rp = r < half ? r + half : r - half;
    const int c0 = 4 * lane, c1 = 128 + 4 * lane;
    const bool in0 = c0 < D, in1 = c1 < D;
    const float4* o4 = reinterpret_cast<const float4*>(a.out + (size_t)r * D);
    const float4* p4 = reinterpret_cast<const float4*>(a.out + (size_t)rp * D);
    const float4 ov0 = in0 ? __ldg(o4 + lane) : zero4, ov1 = in1 ? __ldg(o4 + 32 + lane) : zero4;
    const float4 pv0 = in0 ? __ldg(p4 + lane) : zero4, pv1 = in1 ? __ldg(p4 + 32 + lane) : zero4;
    const float rn = 1.f / norm_loc[i];
    float4 g0 = zero4, g1 = zero4;
    const float4* pg = reinterpret_cast<const float4*>(a.part_g + (size_t)i * D);
    const size_t sstride4 = (size_t)a.n_loc * D / 4;
    constexpr int kS = 9;  // split rows in flight per lane: two rounds for the 18 splits of cfg3
    for (int s0 = 0; s0 < a.n_splits; s0 += kS) {  // fixed order: deterministic
      float4 v0[kS], v1[kS];
#pragma unroll
      for (int u = 0; u < kS; ++u) {
        const bool live = s0 + u < a.n_splits;
        v0[u] = (live && in0) ? __ldcg(pg + (size_t)(s0 + u) * sstride4 + lane) : zero4;
        v1[u] = (live && in1) ? __ldcg(pg + (size_t)(s0 + u) * sstride4 + 32 + lane) : zero4;
      }
#pragma unroll
      for (int u = 0; u < kS; ++u) {
        g0.x += v0[u].x; g0.y += v0[u].y; g0.z += v0[u].z; g0.w += v0[u].w;
        g1.x += v1[u].x; g1.y += v1[u].y; g1.z += v1[u].z; g1.w += v1[u].w;
      }
    }
    g0 = make_float4((g0.x - 2.f * pv0.x) * gscale, (g0.y - 2.f * pv0.y) * gscale, (g0.z - 2.f * pv0.z) * gscale,
                     (g0.w - 2.f * pv0.w) * gscale);
    g1 = make_float4((g1.x - 2.f * pv1.x) * gscale, (g1.y - 2.f * pv1.y) * gscale, (g1.z - 2.f * pv1.z) * gscale,
                     (g1.w - 2.f * pv1.w) * gscale);
    float dot = g0.x * ov0.x + g0.y * ov0.y + g0.z * ov0.z + g0.w * ov0.w + g1.x * ov1.x + g1.y * ov1.y + g1.z * ov1.z +
                g1.w * ov1.w;
    dot = warp_sum(dot);
    float4* d4 = reinterpret_cast<float4*>(dfeat + (size_t)i * D);
    if (in0)
      d4[lane] = make_float4((g0.x - dot * ov0.x) * rn, (g0.y - dot * ov0.y) * rn, (g0.z - dot * ov0.z) * rn,
                             (g0.w - dot * ov0.w) * rn);
    if (in1)
      d4[32 + lane] = make_float4((g1.x - dot * ov1.x) * rn, (g1.y - dot * ov1.y) * rn, (g1.z - dot * ov1.z) * rn,
                                  (g1.w - dot * ov1.w) * rn);
    return;
  }

  // ---- loss: mean_r( log Z_r + 1/T - out_r . out_{r+} / T ) over ALL 2N rows.  A few fat CTAs: every warp walks
  // its share of the pairs (both rows of a pair share the dot product) and keeps a running sum, the CTA folds its
  // warps' sums in a fixed order, and the last CTA to arrive folds the per-CTA sums -- deterministic, one atomic
  // per CTA and no per-row round trip through memory.
  const int n_loss_blocks = gridDim.x - comb_blocks;
  const int lb = (int)blockIdx.x - comb_blocks;
  const bool in0 = 4 * lane < D, in1 = 128 + 4 * lane < D;
  float wsum = 0.f;
  for (int r = lb * kFinWarps + warp; r < half; r += n_loss_blocks * kFinWarps) {  // pair index
    const float4* x4 = reinterpret_cast<const float4*>(a.out + (size_t)r * D);
    const float4* y4 = reinterpret_cast<const float4*>(a.out + (size_t)(r + half) * D);
    const float4 x0 = in0 ? __ldg(x4 + lane) : zero4, y0 = in0 ? __ldg(y4 + lane) : zero4;
    const float4 x1 = in1 ? __ldg(x4 + 32 + lane) : zero4, y1 = in1 ? __ldg(y4 + 32 + lane) : zero4;
    const float za = a.z_all[r], zb = a.z_all[r + half];
    float dot = (x0.x * y0.x + x0.y * y0.y) + (x0.z * y0.z + x0.w * y0.w) + (x1.x * y1.x + x1.y * y1.y) +
                (x1.z * y1.z + x1.w * y1.w);
    dot = warp_sum(dot);
    wsum += (logf(za) + a.inv_T - dot * a.inv_T) + (logf(zb) + a.inv_T - dot * a.inv_T);
  }
  if (lane == 0) s_red[warp] = wsum;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < kFinWarps; ++w) t += s_red[w];
    row_term[lb] = t;
    __threadfence();
    s_last = (atomicAdd(counter, 1u) == (unsigned)n_loss_blocks - 1u) ? 1u : 0u;
  }
  __syncthreads();
  if (s_last) {
    __threadfence();
    float t = 0.f;
    for (int k = threadIdx.x; k < n_loss_blocks; k += blockDim.x) t += __ldcg(row_term + k);
    const float tot = block_sum(t, s_red);
    if (threadIdx.x == 0) {
      *loss_out = tot / (float)a.N2;
      *counter = 0u;
    }
  }
}

template <int DP, bool kGrad>
static int launch_pass(const NtxArgs& a, cudaStream_t s) {
  const size_t smem = simt_smem_bytes(DP, kGrad);
  static unsigned long long configured = 0ull;  // device ordinals already set up
  if (first_use_on_device(configured)) {
    AVSSL_CUDA_OK(cudaFuncSetAttribute(ntxent_pass_kernel<DP, kGrad>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  dim3 grid(a.n_splits, (a.n_loc + kTileI - 1) / kTileI);
  ntxent_pass_kernel<DP, kGrad><<<grid, kSimtThreads, smem, s>>>(a);
  AVSSL_LAUNCH_OK("ntxent_pass_kernel");
  return AVSSL_OK;
}

template <bool kGrad>
static int launch_pass_d(const NtxArgs& a, cudaStream_t s) {
  if (a.D <= 64) return launch_pass<64, kGrad>(a, s);
  if (a.D <= 128) return launch_pass<128, kGrad>(a, s);
  return launch_pass<256, kGrad>(a, s);
}

static int plan_splits(int N2, int n_loc, int* n_splits, int* cols_per_split) {
  int sms = sm_count();
  if (sms <= 0) return -1;
  const int row_blocks = (n_loc + kTileI - 1) / kTileI;
  const int n_tiles = (N2 + kTileJ - 1) / kTileJ;
  int S = (2 * sms + row_blocks - 1) / row_blocks;
  if (S < 1) S = 1;
  if (S > n_tiles) S = n_tiles;
  if (S > 64) S = 64;
  const int tps = (n_tiles + S - 1) / S;
  *n_splits = (n_tiles + tps - 1) / tps;
  *cols_per_split = tps * kTileJ;
  return 0;
}

}  // namespace avssl

using namespace avssl;

extern "C" size_t avssl_ntxent_workspace_bytes(int N2, int D, int n_loc) {
  if (N2 <= 0 || D <= 0 || n_loc <= 0) return 0;
  // counter | row_term[N2] | part_z[64][n_loc] | part_g[64][n_loc][D]
  return 256 + 4 * ((size_t)N2 + 64) + 4 * 64 * ((size_t)n_loc + 64) + 4 * 64 * (size_t)n_loc * D + 1024;
}

static int ntx_setup(NtxArgs& a, const float* out, const void* out_f16, const int* rows, int row0_first, int row1_first, const float* z_all, int N2, int D, int n_loc,
                     float T, void* workspace, size_t workspace_bytes, int impl, bool* use_tc, const char* who) {
  AVSSL_REQUIRE(out && rows && workspace, AVSSL_ERR_INVALID_ARGUMENT, "%s: null pointer", who);
  AVSSL_REQUIRE(N2 > 0 && (N2 % 2) == 0 && n_loc > 0 && n_loc <= N2 && T > 0.f, AVSSL_ERR_INVALID_ARGUMENT,
                "%s: bad sizes N2=%d n_loc=%d", who, N2, n_loc);
  AVSSL_REQUIRE(D % 4 == 0 && D >= 4 && D <= 256, AVSSL_ERR_UNSUPPORTED, "%s: D=%d unsupported (D %% 4 == 0, D <= 256)", who, D);
  AVSSL_REQUIRE(workspace_bytes >= avssl_ntxent_workspace_bytes(N2, D, n_loc), AVSSL_ERR_WORKSPACE, "%s: workspace too small", who);
  a.out = out;
  a.out_f16 = static_cast<const uint16_t*>(out_f16);
  a.rows = rows;
  a.q_row0 = row0_first;
  a.q_row1 = row1_first;
  a.z_all = z_all;
  a.N2 = N2;
  a.D = D;
  a.n_loc = n_loc;
  a.inv_T = 1.f / T;
  AVSSL_REQUIRE(impl == AVSSL_IMPL_AUTO || impl == AVSSL_IMPL_SIMT || impl == AVSSL_IMPL_TC1X, AVSSL_ERR_INVALID_ARGUMENT,
                "%s: impl must be AVSSL_IMPL_AUTO, AVSSL_IMPL_SIMT or AVSSL_IMPL_TC1X (got %d)", who, impl);
  *use_tc = impl != AVSSL_IMPL_SIMT && ntxent_tc_supported(N2, D, n_loc) && out_f16 != nullptr && row0_first >= 0 && row1_first >= 0 &&
            n_loc % 2 == 0 && row0_first + n_loc / 2 <= N2 && row1_first + n_loc / 2 <= N2 &&
            (reinterpret_cast<uintptr_t>(out_f16) & 15u) == 0;
  AVSSL_REQUIRE(*use_tc || impl != AVSSL_IMPL_TC1X, AVSSL_ERR_UNSUPPORTED,
                "%s: the tcgen05 kernel needs D in {64,128,256} and a 16-byte aligned fp16 copy of `out` "
                "(avssl_ntxent_prepare) (D=%d)", who, D);
  AVSSL_REQUIRE((*use_tc ? ntxent_tc_plan(N2, n_loc, z_all != nullptr, &a.n_splits, &a.cols_per_split)
                         : plan_splits(N2, n_loc, &a.n_splits, &a.cols_per_split)) == 0,
                AVSSL_ERR_CUDA, "%s: no CUDA device", who);
  char* w = static_cast<char*>(workspace);
  size_t off = 256 + 4 * ((size_t)N2 + 64);
  off = (off + 255) / 256 * 256;
  a.part_z = reinterpret_cast<float*>(w + off);
  off += 4 * 64 * ((size_t)n_loc + 64);
  off = (off + 255) / 256 * 256;
  a.part_g = reinterpret_cast<float*>(w + off);
  return AVSSL_OK;
}

// out[(v * W + w) * B + b] = gathered[w][v][b] ([q_all ; q2_all], models/contrastive.py:771-775) and its
// fp16 copy (round to nearest), one pass: 128-bit loads, a 128-bit and a 64-bit store per element group.
// kPeer: the gathered rows sit in this rank's NVLink exchange buffer (every rank pushed its [2B, D] block, C4 without
// a collective kernel): wait for all flags of the current epoch first, then read the slot.
template <bool kPeer>
__global__ void __launch_bounds__(256)
ntxent_prepare_kernel(const float4* __restrict__ gathered, const avssl_peer_xchg x, uint32_t* status, int W, int B, int D4,
                      float4* __restrict__ out, uint2* __restrict__ out_f16) {
  if (kPeer) {
    __shared__ int s_slot;
    if (threadIdx.x < 32) {
      const int slot = peer_wait_all_warp(x, status);
      if (threadIdx.x == 0) s_slot = slot;
    }
    __syncthreads();
    gathered = reinterpret_cast<const float4*>(peer_payload(x.base[x.rank], s_slot, x));
  }
  // one warp per destination row (v * W + w) * B + b; the row index arithmetic is per warp, not per element, and
  // all of a row's loads are issued before its first store
  const int lane = threadIdx.x & 31;
  const int n_rows = 2 * W * B;
  constexpr int kMaxPerLane = 4;  // D <= 512 in one go (D4 <= 128)
  for (int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < n_rows; row += gridDim.x * (blockDim.x >> 5)) {
    const int b = row % B, vw = row / B;
    const int w = vw % W, v = vw / W;
    const float4* src = gathered + ((size_t)(w * 2 + v) * B + b) * D4;
    float4* o = out + (size_t)row * D4;
    uint2* oh = out_f16 + (size_t)row * D4;
    for (int c0 = 0; c0 < D4; c0 += 32 * kMaxPerLane) {
      float4 xv[kMaxPerLane];
#pragma unroll
      for (int u = 0; u < kMaxPerLane; ++u) {
        const int c = c0 + lane + 32 * u;
        if (c < D4) xv[u] = kPeer ? __ldcg(src + c) : ldg_stream(src + c);  // peer rows were written by other GPUs: through L2
      }
#pragma unroll
      for (int u = 0; u < kMaxPerLane; ++u) {
        const int c = c0 + lane + 32 * u;
        if (c < D4) {
          o[c] = xv[u];
          const __half2 lo = __floats2half2_rn(xv[u].x, xv[u].y), hi = __floats2half2_rn(xv[u].z, xv[u].w);
          oh[c] = make_uint2(*reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi));
        }
      }
    }
  }
}

static int ntx_prepare(const float* gathered, const avssl_peer_xchg* x, uint32_t* status, int world, int B, int D, float* out,
                       void* out_f16, void* stream);

extern "C" int avssl_ntxent_prepare_peer(const avssl_peer_xchg* x, uint32_t* status_dev, int B, int D, float* out,
                                         void* out_f16, void* stream) {
  int rc = peer_check(x, "ntxent_prepare_peer");
  if (rc != AVSSL_OK) return rc;
  AVSSL_REQUIRE(x->rows_per_rank == 2 * B && x->D == D, AVSSL_ERR_INVALID_ARGUMENT,
                "ntxent_prepare_peer: the exchange carries [%d x %d] per rank, expected [%d x %d]", x->rows_per_rank, x->D, 2 * B, D);
  return ntx_prepare(nullptr, x, status_dev, x->world, B, D, out, out_f16, stream);
}

extern "C" int avssl_ntxent_prepare(const float* gathered, int world, int B, int D, float* out, void* out_f16,
                                    void* stream) {
  AVSSL_REQUIRE(gathered, AVSSL_ERR_INVALID_ARGUMENT, "ntxent_prepare: gathered is null");
  return ntx_prepare(gathered, nullptr, nullptr, world, B, D, out, out_f16, stream);
}

static int ntx_prepare(const float* gathered, const avssl_peer_xchg* x, uint32_t* status, int world, int B, int D, float* out,
                       void* out_f16, void* stream) {
  AVSSL_REQUIRE((gathered || x) && out && out_f16 && world > 0 && B > 0 && D > 0 && D % 4 == 0, AVSSL_ERR_INVALID_ARGUMENT,
                "ntxent_prepare: bad arguments (world=%d B=%d D=%d, D %% 4 == 0)", world, B, D);
  AVSSL_REQUIRE(((reinterpret_cast<uintptr_t>(gathered) | reinterpret_cast<uintptr_t>(out) |
                  reinterpret_cast<uintptr_t>(out_f16)) & 15u) == 0,
                AVSSL_ERR_INVALID_ARGUMENT, "ntxent_prepare: pointers must be 16-byte aligned");
  int grid = (2 * world * B + 7) / 8;  // 8 warps = 8 rows per CTA
  const int cap = 16 * (sm_count() > 0 ? sm_count() : 148);
  if (grid > cap) grid = cap;
  avssl_peer_xchg none;
  memset(&none, 0, sizeof(none));
  if (x)
    ntxent_prepare_kernel<true><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        nullptr, *x, status, world, B, D / 4, reinterpret_cast<float4*>(out), reinterpret_cast<uint2*>(out_f16));
  else
    ntxent_prepare_kernel<false><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const float4*>(gathered), none, nullptr, world, B, D / 4, reinterpret_cast<float4*>(out),
        reinterpret_cast<uint2*>(out_f16));
  AVSSL_LAUNCH_OK("ntxent_prepare_kernel");
  return AVSSL_OK;
}

extern "C" int avssl_ntxent_rowsum(const float* out, const void* out_f16, const int* rows, int row0_first,
                                   int row1_first, int N2, int D, int n_loc, float T,
                                   float* z_loc_out, void* workspace, size_t workspace_bytes, int impl, void* stream) {
  NtxArgs a;
  bool use_tc = false;
  int rc = ntx_setup(a, out, out_f16, rows, row0_first, row1_first, nullptr, N2, D, n_loc, T, workspace, workspace_bytes, impl, &use_tc, "ntxent_rowsum");
  if (rc) return rc;
  AVSSL_REQUIRE(z_loc_out, AVSSL_ERR_INVALID_ARGUMENT, "ntxent_rowsum: null output");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  rc = use_tc ? launch_ntxent_tc(a, false, s) : launch_pass_d<false>(a, s);
  if (rc) return rc;
  ntxent_sum_z_kernel<<<(n_loc + 63) / 64, 64, 0, s>>>(a.part_z, a.n_splits, n_loc, z_loc_out);
  AVSSL_LAUNCH_OK("ntxent_sum_z_kernel");
  return AVSSL_OK;
}

extern "C" int avssl_ntxent_grad(const float* out, const void* out_f16, const int* rows, int row0_first,
                                 int row1_first, const float* z_all, const float* norm_loc, int N2,
                                 int D, int n_loc, float T, float grad_scale, float* loss_out, float* dfeat_out,
                                 void* workspace, size_t workspace_bytes, int impl, void* stream) {
  NtxArgs a;
  bool use_tc = false;
  int rc = ntx_setup(a, out, out_f16, rows, row0_first, row1_first, z_all, N2, D, n_loc, T, workspace, workspace_bytes, impl, &use_tc, "ntxent_grad");
  if (rc) return rc;
  AVSSL_REQUIRE(z_all && norm_loc && loss_out && dfeat_out, AVSSL_ERR_INVALID_ARGUMENT, "ntxent_grad: null pointer");
  AVSSL_REQUIRE(((reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(dfeat_out)) & 15u) == 0,
                AVSSL_ERR_INVALID_ARGUMENT, "ntxent_grad: out and dfeat_out must be 16-byte aligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  rc = use_tc ? launch_ntxent_tc(a, true, s) : launch_pass_d<true>(a, s);
  if (rc) return rc;
  const float gscale = grad_scale * a.inv_T / (float)N2;
  unsigned* counter = static_cast<unsigned*>(workspace);
  float* row_term = reinterpret_cast<float*>(static_cast<char*>(workspace) + 256);
  const int comb_blocks = (n_loc + kFinWarps - 1) / kFinWarps;
  int loss_blocks = (N2 / 2 + kFinWarps - 1) / kFinWarps;  // a few pairs per warp: about two CTAs per SM
  if (loss_blocks > 2 * (sm_count() > 0 ? sm_count() : 148)) loss_blocks = 2 * (sm_count() > 0 ? sm_count() : 148);
  ntxent_finish_kernel<<<comb_blocks + loss_blocks, kFinWarps * 32, 0, s>>>(a, norm_loc, gscale, dfeat_out, comb_blocks,
                                                                           loss_out, row_term, counter);
  AVSSL_LAUNCH_OK("ntxent_finish_kernel");
  return AVSSL_OK;
}
