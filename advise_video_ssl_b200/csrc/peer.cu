// C3 — cross-GPU key exchange over NVLink peer memory (replaces the all_gather + row select of
// _batch_unshuffle, models/contrastive.py:216-230, and cat_all_gather of the keys).
//
// Host side: exchange buffers are cudaMalloc'ed here (a whole allocation, so that the CUDA IPC
// handle maps exactly this buffer), exported with cudaIpcGetMemHandle and opened by the other
// ranks of the box; the caller moves the 64-byte handles between processes (torch.distributed).
// Device side: peer.cuh.  Stand-alone push / wait+gather kernels live here; the push is also
// fused into the EMA launch (ema.cu) and the wait into the InfoNCE launch (infonce_tc.cu).
#include <string.h>

#include "peer.cuh"

namespace avssl {

__global__ void __launch_bounds__(256) peer_push_kernel(const avssl_peer_xchg x, const float* __restrict__ rows) {
  __shared__ unsigned long long s_epoch;
  peer_push_cta<false>(x, rows, blockIdx.x, &s_epoch);
}

// K2 + push in one launch: what travels is l2norm(feat) (Normalize of the key features,
// models/contrastive.py:350 + :216-230); CTA `rank` also keeps a local copy when y_local != null.
__global__ void __launch_bounds__(1024)
l2norm_push_kernel(const avssl_peer_xchg x, const float* __restrict__ feat, float eps, float* __restrict__ y_local) {
  __shared__ unsigned long long s_epoch;
  peer_push_cta<true>(x, feat, blockIdx.x, &s_epoch, eps, y_local);
}

// out[i] = gathered[row_idx ? row_idx[i] : rank * rows_per_rank + i]; bit-exact copy.
__global__ void __launch_bounds__(256)
peer_wait_gather_kernel(const avssl_peer_xchg x, const long long* __restrict__ row_idx, int n_out,
                        float* __restrict__ out, uint32_t* status) {
  __shared__ int s_slot;
  if (threadIdx.x < 32) {
    const int slot = peer_wait_all_warp(x, status);
    if (threadIdx.x == 0) s_slot = slot;
  }
  __syncthreads();
  const float* g = peer_payload(x.base[x.rank], s_slot, x);
  const int D4 = x.D / 4;
  const long long n_rows = (long long)x.world * x.rows_per_rank;
  for (int i = blockIdx.x; i < n_out; i += gridDim.x) {
    const long long r = row_idx ? row_idx[i] : (long long)x.rank * x.rows_per_rank + i;
    if (r < 0 || r >= n_rows) {
      if (threadIdx.x == 0 && status) atomicOr(status, AVSSL_DEVFLAG_BAD_INDEX);
      continue;
    }
    const float4* src = reinterpret_cast<const float4*>(g + (size_t)r * x.D);
    float4* dst = reinterpret_cast<float4*>(out + (size_t)i * x.D);
    for (int c = threadIdx.x; c < D4; c += blockDim.x) dst[c] = __ldcg(src + c);
  }
}

int peer_check(const avssl_peer_xchg* x, const char* who) {
  AVSSL_REQUIRE(x, AVSSL_ERR_INVALID_ARGUMENT, "%s: exchange descriptor is null", who);
  AVSSL_REQUIRE(x->world >= 1 && x->world <= AVSSL_MAX_PEERS && x->rank >= 0 && x->rank < x->world,
                AVSSL_ERR_INVALID_ARGUMENT, "%s: bad world/rank %d/%d (at most %d peers)", who, x->world, x->rank,
                AVSSL_MAX_PEERS);
  AVSSL_REQUIRE(x->rows_per_rank > 0 && x->D > 0 && x->D % 4 == 0, AVSSL_ERR_INVALID_ARGUMENT,
                "%s: rows_per_rank=%d, D=%d (D must be a positive multiple of 4)", who, x->rows_per_rank, x->D);
  for (int r = 0; r < x->world; ++r)
    AVSSL_REQUIRE(x->base[r] && (reinterpret_cast<uintptr_t>(x->base[r]) & 255u) == 0, AVSSL_ERR_INVALID_ARGUMENT,
                  "%s: base[%d] is null or not 256-byte aligned", who, r);
  return AVSSL_OK;
}

}  // namespace avssl

using namespace avssl;

extern "C" size_t avssl_peer_xchg_bytes(int world, int rows_per_rank, int D) {
  if (world < 1 || world > AVSSL_MAX_PEERS || rows_per_rank < 1 || D < 1) return 0;
  return sizeof(PeerHdr) + 2ull * world * rows_per_rank * D * sizeof(float);
}

extern "C" int avssl_peer_alloc(size_t bytes, void** dev_ptr_out, void* ipc_handle_out_host) {
  AVSSL_REQUIRE(dev_ptr_out && ipc_handle_out_host && bytes >= sizeof(PeerHdr), AVSSL_ERR_INVALID_ARGUMENT,
                "peer_alloc: null argument or fewer than %zu bytes", sizeof(PeerHdr));
  static_assert(sizeof(cudaIpcMemHandle_t) == AVSSL_IPC_HANDLE_BYTES, "IPC handle size");
  void* p = nullptr;
  AVSSL_CUDA_OK(cudaMalloc(&p, bytes));
  cudaError_t e = cudaMemset(p, 0, bytes);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(static_cast<cudaIpcMemHandle_t*>(ipc_handle_out_host), p);
  if (e != cudaSuccess) {
    cudaFree(p);
    set_error("peer_alloc: %s", cudaGetErrorString(e));
    return AVSSL_ERR_CUDA;
  }
  *dev_ptr_out = p;
  return AVSSL_OK;
}

extern "C" int avssl_peer_open(const void* ipc_handle_host, void** dev_ptr_out) {
  AVSSL_REQUIRE(ipc_handle_host && dev_ptr_out, AVSSL_ERR_INVALID_ARGUMENT, "peer_open: null argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, ipc_handle_host, sizeof(h));
  AVSSL_CUDA_OK(cudaIpcOpenMemHandle(dev_ptr_out, h, cudaIpcMemLazyEnablePeerAccess));
  return AVSSL_OK;
}

extern "C" int avssl_peer_close(void* dev_ptr) {
  AVSSL_REQUIRE(dev_ptr, AVSSL_ERR_INVALID_ARGUMENT, "peer_close: null pointer");
  AVSSL_CUDA_OK(cudaIpcCloseMemHandle(dev_ptr));
  return AVSSL_OK;
}

extern "C" int avssl_peer_free(void* dev_ptr) {
  AVSSL_REQUIRE(dev_ptr, AVSSL_ERR_INVALID_ARGUMENT, "peer_free: null pointer");
  AVSSL_CUDA_OK(cudaFree(dev_ptr));
  return AVSSL_OK;
}

extern "C" int avssl_peer_push_rows(const avssl_peer_xchg* x, const float* rows, void* stream) {
  int rc = peer_check(x, "peer_push_rows");
  if (rc != AVSSL_OK) return rc;
  AVSSL_REQUIRE(rows && (reinterpret_cast<uintptr_t>(rows) & 15u) == 0, AVSSL_ERR_INVALID_ARGUMENT,
                "peer_push_rows: rows is null or not 16-byte aligned");
  peer_push_kernel<<<x->world, 256, 0, static_cast<cudaStream_t>(stream)>>>(*x, rows);
  AVSSL_LAUNCH_OK("peer_push_kernel");
  return AVSSL_OK;
}

extern "C" int avssl_l2norm_push_rows(const avssl_peer_xchg* x, const float* feat, float eps, float* y_local_out,
                                      void* stream) {
  int rc = peer_check(x, "l2norm_push_rows");
  if (rc != AVSSL_OK) return rc;
  AVSSL_REQUIRE(feat && eps >= 0.f, AVSSL_ERR_INVALID_ARGUMENT, "l2norm_push_rows: feat is null or eps < 0");
  // one warp per row where possible: the launch is latency-bound (each warp: load row, reduce, divide, store)
  const int threads = x->rows_per_rank >= 32 ? 1024 : (x->rows_per_rank >= 16 ? 512 : 256);
  l2norm_push_kernel<<<x->world, threads, 0, static_cast<cudaStream_t>(stream)>>>(*x, feat, eps, y_local_out);
  AVSSL_LAUNCH_OK("l2norm_push_kernel");
  return AVSSL_OK;
}

extern "C" int avssl_peer_wait_gather(const avssl_peer_xchg* x, const int64_t* row_idx, int n_out, float* out,
                                      uint32_t* status_dev, void* stream) {
  int rc = peer_check(x, "peer_wait_gather");
  if (rc != AVSSL_OK) return rc;
  AVSSL_REQUIRE(out && n_out > 0 && (reinterpret_cast<uintptr_t>(out) & 15u) == 0, AVSSL_ERR_INVALID_ARGUMENT,
                "peer_wait_gather: out is null / misaligned or n_out=%d", n_out);
  AVSSL_REQUIRE(row_idx || n_out <= x->rows_per_rank, AVSSL_ERR_INVALID_ARGUMENT,
                "peer_wait_gather: n_out=%d exceeds rows_per_rank=%d without row_idx", n_out, x->rows_per_rank);
  const int grid = n_out < 64 ? n_out : 64;
  peer_wait_gather_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      *x, reinterpret_cast<const long long*>(row_idx), n_out, out, status_dev);
  AVSSL_LAUNCH_OK("peer_wait_gather_kernel");
  return AVSSL_OK;
}
