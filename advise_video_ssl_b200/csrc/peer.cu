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

// push for LARGE blocks (the SimCLR row gather, C4: 1 MiB per rank and destination at cfg3): M CTAs per destination
// split the block's bytes; the last of them (counter per destination in the header) publishes the flag, the last CTA
// of the whole grid advances the local epoch.
__global__ void __launch_bounds__(512)
peer_push_wide_kernel(const avssl_peer_xchg x, const float4* __restrict__ rows, int M) {
  __shared__ unsigned long long s_epoch;
  PeerHdr* me = static_cast<PeerHdr*>(x.base[x.rank]);
  if (threadIdx.x == 0) s_epoch = *reinterpret_cast<volatile unsigned long long*>(&me->epoch) + 1ull;
  __syncthreads();
  const unsigned long long e = s_epoch;
  const int d = blockIdx.x / M, m = blockIdx.x % M;
  const size_t n4 = (size_t)x.rows_per_rank * x.D / 4;
  const size_t lo = n4 * m / M, hi = n4 * (m + 1) / M;
  float4* out = reinterpret_cast<float4*>(peer_payload(x.base[d], (int)(e & 1ull), x) + (size_t)x.rank * x.rows_per_rank * x.D);
#pragma unroll 8
  for (size_t i = lo + threadIdx.x; i < hi; i += blockDim.x) out[i] = __ldg(rows + i);
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    if (atomicAdd(&me->pad_[d], 1u) == (unsigned)M - 1u) {  // last CTA of destination d
      me->pad_[d] = 0u;
      __threadfence_system();
      st_release_sys_u64(&static_cast<PeerHdr*>(x.base[d])->flags[x.rank], e);
    }
    if (atomicAdd(&me->done, 1u) == gridDim.x - 1u) {
      me->done = 0u;
      *reinterpret_cast<volatile unsigned long long*>(&me->epoch) = e;
      __threadfence();
    }
  }
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

// ---------------------------------------------------------------------------------------------------
// C1 as a scatter over NVLink peer stores (the shuffle of models/contrastive.py:174-214), ONE launch.
//
// Geometry: the same header as the key exchange, then payload[2][rows_per_rank * D] (D = floats per row,
// whatever the tensor's dtype): a rank receives exactly rows_per_rank rows per step, each written by the rank
// that owns it STRAIGHT INTO ITS FINAL POSITION -- local row j of this rank has global id g = rank*rows + j and
// goes to position dest_pos[j] = argsort(perm)[g] of the rank-major shuffled batch, i.e. to rank
// dest_pos[j] / rows, row dest_pos[j] % rows.  One read of the source row, one NVLink write, no send-buffer
// gather, no reordering on arrival (the reference all_gathers world x the bytes and discards (world-1)/world).
//
// Grid = world * M CTAs.  CTA (d, m) stores the m-th of M byte ranges of every local row destined to rank d,
// then the last of the M CTAs of destination d (counter in the header) publishes flags[rank] = epoch at rank d.
// Every CTA then waits for all `world` local flags (no rank's scatter depends on another's, so this cannot
// deadlock) and copies its share of the received rows out of the slot to `out` -- the slot alternates with the
// epoch, and a CUDA graph needs a fixed destination.  With small rows M = 1 and the dependent chain is
// loads -> NVLink stores -> fence -> flag | poll -> copy: it has to fit under the EMA kernel it runs beside.
__device__ __forceinline__ uint4* scatter_slot(void* base, int slot, const avssl_peer_xchg& x) {
  return reinterpret_cast<uint4*>(static_cast<char*>(base) + sizeof(PeerHdr)) + (size_t)slot * x.rows_per_rank * (x.D / 4);
}

// Destination positions passed BY VALUE in the kernel parameters (batches of up to kInlinePos rows): the launch then
// depends on no host-to-device copy -- beside a kernel that saturates HBM, that copy and the scheduling gap behind it
// cost ~15 us of the window the exchange has to fit in.
constexpr int kInlinePos = 256;
struct InlinePos {
  uint32_t pos[kInlinePos];
};

template <bool kInline>
__global__ void __launch_bounds__(512)
peer_scatter_exchange_kernel(const avssl_peer_xchg x, const uint4* __restrict__ rows, const long long* __restrict__ dest_pos,
                             const __grid_constant__ InlinePos inl, uint4* __restrict__ out, int M, uint32_t* status) {
  __shared__ unsigned long long s_epoch;
  PeerHdr* me = static_cast<PeerHdr*>(x.base[x.rank]);
  if (threadIdx.x == 0) s_epoch = *reinterpret_cast<volatile unsigned long long*>(&me->epoch) + 1ull;
  __syncthreads();
  const unsigned long long e = s_epoch;
  const int slot = (int)(e & 1ull);
  const int d = blockIdx.x / M, m = blockIdx.x % M;
  const size_t row16 = (size_t)x.D / 4;  // 16-byte units per row
  const size_t c_lo = row16 * m / M, c_hi = row16 * (m + 1) / M;
  const long long n_pos = (long long)x.world * x.rows_per_rank;
  uint4* dst_slot = scatter_slot(x.base[d], slot, x);
  // rows in batches of blockDim: one coalesced load of the destinations, then one WARP per row (all rows of
  // a batch are in flight at once -- the launch runs beside a kernel that saturates HBM, so every dependent
  // memory round trip costs microseconds)
  __shared__ long long s_pos[512];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
  for (int j0 = 0; j0 < x.rows_per_rank; j0 += (int)blockDim.x) {
    const int nj = min((int)blockDim.x, x.rows_per_rank - j0);
    __syncthreads();
    if ((int)threadIdx.x < nj)
      s_pos[threadIdx.x] = kInline ? (long long)inl.pos[j0 + threadIdx.x] : __ldg(dest_pos + j0 + threadIdx.x);
    __syncthreads();
    for (int jj = warp; jj < nj; jj += n_warps) {
      const long long pos = s_pos[jj];
      if (pos < 0 || pos >= n_pos) {
        if (lane == 0 && d == 0 && m == 0 && status) atomicOr(status, AVSSL_DEVFLAG_BAD_INDEX);
        continue;
      }
      if ((int)(pos / x.rows_per_rank) != d) continue;
      const uint4* src = rows + (size_t)(j0 + jj) * row16;
      uint4* o = dst_slot + (size_t)(pos % x.rows_per_rank) * row16;
#pragma unroll 4
      for (size_t c = c_lo + lane; c < c_hi; c += 32) o[c] = __ldg(src + c);
    }
  }
  __syncthreads();  // the CTA's stores are ordered before thread 0's fence (barrier + cumulativity)
  if (threadIdx.x == 0) {
    __threadfence_system();
    bool publish = true;
    if (M > 1) publish = atomicAdd(&me->pad_[d], 1u) == (unsigned)M - 1u;  // last CTA of destination d
    if (publish) {
      if (M > 1) {
        me->pad_[d] = 0u;
        __threadfence_system();
      }
      st_release_sys_u64(&static_cast<PeerHdr*>(x.base[d])->flags[x.rank], e);
    }
    // every CTA has read `epoch` before it arrives here, so the last one may advance it
    if (atomicAdd(&me->done, 1u) == gridDim.x - 1u) {
      me->done = 0u;
      *reinterpret_cast<volatile unsigned long long*>(&me->epoch) = e;
      __threadfence();
    }
  }
  // ---- wait for every rank's rows of this epoch (bounded spin, see peer_wait_all_warp), then copy out
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    if (lane < x.world) {
      unsigned spins = 0;
      unsigned long long t0 = 0ull;
      while (ld_relaxed_sys_u64(&me->flags[lane]) < e) {
        __nanosleep(32);
        if (x.timeout_ms != 0u && (++spins & 1023u) == 0u) {
          const unsigned long long now = global_timer_ns();
          if (t0 == 0ull) t0 = now;
          else if (now - t0 > (unsigned long long)x.timeout_ms * 1000000ull) {
            if (status) atomicOr(status, AVSSL_DEVFLAG_PEER_TIMEOUT);
            break;
          }
        }
      }
    }
    __syncwarp();
    __threadfence_system();
  }
  __syncthreads();
  const uint4* src = scatter_slot(x.base[x.rank], slot, x);
  const size_t n16 = (size_t)x.rows_per_rank * row16;
  for (size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x; c < n16; c += (size_t)gridDim.x * blockDim.x)
    out[c] = __ldcg(src + c);
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
  const size_t bytes = (size_t)x->rows_per_rank * x->D * 4;
  if (bytes <= 65536) {  // small blocks (the MoCo keys): one CTA per destination, the shortest dependent chain
    peer_push_kernel<<<x->world, 256, 0, static_cast<cudaStream_t>(stream)>>>(*x, rows);
    AVSSL_LAUNCH_OK("peer_push_kernel");
    return AVSSL_OK;
  }
  const int sms = sm_count() > 0 ? sm_count() : 148;
  // ~128 KiB per CTA: every CTA ends with a system-scope fence that waits for the GPU's outstanding NVLink writes, so
  // few fat CTAs beat many thin ones (256 CTAs of 32 KiB: 32 us for 8 x 1 MiB at N = 8)
  int M = (int)((bytes + 131071) / 131072);
  const int cap = 2 * sms / x->world > 1 ? 2 * sms / x->world : 1;
  if (M > cap) M = cap;
  peer_push_wide_kernel<<<x->world * M, 512, 0, static_cast<cudaStream_t>(stream)>>>(*x, reinterpret_cast<const float4*>(rows), M);
  AVSSL_LAUNCH_OK("peer_push_wide_kernel");
  return AVSSL_OK;
}

extern "C" size_t avssl_peer_scatter_bytes(int rows_per_rank, int64_t row_bytes) {
  if (rows_per_rank < 1 || row_bytes < 16 || row_bytes % 16 != 0) return 0;
  return sizeof(PeerHdr) + 2ull * rows_per_rank * (size_t)row_bytes;
}

extern "C" int avssl_peer_scatter_exchange(const avssl_peer_xchg* x, const void* rows, const int64_t* dest_pos_dev,
                                           const int64_t* dest_pos_host, void* out, uint32_t* status_dev,
                                           void* stream) {
  int rc = peer_check(x, "peer_scatter_exchange");
  if (rc != AVSSL_OK) return rc;
  AVSSL_REQUIRE(rows && (dest_pos_dev || dest_pos_host) && out &&
                    ((reinterpret_cast<uintptr_t>(rows) | reinterpret_cast<uintptr_t>(out)) & 15u) == 0,
                AVSSL_ERR_INVALID_ARGUMENT, "peer_scatter_exchange: null pointer, or rows / out not 16-byte aligned");
  static_assert(sizeof(((PeerHdr*)nullptr)->pad_) >= AVSSL_MAX_PEERS * sizeof(unsigned), "per-destination counters");
  // M CTAs per destination: one for small rows (latency-bound: the shortest dependent chain), more for large
  // rows so that the stores keep NVLink busy (about 64 KiB per CTA), never more than 2 CTAs per SM
  const int sms = sm_count() > 0 ? sm_count() : 148;
  const size_t bytes_per_dest = (size_t)x->rows_per_rank * x->D * 4 / x->world;
  int M = (int)((bytes_per_dest + 65535) / 65536);
  // all CTAs spin on the peers' flags after their stores: the whole grid must be able to be co-resident
  const int cap = 2 * sms / x->world > 1 ? 2 * sms / x->world : 1;
  if (M > cap) M = cap;
  if (M < 1) M = 1;
  if ((size_t)M > (size_t)x->D / 4) M = x->D / 4;
  InlinePos inl;
  bool by_value = dest_pos_host != nullptr && x->rows_per_rank <= kInlinePos;
  if (by_value) {
    const int64_t n_pos = (int64_t)x->world * x->rows_per_rank;
    for (int j = 0; j < x->rows_per_rank; ++j) {
      AVSSL_REQUIRE(dest_pos_host[j] >= 0 && dest_pos_host[j] < n_pos, AVSSL_ERR_INVALID_ARGUMENT,
                    "peer_scatter_exchange: dest_pos[%d] = %lld outside [0, %lld)", j, (long long)dest_pos_host[j],
                    (long long)n_pos);
      inl.pos[j] = (uint32_t)dest_pos_host[j];
    }
  }
  AVSSL_REQUIRE(by_value || dest_pos_dev, AVSSL_ERR_INVALID_ARGUMENT,
                "peer_scatter_exchange: more than %d rows per rank need the device copy of dest_pos", kInlinePos);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (by_value)
    peer_scatter_exchange_kernel<true><<<x->world * M, 512, 0, st>>>(*x, static_cast<const uint4*>(rows), nullptr, inl,
                                                                      static_cast<uint4*>(out), M, status_dev);
  else
    peer_scatter_exchange_kernel<false><<<x->world * M, 512, 0, st>>>(
        *x, static_cast<const uint4*>(rows), reinterpret_cast<const long long*>(dest_pos_dev), inl, static_cast<uint4*>(out),
        M, status_dev);
  AVSSL_LAUNCH_OK("peer_scatter_exchange_kernel");
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
