// Device side of the NVLink peer-memory row exchange (C3: the cross-GPU key gather of
// models/contrastive.py:216-230 without a collective kernel).
//
// Every rank owns one exchange buffer, mapped into every other rank's address space
// (CUDA IPC, see peer.cu):
//
//   offset   0  u64 flags[16]   flags[src] = epoch of the last push completed by rank `src`
//                               (written by `src` over NVLink, release at system scope)
//   offset 128  u64 epoch       pushes issued by the OWNER so far (local)
//   offset 136  u32 done        CTA-finish counter of the running push (local, self-resetting)
//   offset 256  float payload[2][world * rows * D]   slot = epoch & 1
//
// push(e):  rank r stores its [rows, D] block into payload[e & 1][r] of EVERY rank (one CTA per
//           destination), fences at system scope and then publishes flags[r] = e there.
// wait(e):  one consumer warp polls all `world` local flags until they are >= e, then fences (system scope).
//
// Two slots suffice: rank A can issue push e+2 only after its own consumer of epoch e+1 has
// finished, which waited for B's push e+1, which B issued after ITS consumer of epoch e was done
// with slot e & 1 (same stream).  Flags are monotonic, so a peer running one step ahead is harmless.
#pragma once
#include "common.cuh"

namespace avssl {

struct PeerHdr {
  unsigned long long flags[AVSSL_MAX_PEERS];
  unsigned long long epoch;
  unsigned int done;
  unsigned int pad_[29];
};
static_assert(sizeof(PeerHdr) == 256, "exchange header is 256 bytes");

__device__ __forceinline__ void st_release_sys_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__device__ __forceinline__ float* peer_payload(void* base, int slot, const avssl_peer_xchg& x) {
  return reinterpret_cast<float*>(static_cast<char*>(base) + sizeof(PeerHdr)) +
         (size_t)slot * x.world * x.rows_per_rank * x.D;
}

// All threads of one CTA: push this rank's rows to rank `dst`.  `world` CTAs (dst = 0..world-1)
// make one push; the last of them to finish advances the local epoch.  `s_epoch` is a shared
// scratch word.  rows * D must be a multiple of 4 and `rows` 16-byte aligned (checked by the host).
// kNormalize: the rows are raw features and what travels is x / max(||x||, eps), computed with the
// arithmetic of l2norm_fwd_kernel (same summation order, true division), so the receivers hold the
// bits a separate Normalize launch would have produced; the CTA whose destination is this rank
// also stores them to y_local (may be null).
template <bool kNormalize>
__device__ __forceinline__ void peer_push_cta(const avssl_peer_xchg& x, const float* __restrict__ rows, int dst,
                                              unsigned long long* s_epoch, float eps = 0.f,
                                              float* __restrict__ y_local = nullptr) {
  PeerHdr* me = static_cast<PeerHdr*>(x.base[x.rank]);
  if (threadIdx.x == 0) *s_epoch = *reinterpret_cast<volatile unsigned long long*>(&me->epoch) + 1ull;
  __syncthreads();
  const unsigned long long e = *s_epoch;
  float* out_f = peer_payload(x.base[dst], (int)(e & 1ull), x) + (size_t)x.rank * x.rows_per_rank * x.D;
  if (kNormalize) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
    const bool keep = y_local != nullptr && dst == x.rank;
    for (int r = warp; r < x.rows_per_rank; r += n_warps) {
      const float* xr = rows + (size_t)r * x.D;
      const float den = fmaxf(sqrtf(row_sumsq(xr, x.D, lane)), eps);
      for (int c = lane; c < x.D; c += 32) {
        const float y = xr[c] / den;
        out_f[(size_t)r * x.D + c] = y;
        if (keep) y_local[(size_t)r * x.D + c] = y;
      }
    }
  } else {
    const int n4 = x.rows_per_rank * x.D / 4;
    const float4* src = reinterpret_cast<const float4*>(rows);
    float4* out = reinterpret_cast<float4*>(out_f);
    for (int i = threadIdx.x; i < n4; i += blockDim.x) out[i] = __ldg(src + i);
  }
  // Publication: the CTA barrier orders every thread's payload stores before thread 0's system-scope fence
  // (release patterns are cumulative over barrier synchronisation in the PTX memory model -- the pattern of
  // cooperative-groups grid sync), so ONE fence and one release store publish the whole block; a fence per
  // thread cost a second NVLink round trip on the step's critical path.
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    st_release_sys_u64(&static_cast<PeerHdr*>(x.base[dst])->flags[x.rank], e);
    // every CTA has read `epoch` before it arrives here, so the last one may advance it
    const unsigned prev = atomicAdd(&me->done, 1u);
    if (prev == (unsigned)x.world - 1u) {
      me->done = 0u;
      *reinterpret_cast<volatile unsigned long long*>(&me->epoch) = e;
      __threadfence();
    }
  }
}

// All threads of ONE CTA: normalise this rank's raw rows and store them to EVERY rank of the exchange (one push by
// a single CTA: the extra CTA of the head launch).  Same arithmetic as peer_push_cta<true>; `world` threads publish
// the flags after the barrier, the epoch advances locally.
__device__ __forceinline__ void peer_push_all_cta(const avssl_peer_xchg& x, const float* __restrict__ rows, float eps,
                                                  unsigned long long* s_epoch) {
  PeerHdr* me = static_cast<PeerHdr*>(x.base[x.rank]);
  if (threadIdx.x == 0) *s_epoch = *reinterpret_cast<volatile unsigned long long*>(&me->epoch) + 1ull;
  __syncthreads();
  const unsigned long long e = *s_epoch;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
  const size_t block_off = (size_t)x.rank * x.rows_per_rank * x.D;
  for (int r = warp; r < x.rows_per_rank; r += n_warps) {
    const float* xr = rows + (size_t)r * x.D;
    const float den = fmaxf(sqrtf(row_sumsq(xr, x.D, lane)), eps);
    for (int c = lane; c < x.D; c += 32) {
      const float y = xr[c] / den;
      for (int d = 0; d < x.world; ++d) peer_payload(x.base[d], (int)(e & 1ull), x)[block_off + (size_t)r * x.D + c] = y;
    }
  }
  __syncthreads();
  if ((int)threadIdx.x < x.world) {
    __threadfence_system();
    st_release_sys_u64(&static_cast<PeerHdr*>(x.base[threadIdx.x])->flags[x.rank], e);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    *reinterpret_cast<volatile unsigned long long*>(&me->epoch) = e;
    __threadfence();
  }
}

__device__ __forceinline__ unsigned long long ld_relaxed_sys_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// wait, by ONE WARP (all 32 lanes call it): until every rank's push of the current local epoch has
// landed here.  Lane r watches rank r's flag with relaxed loads, so the wait costs one flag latency
// whatever the world size (W sequential ld.acquire.sys cost +14 us per step at W = 4), and a single
// system-scope fence orders the payload reads that follow the caller's __syncthreads().
// The spin is bounded: after x.timeout_ms (0 = never) without the flag the lane gives up, ors
// AVSSL_DEVFLAG_PEER_TIMEOUT into *status (may be null) and the kernel carries on with whatever the
// slot holds -- a dead peer then costs one garbage step that the host sees in the status word
// instead of a hung cooperative kernel.  Returns the payload slot to read.
__device__ __forceinline__ int peer_wait_all_warp(const avssl_peer_xchg& x, uint32_t* status) {
  PeerHdr* me = static_cast<PeerHdr*>(x.base[x.rank]);
  const int lane = threadIdx.x & 31;
  const unsigned long long e = *reinterpret_cast<volatile unsigned long long*>(&me->epoch);
  if (lane < x.world) {
    unsigned spins = 0;
    unsigned long long t0 = 0ull;
    while (ld_relaxed_sys_u64(&me->flags[lane]) < e) {
      __nanosleep(32);
      if (x.timeout_ms != 0u && (++spins & 1023u) == 0u) {  // look at the clock every ~1000 polls
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
  return (int)(e & 1ull);
}

// Host-side validation of an exchange descriptor (peer.cu).
int peer_check(const avssl_peer_xchg* x, const char* who);

}  // namespace avssl
