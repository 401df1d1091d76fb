// K5 / K14 — memory-bank update and gather-dot
// (replaces Memory.get/update models/contrastive.py:966-1036, Memory1D.update
//  :1066-1080, knn_mem_update :131-140 and the mem-mode score gather :429-433).
//
// One warp per item.  Row indices are int64 and are applied bit-exactly; duplicate
// (ind, time) targets resolve the way the reference's CPU index_put does: updates
// are all computed from the OLD bank rows and the LAST occurrence wins.  Only the
// winning warp touches its row (losers exit before reading), so one launch is
// race-free without a grid barrier.
#include "common.cuh"

namespace avssl {

struct BankArgs {
  float* bank;
  int64_t L;
  int duration, D;
  const float* mem;
  const int64_t* ind;
  const int64_t* time_i;  // may be null (== 0)
  const float* time_f;    // interp only
  int n;
  float m, om;
  int interp;
  uint32_t* status;
};

// Target row (flattened ind*duration + t) of entry e in the reference's statement
// order: entries [0,n) are the t0 writes (:1025), [n,2n) the t1 writes (:1026).
__device__ __forceinline__ int64_t bank_target(const BankArgs& a, int e, float* w_out) {
  const int i = e < a.n ? e : e - a.n;
  int64_t r = a.ind[i];
  if (r < 0) r += a.L;
  if (r < 0 || r >= a.L) return -1;
  int64_t t;
  float w = 1.f;
  if (a.interp) {
    const float tf = a.time_f[i];
    int64_t t0 = (int64_t)floorf(tf);
    t0 = t0 < 0 ? 0 : (t0 > a.duration - 1 ? a.duration - 1 : t0);
    int64_t t1 = t0 + 1;
    t1 = t1 > a.duration - 1 ? a.duration - 1 : t1;
    const float w_t1 = 1.f - (tf - (float)t0);  // "hack for inverse" (:1003-1004)
    const float w_t0 = 1.f - w_t1;
    if (e < a.n) {
      t = t0;
      w = w_t0;
    } else {
      t = t1;
      w = w_t1;
    }
  } else {
    t = a.time_i ? a.time_i[i] : 0;
    if (t < 0) t += a.duration;
    if (t < 0 || t >= a.duration) return -1;
  }
  if (w_out) *w_out = w;
  return r * a.duration + t;
}

__global__ void __launch_bounds__(256) membank_update_kernel(const BankArgs a) {
  const int lane = threadIdx.x & 31;
  const int e = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int n_entries = a.interp ? 2 * a.n : a.n;
  if (e >= n_entries) return;
  float w = 1.f;
  const int64_t tgt = bank_target(a, e, &w);
  if (tgt < 0) {
    if (lane == 0 && a.status) atomicOr(a.status, AVSSL_DEVFLAG_BAD_INDEX);
    return;
  }
  // last occurrence wins
  bool later = false;
  for (int e2 = e + 1 + lane; e2 < n_entries; e2 += 32) later |= (bank_target(a, e2, nullptr) == tgt);
  if (__any_sync(0xffffffffu, later)) return;

  const int i = e < a.n ? e : e - a.n;
  float* row = a.bank + tgt * a.D;
  const float* src = a.mem + (int64_t)i * a.D;
  // upd = mem [* w] * momentum + old * (1 - momentum), separately rounded as in ATen
  float ss = 0.f;
  for (int c = lane; c < a.D; c += 32) {
    float x = src[c];
    if (a.interp) x = __fmul_rn(x, w);
    const float u = __fadd_rn(__fmul_rn(x, a.m), __fmul_rn(row[c], a.om));
    ss = fmaf(u, u, ss);
  }
  ss = warp_sum(ss);
  const float nrm = sqrtf(ss);
  for (int c = lane; c < a.D; c += 32) {
    float x = src[c];
    if (a.interp) x = __fmul_rn(x, w);
    const float u = __fadd_rn(__fmul_rn(x, a.m), __fmul_rn(row[c], a.om));
    row[c] = u / nrm;  // Normalize: no eps (:929-934)
  }
}

struct DotArgs {
  const float* bank;
  int64_t L;
  int duration, D;
  const float* q;
  const int64_t* ind;
  const int64_t* time_i;
  const float* time_f;
  int B, Kp;
  float inv_T;
  int interp;
  float* prod;
  uint32_t* status;
};

// prod[n,k] = q_n . bank[ind[n,k], time[n,k]] / T without materialising [B,K+1,D].
__global__ void __launch_bounds__(256) membank_gather_dot_kernel(const DotArgs a) {
  extern __shared__ float qs[];  // this CTA's query row
  const int n = blockIdx.y;
  for (int c = threadIdx.x; c < a.D; c += blockDim.x) qs[c] = a.q[(int64_t)n * a.D + c];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int k = blockIdx.x * nw + warp; k < a.Kp; k += gridDim.x * nw) {
    const int64_t flat = (int64_t)n * a.Kp + k;
    int64_t r = a.ind[flat];
    if (r < 0) r += a.L;
    if (r < 0 || r >= a.L) {
      if (lane == 0 && a.status) atomicOr(a.status, AVSSL_DEVFLAG_BAD_INDEX);
      continue;
    }
    float dot = 0.f;
    if (a.interp) {
      const float tf = a.time_f[flat];
      int64_t t0 = (int64_t)floorf(tf);
      t0 = t0 < 0 ? 0 : (t0 > a.duration - 1 ? a.duration - 1 : t0);
      int64_t t1 = t0 + 1;
      t1 = t1 > a.duration - 1 ? a.duration - 1 : t1;
      const float w1 = 1.f - (tf - (float)t0);
      const float w0 = 1.f - w1;
      const float* r0 = a.bank + (r * a.duration + t0) * a.D;
      const float* r1 = a.bank + (r * a.duration + t1) * a.D;
      for (int c = lane; c < a.D; c += 32)
        dot = fmaf(qs[c], __fadd_rn(__fmul_rn(r0[c], w0), __fmul_rn(r1[c], w1)), dot);
    } else {
      int64_t t = a.time_i ? a.time_i[flat] : 0;
      if (t < 0) t += a.duration;
      if (t < 0 || t >= a.duration) {
        if (lane == 0 && a.status) atomicOr(a.status, AVSSL_DEVFLAG_BAD_INDEX);
        continue;
      }
      const float* row = a.bank + (r * a.duration + t) * a.D;
      if ((a.D & 3) == 0) {
        for (int c = lane * 4; c < a.D; c += 128) {
          const float4 v = *reinterpret_cast<const float4*>(row + c);
          dot = fmaf(qs[c], v.x, dot);
          dot = fmaf(qs[c + 1], v.y, dot);
          dot = fmaf(qs[c + 2], v.z, dot);
          dot = fmaf(qs[c + 3], v.w, dot);
        }
      } else {
        for (int c = lane; c < a.D; c += 32) dot = fmaf(qs[c], row[c], dot);
      }
    }
    dot = warp_sum(dot);
    if (lane == 0) a.prod[flat] = dot * a.inv_T;
  }
}

}  // namespace avssl

using namespace avssl;

extern "C" int avssl_membank_update(float* bank, int64_t L, int duration, int D, const float* mem,
                                    const int64_t* ind, const int64_t* time_i64, const float* time_f32, int n,
                                    float momentum, float one_minus_momentum, int interp, uint32_t* status_dev,
                                    void* stream) {
  AVSSL_REQUIRE(bank && mem && ind, AVSSL_ERR_INVALID_ARGUMENT, "membank_update: null pointer");
  AVSSL_REQUIRE(L > 0 && duration > 0 && D > 0 && n >= 0, AVSSL_ERR_INVALID_ARGUMENT, "membank_update: bad sizes");
  AVSSL_REQUIRE(!interp || time_f32, AVSSL_ERR_INVALID_ARGUMENT, "membank_update: interp needs float times");
  if (n == 0) return AVSSL_OK;
  BankArgs a{bank, L, duration, D, mem, ind, time_i64, time_f32, n, momentum, one_minus_momentum, interp ? 1 : 0, status_dev};
  const int entries = interp ? 2 * n : n;
  membank_update_kernel<<<(entries + 7) / 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(a);
  AVSSL_LAUNCH_OK("membank_update_kernel");
  return AVSSL_OK;
}

extern "C" int avssl_membank_gather_dot(const float* bank, int64_t L, int duration, int D, const float* q,
                                        const int64_t* ind, const int64_t* time_i64, const float* time_f32, int B,
                                        int Kp, float T, int interp, float* prod, uint32_t* status_dev, void* stream) {
  AVSSL_REQUIRE(bank && q && ind && prod, AVSSL_ERR_INVALID_ARGUMENT, "membank_gather_dot: null pointer");
  AVSSL_REQUIRE(L > 0 && duration > 0 && D > 0 && B > 0 && Kp > 0 && T > 0.f, AVSSL_ERR_INVALID_ARGUMENT,
                "membank_gather_dot: bad sizes");
  AVSSL_REQUIRE(B <= 65535, AVSSL_ERR_UNSUPPORTED, "membank_gather_dot: B > 65535");
  AVSSL_REQUIRE(!interp || time_f32, AVSSL_ERR_INVALID_ARGUMENT, "membank_gather_dot: interp needs float times");
  DotArgs a{bank, L, duration, D, q, ind, time_i64, time_f32, B, Kp, 1.0f / T, interp ? 1 : 0, prod, status_dev};
  int gx = (Kp + 7) / 8;
  const int cap = (sm_count() > 0 ? sm_count() : 148) * 8 / (B < 8 ? B : 8);
  if (gx > cap) gx = cap < 1 ? 1 : cap;
  dim3 grid(gx, B);
  membank_gather_dot_kernel<<<grid, 256, sizeof(float) * D, static_cast<cudaStream_t>(stream)>>>(a);
  AVSSL_LAUNCH_OK("membank_gather_dot_kernel");
  return AVSSL_OK;
}
