// Shared building blocks of the CUDA-core (SIMT) tiled kernels: a 64x64 score tile
// S = A . B^T from two shared-memory operands and the follow-up P . B product.
// 256 threads; thread (ty = tid/16, tx = tid%16) owns rows ty*4+ii.
#pragma once
#include "common.cuh"

namespace avssl {

constexpr int kTileJ = 64;   // queue rows per tile
constexpr int kTileI = 64;   // query rows per CTA
constexpr int kSimtThreads = 256;
constexpr int kPsStride = kTileJ + 4;

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool valid) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

template <int DP>
__device__ __forceinline__ void load_tile(float* ks, const float* __restrict__ queue, int D, int j0, int j_end) {
  constexpr int KS = DP + 4;
  constexpr int CH = DP / 4;  // 16-byte chunks per row
  for (int idx = threadIdx.x; idx < kTileJ * CH; idx += kSimtThreads) {
    const int r = idx / CH, ch = idx % CH;
    const int j = j0 + r;
    const bool valid = (j < j_end) && (ch * 4 < D);
    const float* src = valid ? queue + (size_t)j * D + ch * 4 : queue;
    cp_async16(ks + r * KS + ch * 4, src, valid);
  }
}


// S[ii][jj] = sum_c qs[ty*4+ii][c] * ks[tx+16*jj][c]
template <int DP>
__device__ __forceinline__ void simt_s_tile(const float* __restrict__ qs, const float* __restrict__ ks, int ty, int tx,
                                            float (&s)[4][4]) {
  constexpr int KS = DP + 4;
#pragma unroll
  for (int ii = 0; ii < 4; ++ii)
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) s[ii][jj] = 0.f;
#pragma unroll 4
  for (int c = 0; c < DP; c += 4) {
    float4 qv[4], kv[4];
#pragma unroll
    for (int ii = 0; ii < 4; ++ii) qv[ii] = *reinterpret_cast<const float4*>(qs + (ty * 4 + ii) * KS + c);
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) kv[jj] = *reinterpret_cast<const float4*>(ks + (tx + 16 * jj) * KS + c);
#pragma unroll
    for (int ii = 0; ii < 4; ++ii)
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        s[ii][jj] = fmaf(qv[ii].x, kv[jj].x, s[ii][jj]);
        s[ii][jj] = fmaf(qv[ii].y, kv[jj].y, s[ii][jj]);
        s[ii][jj] = fmaf(qv[ii].z, kv[jj].z, s[ii][jj]);
        s[ii][jj] = fmaf(qv[ii].w, kv[jj].w, s[ii][jj]);
      }
  }
}

// acc[ii][cc] += sum_j ps[ty*4+ii][j] * ks[j][tx*4 + 64*cc .. +3]
template <int DP>
__device__ __forceinline__ void simt_pv_tile(const float* __restrict__ ps, const float* __restrict__ ks, int ty, int tx,
                                             float4 (&acc)[4][DP / 64]) {
  constexpr int KS = DP + 4;
  constexpr int CC = DP / 64;
#pragma unroll 2
  for (int j0 = 0; j0 < kTileJ; j0 += 4) {
    float4 pv[4];
#pragma unroll
    for (int ii = 0; ii < 4; ++ii) pv[ii] = *reinterpret_cast<const float4*>(ps + (ty * 4 + ii) * kPsStride + j0);
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      float4 kv[CC];
#pragma unroll
      for (int cc = 0; cc < CC; ++cc) kv[cc] = *reinterpret_cast<const float4*>(ks + (j0 + jj) * KS + tx * 4 + 64 * cc);
#pragma unroll
      for (int ii = 0; ii < 4; ++ii) {
        const float pj = jj == 0 ? pv[ii].x : jj == 1 ? pv[ii].y : jj == 2 ? pv[ii].z : pv[ii].w;
#pragma unroll
        for (int cc = 0; cc < CC; ++cc) {
          acc[ii][cc].x = fmaf(pj, kv[cc].x, acc[ii][cc].x);
          acc[ii][cc].y = fmaf(pj, kv[cc].y, acc[ii][cc].y);
          acc[ii][cc].z = fmaf(pj, kv[cc].z, acc[ii][cc].z);
          acc[ii][cc].w = fmaf(pj, kv[cc].w, acc[ii][cc].w);
        }
      }
    }
  }
}

constexpr size_t simt_smem_bytes(int DP, bool with_ps) {
  return sizeof(float) * ((size_t)kTileI * (DP + 4) + 2 * (size_t)kTileJ * (DP + 4) + (with_ps ? kTileI * kPsStride : 0));
}

}  // namespace avssl
