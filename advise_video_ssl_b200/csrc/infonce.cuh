// Shared declarations for the MoCo l2-norm + logits + InfoNCE kernels (K2+K3).
#pragma once
#include "common.cuh"
#include "peer.cuh"

namespace avssl {

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr int kSimtTileRows = 64;  // queue rows per tile, CUDA-core kernel
constexpr int kTcTileRows = 64;    // queue rows per tile, tcgen05 kernel

// Launch-time description shared by the split kernels (SIMT / tcgen05) and the
// combine kernel.  Partials are kept in the log2 domain:
//   part_m[s][i]      running max of s_ij * log2e/T over the split's queue rows
//   part_l[s][i]      sum_j 2^(s2_ij - m)
//   part_acc[s][i][c] sum_j 2^(s2_ij - m) * queue[j][c]
struct InfoNceParams {
  const float* feat_q;
  const float* keys[AVSSL_MAX_KEYS];
  int n_keys;
  const float* queue;
  int B, D, K;
  float inv_T;
  float* q_out;
  float* loss_out;
  float* dfeat_out;
  float* row_lse_out;
  float* logits_out;
  // workspace
  unsigned* counter;  // [0] finish counter, [1] grid-barrier counter; zero between launches
  float* row_loss;
  float* part_m;
  float* part_l;
  float* part_acc;
  int n_splits;
  int rows_per_split;  // queue rows handled by one split (multiple of the tile)
  // optional fused K4 (models/contrastive.py:263-292): queue[ptr:ptr+B] = keys[0], ptr advanced,
  // performed after every CTA has finished reading the queue.  enq_ptr == nullptr: no enqueue.
  float* queue_rw;
  long long* enq_ptr;
  uint32_t* enq_status;
  // optional fused C3 wait (tcgen05 kernel only): key row i is row
  // (peer_row_idx ? peer_row_idx[i] : peer.rank * B + i) of the peer exchange buffer, read after
  // the merge CTAs have waited for every rank's push of the current epoch.  keys[0] is unused then.
  int use_peer;
  const long long* peer_row_idx;
  // rows written by the fused enqueue: gathered[enq_row_idx[e]] (peer) or keys[0][enq_row_idx ? enq_row_idx[e] : e], e < n_enq
  const long long* enq_row_idx;
  int n_enq;
  // indexed plain keys (tcgen05 kernel, n_keys == 1): keys[0] is [n_key_rows, D] and key row i is
  // keys[0][peer_row_idx[i]] -- the un-shuffle of models/contrastive.py:216-230 folded into the head launch.
  // keys_raw: keys[0] holds the key encoder's RAW output; Normalize (:350) is applied where the rows are read
  // (positive logit and enqueue), with the arithmetic of l2norm_fwd_kernel.
  int n_key_rows;
  int keys_raw;
  // fused C3 push (tcgen05 kernel with use_peer): this rank's RAW key rows [peer.rows_per_rank, D]; one extra CTA
  // of the launch normalises them (eps as Normalize: 0) and stores them into every rank's exchange buffer
  const float* push_feat;
  float push_eps;
  avssl_peer_xchg peer;
  // tcgen05 path in two steps (kPhaseSweep: the sweep kernel + infonce_merge_partials_kernel; kPhaseFinish:
  // infonce_finish_kernel on the same workspace): the sweep only needs feat_q and the queue, so it can run on another
  // stream while the key path (momentum update, shuffle, key encoder) is still in flight; kPhaseFused is the single
  // cooperative launch with the grid barrier in between.
  int phase;
};
constexpr int kPhaseFused = 0, kPhaseSweep = 1, kPhaseFinish = 2;

// 1 / ||row|| computed by ONE warp with a fixed summation order, so that every
// kernel that normalises the same row obtains the same bits.
__device__ __forceinline__ float warp_row_norm(const float* __restrict__ row, int D, int lane) {
  return sqrtf(row_sumsq(row, D, lane));  // Normalize: x / sum(x^2)^(1/2), no eps (models/contrastive.py:929-934)
}

int launch_infonce_simt(const InfoNceParams& p, cudaStream_t s);
int launch_infonce_combine(const InfoNceParams& p, cudaStream_t s);
// tcgen05 path; returns AVSSL_ERR_UNSUPPORTED when the shape does not fit.
int launch_infonce_tc(const InfoNceParams& p, int three_term, cudaStream_t s);
// Two-launch form, behind the sweep on its stream: reduce the n_splits partials of every query row to ONE partial
// (m, l, acc) in (m_out[B], l_out[B], acc_out[B, D]), which kPhaseFinish then reads with n_splits = 1.
int launch_infonce_finish(const InfoNceParams& p, cudaStream_t s);  // kPhaseFinish: p.part_* = the merged partials, n_splits = 1
int launch_infonce_merge_partials(const InfoNceParams& p, float* m_out, float* l_out, float* acc_out, cudaStream_t s);
bool infonce_tc_supported(int B, int D, int K);

}  // namespace avssl
