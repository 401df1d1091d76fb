// K2+K3 front door: argument checks, workspace carving, implementation choice.
#include <string.h>

#include "infonce.cuh"

using namespace avssl;

namespace {

constexpr size_t kAlign = 256;
inline size_t align_up(size_t v) { return (v + kAlign - 1) / kAlign * kAlign; }

int max_splits() {
  int sms = sm_count();
  if (sms <= 0) {
    cudaGetLastError();
    sms = 148;
  }
  return 2 * sms;
}

struct Carve {
  size_t counter, row_loss, part_m, part_l, part_acc, one_m, one_l, one_acc, total;
};
Carve carve(int B, int D, int n_keys, int S) {
  Carve c;
  size_t off = 0;
  c.counter = off;
  off += kAlign;
  c.row_loss = off;
  off += align_up(sizeof(float) * (size_t)n_keys * B);
  c.part_m = off;
  off += align_up(sizeof(float) * (size_t)S * B);
  c.part_l = off;
  off += align_up(sizeof(float) * (size_t)S * B);
  c.part_acc = off;
  off += align_up(sizeof(float) * (size_t)S * B * D);
  // two-launch form: the partials of every row merged to one (infonce_merge_partials_kernel)
  c.one_m = off;
  off += align_up(sizeof(float) * (size_t)B);
  c.one_l = off;
  off += align_up(sizeof(float) * (size_t)B);
  c.one_acc = off;
  off += align_up(sizeof(float) * (size_t)B * D);
  c.total = off;
  return c;
}

}  // namespace

extern "C" size_t avssl_moco_infonce_workspace_bytes(int B, int D, int K, int n_keys) {
  if (B <= 0 || D <= 0 || K <= 0 || n_keys <= 0) return 0;
  return carve(B, D, n_keys, max_splits()).total;
}

namespace {

int moco_infonce_run(const float* feat_q, const float* const* keys_host, int n_keys, const float* queue,
                     float* queue_rw, int64_t* ptr_dev, uint32_t* status_dev, int B, int D, int K, float T,
                     float* q_out, float* loss_out, float* dfeat_out, float* row_lse_out, float* logits_out,
                     void* workspace, size_t workspace_bytes, int impl, void* stream,
                     const avssl_peer_xchg* peer = nullptr, const int64_t* peer_row_idx = nullptr,
                     const int64_t* enq_row_idx = nullptr, int n_enq = 0, int n_key_rows = 0, int keys_raw = 0,
                     const float* push_feat = nullptr) {
  // impl carries the two-launch form in its upper bits (avssl_b200.h: AVSSL_HEAD_SWEPT, avssl_moco_infonce_sweep)
  const int phase = (impl & 0x100) ? kPhaseFinish : ((impl & 0x200) ? kPhaseSweep : kPhaseFused);
  const int sweep_ctas = (impl >> 16) & 0xffff;
  impl &= 0xff;
  AVSSL_REQUIRE(feat_q && queue && workspace, AVSSL_ERR_INVALID_ARGUMENT, "moco_infonce: null pointer");
  AVSSL_REQUIRE(phase == kPhaseSweep || ((keys_host || peer) && q_out && loss_out && dfeat_out),
                AVSSL_ERR_INVALID_ARGUMENT, "moco_infonce: null pointer");
  AVSSL_REQUIRE(B > 0 && D > 0 && K > 0, AVSSL_ERR_INVALID_ARGUMENT, "moco_infonce: bad sizes B=%d D=%d K=%d", B, D, K);
  AVSSL_REQUIRE(n_keys >= 1 && n_keys <= AVSSL_MAX_KEYS, AVSSL_ERR_INVALID_ARGUMENT,
                "moco_infonce: n_keys=%d outside [1, %d]", n_keys, AVSSL_MAX_KEYS);
  AVSSL_REQUIRE(T > 0.f, AVSSL_ERR_INVALID_ARGUMENT, "moco_infonce: temperature must be positive");
  AVSSL_REQUIRE((reinterpret_cast<uintptr_t>(feat_q) & 15u) == 0 && (reinterpret_cast<uintptr_t>(queue) & 15u) == 0 &&
                    (reinterpret_cast<uintptr_t>(workspace) & 255u) == 0,
                AVSSL_ERR_INVALID_ARGUMENT, "moco_infonce: feat_q/queue need 16-byte, workspace 256-byte alignment");
  const int sms = sm_count();
  AVSSL_REQUIRE(sms > 0, AVSSL_ERR_CUDA, "moco_infonce: no CUDA device (there is no CPU fallback)");
  cudaStream_t s = static_cast<cudaStream_t>(stream);

  InfoNceParams p;
  p.feat_q = feat_q;
  for (int k = 0; k < AVSSL_MAX_KEYS; ++k) p.keys[k] = (k < n_keys && !peer && keys_host) ? keys_host[k] : nullptr;
  for (int k = 0; k < n_keys && !peer && phase != kPhaseSweep; ++k)
    AVSSL_REQUIRE(p.keys[k], AVSSL_ERR_INVALID_ARGUMENT, "moco_infonce: keys[%d] is null", k);
  p.use_peer = peer ? 1 : 0;
  p.peer_row_idx = reinterpret_cast<const long long*>(peer_row_idx);
  p.enq_row_idx = reinterpret_cast<const long long*>(enq_row_idx);
  p.n_enq = (ptr_dev && n_enq > 0) ? n_enq : B;
  p.n_key_rows = n_key_rows > 0 ? n_key_rows : B;
  p.keys_raw = keys_raw ? 1 : 0;
  p.push_feat = phase == kPhaseSweep ? nullptr : push_feat;
  p.phase = phase;
  p.push_eps = 0.f;  // Normalize has no epsilon (models/contrastive.py:929-934)
  AVSSL_REQUIRE(!push_feat || (peer && (reinterpret_cast<uintptr_t>(push_feat) & 15u) == 0), AVSSL_ERR_INVALID_ARGUMENT,
                "moco_infonce_peer: push_feat needs an exchange descriptor and 16-byte alignment");
  const bool indexed = !peer && (peer_row_idx || enq_row_idx || keys_raw || p.n_key_rows != B);
  AVSSL_REQUIRE(!indexed || n_keys == 1, AVSSL_ERR_INVALID_ARGUMENT,
                "moco_infonce_indexed: one key tensor only (got %d)", n_keys);
  memset(&p.peer, 0, sizeof(p.peer));
  if (peer) {
    const int rc = peer_check(peer, "moco_infonce_peer");
    if (rc != AVSSL_OK) return rc;
    AVSSL_REQUIRE(n_keys == 1 && peer->D == D && (peer_row_idx || peer->rows_per_rank == B), AVSSL_ERR_INVALID_ARGUMENT,
                  "moco_infonce_peer: exchange is [%d x %d] per rank, head has B=%d D=%d (one key tensor only)",
                  peer->rows_per_rank, peer->D, B, D);
    p.peer = *peer;
  }
  p.n_keys = n_keys;
  p.queue = queue;
  p.B = B;
  p.D = D;
  p.K = K;
  p.inv_T = 1.0f / T;
  p.q_out = q_out;
  p.loss_out = loss_out;
  p.dfeat_out = dfeat_out;
  p.row_lse_out = row_lse_out;
  p.logits_out = logits_out;
  p.queue_rw = queue_rw;
  p.enq_ptr = reinterpret_cast<long long*>(ptr_dev);
  p.enq_status = status_dev;
  if (ptr_dev) {
    // models/contrastive.py:284  assert self.k % num_items == 0
    AVSSL_REQUIRE(K % p.n_enq == 0, AVSSL_ERR_INVALID_ARGUMENT, "moco_infonce: K=%d is not a multiple of the %d enqueued rows", K,
                  p.n_enq);
    AVSSL_REQUIRE(enq_row_idx != nullptr || p.n_enq == (peer ? peer->rows_per_rank : B), AVSSL_ERR_INVALID_ARGUMENT,
                  "moco_infonce: n_enq=%d needs a row list into the key buffer", p.n_enq);
    AVSSL_REQUIRE(queue_rw && (peer || (reinterpret_cast<uintptr_t>(p.keys[0]) & 15u) == 0) && D % 4 == 0,
                  AVSSL_ERR_INVALID_ARGUMENT, "moco_infonce: fused enqueue needs a writable queue and 16-byte aligned keys[0]");
  }

  int use = impl;
  if (use == AVSSL_IMPL_AUTO) use = infonce_tc_supported(B, D, K) ? AVSSL_IMPL_TC3X : AVSSL_IMPL_SIMT;
  AVSSL_REQUIRE(use == AVSSL_IMPL_SIMT || use == AVSSL_IMPL_TC3X || use == AVSSL_IMPL_TC1X,
                AVSSL_ERR_INVALID_ARGUMENT, "moco_infonce: unknown impl %d", impl);
  AVSSL_REQUIRE(!indexed || use != AVSSL_IMPL_SIMT, AVSSL_ERR_UNSUPPORTED,
                "moco_infonce_indexed: row indices / raw keys need the tcgen05 kernel (D in {32,64,96,128}); normalise and "
                "gather the keys first and call the plain entry point instead");
  AVSSL_REQUIRE(phase == kPhaseFused || use != AVSSL_IMPL_SIMT, AVSSL_ERR_UNSUPPORTED,
                "moco_infonce: the two-launch form (sweep, then AVSSL_HEAD_SWEPT) needs the tcgen05 kernel (D in {32,64,96,128})");
  AVSSL_REQUIRE(!peer || use != AVSSL_IMPL_SIMT, AVSSL_ERR_UNSUPPORTED,
                "moco_infonce_peer: the fused exchange wait needs the tcgen05 kernel (D in {32,64,96,128}); "
                "call avssl_peer_wait_gather and the plain entry point instead");

  // split the queue over the SMs in whole tiles
  const int tile = (use == AVSSL_IMPL_SIMT) ? kSimtTileRows : kTcTileRows;
  const int n_tiles = (K + tile - 1) / tile;
  const int row_blocks = (use == AVSSL_IMPL_SIMT) ? (B + 63) / 64 : (B + 127) / 128;
  // (the fused key push takes one more column of CTAs: leave it an SM, the launch is cooperative)
  // (two launches: the caller may give the sweep fewer CTAs than SMs, leaving the rest to whatever runs beside it)
  int S = (push_feat && phase == kPhaseFused ? sms - row_blocks : sms) / row_blocks;
  if (phase != kPhaseFused && sweep_ctas > 0 && sweep_ctas / row_blocks < S) S = sweep_ctas / row_blocks;
  if (S < 1) S = 1;
  if (S > n_tiles) S = n_tiles;
  const int tiles_per_split = (n_tiles + S - 1) / S;
  S = (n_tiles + tiles_per_split - 1) / tiles_per_split;  // drop empty splits
  p.n_splits = S;
  p.rows_per_split = tiles_per_split * tile;

  const Carve c = carve(B, D, n_keys, S);
  AVSSL_REQUIRE(workspace_bytes >= c.total, AVSSL_ERR_WORKSPACE,
                "moco_infonce: workspace of %zu bytes, need %zu", workspace_bytes, c.total);
  char* w = static_cast<char*>(workspace);
  p.counter = reinterpret_cast<unsigned*>(w + c.counter);
  p.row_loss = reinterpret_cast<float*>(w + c.row_loss);
  p.part_m = reinterpret_cast<float*>(w + c.part_m);
  p.part_l = reinterpret_cast<float*>(w + c.part_l);
  p.part_acc = reinterpret_cast<float*>(w + c.part_acc);

  if (use != AVSSL_IMPL_SIMT) {
    float* one_m = reinterpret_cast<float*>(w + c.one_m);
    float* one_l = reinterpret_cast<float*>(w + c.one_l);
    float* one_acc = reinterpret_cast<float*>(w + c.one_acc);
    if (phase == kPhaseSweep) {  // the sweep, then (same stream) its partials reduced to one per query row
      const int rc = launch_infonce_tc(p, use == AVSSL_IMPL_TC3X ? 1 : 0, s);
      return rc != AVSSL_OK ? rc : launch_infonce_merge_partials(p, one_m, one_l, one_acc, s);
    }
    if (phase == kPhaseFinish) {  // key term, loss, gradient, enqueue against the one merged partial per row
      p.part_m = one_m;
      p.part_l = one_l;
      p.part_acc = one_acc;
      p.n_splits = 1;
      return launch_infonce_finish(p, s);
    }
    // kPhaseFused: one cooperative launch: sweep + grid barrier + merge (+ the queue ring write of K4)
    return launch_infonce_tc(p, use == AVSSL_IMPL_TC3X ? 1 : 0, s);
  }
  int rc = launch_infonce_simt(p, s);
  if (rc != AVSSL_OK) return rc;
  rc = launch_infonce_combine(p, s);
  if (rc != AVSSL_OK || !ptr_dev) return rc;
  return avssl_queue_enqueue(queue_rw, ptr_dev, p.keys[0], B, K, D, status_dev, stream);
}

}  // namespace

extern "C" int avssl_moco_infonce_sweep(const float* feat_q, const float* queue, int B, int D, int K, float T, int n_keys,
                                        float* logits_out, void* workspace, size_t workspace_bytes, int impl,
                                        int sweep_ctas, void* stream) {
  AVSSL_REQUIRE(impl >= 0 && impl <= 0xff && sweep_ctas >= 0 && sweep_ctas <= 0xffff, AVSSL_ERR_INVALID_ARGUMENT,
                "moco_infonce_sweep: bad impl %d / sweep_ctas %d", impl, sweep_ctas);
  return moco_infonce_run(feat_q, nullptr, n_keys, queue, nullptr, nullptr, nullptr, B, D, K, T, nullptr, nullptr, nullptr,
                          nullptr, logits_out, workspace, workspace_bytes, impl | 0x200 | (sweep_ctas << 16), stream);
}

extern "C" int avssl_moco_infonce_fwd_bwd(const float* feat_q, const float* const* keys_host, int n_keys,
                                          const float* queue, int B, int D, int K, float T, float* q_out,
                                          float* loss_out, float* dfeat_out, float* row_lse_out,
                                          float* logits_out, void* workspace, size_t workspace_bytes,
                                          int impl, void* stream) {
  return moco_infonce_run(feat_q, keys_host, n_keys, queue, nullptr, nullptr, nullptr, B, D, K, T, q_out, loss_out,
                          dfeat_out, row_lse_out, logits_out, workspace, workspace_bytes, impl, stream);
}

extern "C" int avssl_moco_infonce_fwd_bwd_enqueue(const float* feat_q, const float* const* keys_host, int n_keys,
                                                  float* queue, int64_t* ptr_dev, uint32_t* status_dev, int B,
                                                  int D, int K, float T, float* q_out, float* loss_out,
                                                  float* dfeat_out, float* row_lse_out, float* logits_out,
                                                  void* workspace, size_t workspace_bytes, int impl, void* stream) {
  AVSSL_REQUIRE(ptr_dev, AVSSL_ERR_INVALID_ARGUMENT, "moco_infonce_enqueue: ptr_dev is null");
  return moco_infonce_run(feat_q, keys_host, n_keys, queue, queue, ptr_dev, status_dev, B, D, K, T, q_out, loss_out,
                          dfeat_out, row_lse_out, logits_out, workspace, workspace_bytes, impl, stream);
}

extern "C" int avssl_moco_infonce_fwd_bwd_enqueue_peer(const float* feat_q, const avssl_peer_xchg* x,
                                                       const int64_t* row_idx, const int64_t* enq_row_idx, int n_enq,
                                                       const float* push_feat, float* queue, int64_t* ptr_dev,
                                                       uint32_t* status_dev, int B, int D, int K, float T,
                                                       float* q_out, float* loss_out, float* dfeat_out,
                                                       float* row_lse_out, float* logits_out, void* workspace,
                                                       size_t workspace_bytes, int impl, void* stream) {
  AVSSL_REQUIRE(x, AVSSL_ERR_INVALID_ARGUMENT, "moco_infonce_peer: exchange descriptor is null");
  return moco_infonce_run(feat_q, nullptr, 1, queue, ptr_dev ? queue : nullptr, ptr_dev, status_dev, B, D, K, T, q_out,
                          loss_out, dfeat_out, row_lse_out, logits_out, workspace, workspace_bytes, impl, stream, x,
                          row_idx, enq_row_idx, n_enq, 0, 0, push_feat);
}

extern "C" int avssl_moco_infonce_fwd_bwd_enqueue_indexed(const float* feat_q, const float* key_rows, int n_key_rows,
                                                          int keys_raw, const int64_t* row_idx,
                                                          const int64_t* enq_row_idx, int n_enq, float* queue,
                                                          int64_t* ptr_dev, uint32_t* status_dev, int B, int D, int K,
                                                          float T, float* q_out, float* loss_out, float* dfeat_out,
                                                          float* row_lse_out, float* logits_out, void* workspace,
                                                          size_t workspace_bytes, int impl, void* stream) {
  AVSSL_REQUIRE(key_rows && n_key_rows > 0, AVSSL_ERR_INVALID_ARGUMENT, "moco_infonce_indexed: key_rows is null or empty");
  AVSSL_REQUIRE(row_idx || n_key_rows >= B, AVSSL_ERR_INVALID_ARGUMENT,
                "moco_infonce_indexed: %d key rows for B=%d without a row index", n_key_rows, B);
  const float* keys_host[1] = {key_rows};
  return moco_infonce_run(feat_q, keys_host, 1, queue, ptr_dev ? queue : nullptr, ptr_dev, status_dev, B, D, K, T, q_out,
                          loss_out, dfeat_out, row_lse_out, logits_out, workspace, workspace_bytes, impl, stream, nullptr,
                          row_idx, enq_row_idx, n_enq, n_key_rows, keys_raw);
}
