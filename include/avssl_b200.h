/*
 * avssl_b200.h — C-ABI of the B200-native contrastive hot path.
 *
 * Drop-in boundary for the contrastive head of JingwWu/advise-video-ssl
 * (`models/contrastive.py`, `models/losses.py:15-25`, `utils/distributed.py:79-155`).
 * The reference is pure Python/torch and has no FFI of its own; these are the
 * entry points SURVEY.md §8(b) lists for a C-ABI replacement.  Each function
 * cites the reference lines whose op sequence it replaces.
 *
 * Conventions
 *   - plain C: raw pointers, sizes, `cudaStream_t` passed as `void*`; no torch types.
 *   - device pointers unless the name ends in `_host`.
 *   - every call is asynchronous on `stream` and never synchronises unless stated.
 *   - return value: 0 = AVSSL_OK, otherwise an avssl_status code; the message for
 *     the calling thread's last failure is available from avssl_last_error().
 *   - the caller owns all memory; the library keeps no global device state.  The one exception
 *     is the peer-exchange buffer of avssl_peer_alloc(): it must be a whole cudaMalloc allocation
 *     to be exportable over CUDA IPC, so the library allocates it and the caller frees it with
 *     avssl_peer_free().
 *   - there is NO CPU fallback: without a CUDA device every compute call fails.
 *   - fp32 tensors are row-major and contiguous; indices / pointers are int64.
 */
#ifndef AVSSL_B200_H_
#define AVSSL_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define AVSSL_API __attribute__((visibility("default")))
#else
#define AVSSL_API
#endif

typedef enum {
  AVSSL_OK = 0,
  AVSSL_ERR_INVALID_ARGUMENT = 1,
  AVSSL_ERR_CUDA = 2,
  AVSSL_ERR_UNSUPPORTED = 3,
  AVSSL_ERR_WORKSPACE = 4
} avssl_status;

/* Bits of the device-side status word written by kernels that replace host asserts. */
#define AVSSL_DEVFLAG_QUEUE_OVERRUN 1u /* ptr + n > K   (models/contrastive.py:285) */
#define AVSSL_DEVFLAG_BAD_INDEX 2u     /* bank index out of range                   */
#define AVSSL_DEVFLAG_PEER_TIMEOUT 4u  /* a peer's push did not arrive within avssl_peer_xchg.timeout_ms */

AVSSL_API int avssl_abi_version(void);
AVSSL_API const char* avssl_last_error(void);
/* Number of SMs of the current device (grid sizing), or <0 on error. */
AVSSL_API int avssl_device_sm_count(void);

/* ------------------------------------------------------------------ K1: momentum EMA
 * Replaces ContrastiveModel._update_history (models/contrastive.py:158-172):
 *     hist <- online * (1 - m) + hist * m            for every parameter tensor,
 * three separately rounded fp32 ops (mul, mul, add — no FMA contraction) so the
 * result is bit-identical to the reference's ATen sequence.  When *iter_dev == 0
 * the history is first replaced by the online weights (:167-169) and then blended.
 *
 * The pointer table is a device array of avssl_ema_chunk built once on the host by
 * avssl_ema_plan_fill() and reused every step.
 */
typedef struct {
  const float* online; /* start of this chunk in the online (query) encoder tensor */
  float* hist;         /* same offset in the momentum (key) encoder tensor         */
  uint32_t n;          /* elements in this chunk (<= chunk_elems)                  */
  uint32_t flags;      /* bit0: both pointers 16-byte aligned                      */
} avssl_ema_chunk;

/* Elements per chunk used by the kernel's fast path. */
AVSSL_API int64_t avssl_ema_chunk_elems(void);
/* Number of chunks for a parameter list (host; no CUDA). */
AVSSL_API int64_t avssl_ema_plan_chunks(const int64_t* numel_host, int n_tensors);
/* Fill a host table of avssl_ema_plan_chunks() entries (host; no CUDA). */
AVSSL_API int avssl_ema_plan_fill(const uint64_t* online_ptrs_host, const uint64_t* hist_ptrs_host,
                        const int64_t* numel_host, int n_tensors,
                        avssl_ema_chunk* table_host, int64_t n_chunks);
/*
 * m, one_minus_m: the fp32 roundings of the host doubles m and (1.0 - m), exactly
 *   what ATen uses for `tensor * python_float`.
 * iter_dev: int64[1] device-resident step counter (buffer `iter`, :90); never copied
 *   to the host on the step path (the reference's int(self.iter) D2H sync, :161, is gone).
 * first_iter: 1 / 0 when the caller knows whether iter == 0 (it mirrors the counter on
 *   the host), -1 to have the kernel read *iter_dev itself.
 * bump_iter != 0: also performs `self.iter += 1` (compute_key_feat, :314).  With
 *   first_iter = -1 the increment is made by the last block to finish and needs
 *   done_counter (uint32[1], zero-initialised, self-resetting); otherwise it is free.
 */
AVSSL_API int avssl_ema_multi_tensor(const avssl_ema_chunk* table_dev, int64_t n_chunks, float m,
                           float one_minus_m, int64_t* iter_dev, int first_iter, int bump_iter,
                           uint32_t* done_counter_dev, void* stream);

/* -------------------------------------------- K2+K3: l2-norm + MoCo logits + InfoNCE
 * Replaces the MoCo score/loss block (models/contrastive.py:462, 486-500) and
 * ContrastiveLoss.forward (models/losses.py:20-25), forward AND backward in one
 * pass over the queue:
 *     q = f / ||f||;  s_ij = q_i . queue_j / T;  s0_ki = q_i . key_ki / T
 *     loss = mean_{k,i} ( log( e^{s0_ki} + sum_j e^{s_ij} ) - s0_ki )
 *     dq_i = sum_k ( sum_j p_kij queue_j + p_ki0 key_ki - key_ki ) / (T n_rows)
 *     df_i = ( dq_i - (dq_i . q_i) q_i ) / ||f_i||
 * keys_host: host array of n_keys (<= AVSSL_MAX_KEYS) device pointers, each [B, D],
 *   already normalised and detached (compute_key_feat output).
 * logits_out: optional [n_keys*B, K+1] (the tensor the reference returns, :498,506);
 *   NULL skips the 4*n_keys*B*(K+1) bytes of writes.
 * row_lse_out: optional [n_keys*B] log-sum-exp per logits row.
 * impl: AVSSL_IMPL_AUTO picks the tcgen05 kernel when the shape allows it.
 * workspace: avssl_moco_infonce_workspace_bytes() bytes, 256-byte aligned, zero-filled
 *   ONCE by the caller when it is allocated; the call leaves it reusable.
 */
#define AVSSL_MAX_KEYS 8
#define AVSSL_IMPL_AUTO 0
#define AVSSL_IMPL_SIMT 1   /* fp32 CUDA-core kernel, any D % 4 == 0, D <= 256 */
#define AVSSL_IMPL_TC3X 2   /* tcgen05 kind::tf32, 3-term error-compensated (fp32-grade) */
#define AVSSL_IMPL_TC1X 3   /* tcgen05 kind::tf32, single pass (tf32-grade)    */

AVSSL_API size_t avssl_moco_infonce_workspace_bytes(int B, int D, int K, int n_keys);
AVSSL_API int avssl_moco_infonce_fwd_bwd(const float* feat_q, const float* const* keys_host, int n_keys,
                               const float* queue, int B, int D, int K, float T, float* q_out,
                               float* loss_out, float* dfeat_out, float* row_lse_out,
                               float* logits_out, void* workspace, size_t workspace_bytes,
                               int impl, void* stream);
/* Same, followed by the queue ring write of K4 for keys_host[0] (the non-multi-view
 * `_dequeue_and_enqueue(keys)` of models/contrastive.py:502-503, 263-292):
 *     queue[ptr:ptr+B] = keys[0];  ptr = (ptr + B == K) ? 0 : ptr + B
 * The logits, loss and gradient are computed against the queue BEFORE the write, as in
 * the reference.  With the tcgen05 kernels the write happens inside the same launch, after
 * a grid-wide barrier behind the last read of the queue; ptr_dev / status_dev as in
 * avssl_queue_enqueue (requires K % B == 0; ptr + B <= K is checked on the device).
 */
AVSSL_API int avssl_moco_infonce_fwd_bwd_enqueue(const float* feat_q, const float* const* keys_host, int n_keys,
                                       float* queue, int64_t* ptr_dev, uint32_t* status_dev, int B, int D,
                                       int K, float T, float* q_out, float* loss_out, float* dfeat_out,
                                       float* row_lse_out, float* logits_out, void* workspace,
                                       size_t workspace_bytes, int impl, void* stream);

/* --------------------------------------------------------------- K4: queue ring write
 * Replaces _dequeue_and_enqueue (models/contrastive.py:263-292) for one key tensor:
 *     queue[ptr:ptr+n] = keys;  ptr = (ptr + n == K) ? 0 : ptr + n
 * with ptr device-resident (no .item() sync, :265).  The host assert K % n == 0
 * (:284) is checked by the call; `ptr + n <= K` (:285) is checked on the device:
 * on violation nothing is written and AVSSL_DEVFLAG_QUEUE_OVERRUN is or-ed into
 * *status_dev (uint32[1], may be NULL).
 */
AVSSL_API int avssl_queue_enqueue(float* queue, int64_t* ptr_dev, const float* keys, int n, int K, int D,
                        uint32_t* status_dev, void* stream);

/* ----------------------------------------------------- K5: memory-bank update (scatter)
 * Replaces Memory.update (models/contrastive.py:989-1036), Memory1D.update (:1066-1080)
 * and knn_mem_update (:131-140), after the all_gather of (mem, ind, time):
 *   interp == 0:  bank[ind, time] <- l2norm( mem * m + bank[ind, time] * (1 - m) )
 *   interp != 0:  the two-row time-interpolated update of :995-1026 (float times).
 * bank is [L, duration, D] (Memory1D: duration = 1).  Indices are applied bit-exactly;
 * duplicate targets resolve like the reference's CPU index_put (all updates computed
 * from the old rows, last occurrence wins).  Out-of-range indices set
 * AVSSL_DEVFLAG_BAD_INDEX in *status_dev and are skipped.  time_i64 may be NULL (= 0).
 */
AVSSL_API int avssl_membank_update(float* bank, int64_t L, int duration, int D, const float* mem,
                         const int64_t* ind, const int64_t* time_i64, const float* time_f32, int n,
                         float momentum, float one_minus_momentum, int interp, uint32_t* status_dev,
                         void* stream);

/* ------------------------------------------------- K14: mem-mode logits (fused gather-dot)
 * Replaces Memory.get + einsum + div (models/contrastive.py:429-433, :966-987):
 *   prod[n, k] = q_n . bank[ind[n,k], time[n,k]] / T        (never builds [B, K+1, D])
 */
AVSSL_API int avssl_membank_gather_dot(const float* bank, int64_t L, int duration, int D, const float* q,
                             const int64_t* ind, const int64_t* time_i64, const float* time_f32, int B,
                             int Kp, float T, int interp, float* prod, uint32_t* status_dev, void* stream);

/* --------------------------------------------------------- K2: row l2-normalisation
 * y = x / max(||x||, eps) per row of x [n, D].  eps = 0 is Normalize
 * (models/contrastive.py:923-934); eps = 1e-12 is F.normalize (:617-621, :850, :867).
 * norm_out [n] receives ||x|| (needed by the backward).
 */
AVSSL_API int avssl_l2norm_fwd(const float* x, int n, int D, float eps, float* y, float* norm_out, void* stream);
AVSSL_API int avssl_l2norm_bwd(const float* y, const float* norm, const float* dy, int n, int D, float eps,
                     float* dx, void* stream);

/* ------------------------------------------------------- K7: BYOL similarity loss
 * Replaces sim_loss (models/contrastive.py:243-249) fused with the predictor
 * l2-norm (:533) when normalize != 0:  loss = -mean_n( p_n . key_n ) / T, and
 * dpred_out (optional) = d loss / d pred.  workspace: zero-filled once, reusable.
 */
AVSSL_API size_t avssl_byol_simloss_workspace_bytes(int n);
AVSSL_API int avssl_byol_simloss_fwd_bwd(const float* pred, const float* key, int n, int D, float T, int normalize,
                               float* loss_out, float* dpred_out, void* workspace, size_t workspace_bytes,
                               void* stream);

/* ------------------------------------------ ContrastiveLoss on materialised logits
 * Replaces ContrastiveLoss.forward (models/losses.py:15-25) for callers that hold a
 * logits tensor (mem mode, models/contrastive.py:436): mean cross-entropy against
 * class 0.  bwd: dlogits = (softmax - onehot0) * (*grad_out_dev) / n.
 */
AVSSL_API size_t avssl_ce_target0_workspace_bytes(int n);
AVSSL_API int avssl_ce_target0_fwd(const float* logits, int n, int C, float* loss_out, float* row_lse_out,
                         void* workspace, size_t workspace_bytes, void* stream);
AVSSL_API int avssl_ce_target0_bwd(const float* logits, const float* row_lse, int n, int C,
                         const float* grad_out_dev, float* dlogits, void* stream);

/* ------------------------------------------------------------- K6: SimCLR NT-Xent
 * Replaces the live SimCLR branch (models/contrastive.py:770-792) and the gradient
 * bookkeeping of AllGatherWithGradient (utils/distributed.py:131-155), computing
 * only the rows this rank owns.  out = [q_all ; q2_all] is [N2, D] with unit rows
 * (already gathered); rows[n_loc] are the global row ids of this rank.
 *   avssl_ntxent_rowsum : z_loc[i] = sum_{c != r_i} exp((out_r . out_c - 1) / T)
 *   avssl_ntxent_grad   : given z for ALL rows, loss = mean_r(log z_r + 1/T - s_{r,r+})
 *       and dfeat[i] = grad_scale * dLoss/dout_{r_i} chained through the row
 *       l2-normalisation (norm_loc[i] = ||f_i||).  grad_scale = world size reproduces
 *       the reference's all_reduce(SUM)-then-slice backward.
 * impl: AVSSL_IMPL_AUTO picks the tcgen05 kernels (kind::f16 on fp16 copies of the unit rows -- on [-1, 1]
 *   the 11 significant bits of round-to-nearest tf32 -- with fp32 accumulation: loss ~1e-5, gradient ~3e-4
 *   relative, inside the 1e-3 fp32 tolerance) when D is 64/128/256 and out_f16 is given; AVSSL_IMPL_SIMT
 *   forces the exact-fp32 CUDA-core kernels (out_f16 may then be NULL), which also serve every other D.
 * row0_first, row1_first: the tcgen05 kernels take this rank's rows as two blocks of n_loc/2 consecutive global rows
 *   (its q rows and its q2 rows): rows[i] = row0_first + i for i < n_loc/2, row1_first + (i - n_loc/2) after.  Pass -1
 *   when `rows` has another structure (CUDA-core kernels then).
 * out_f16: `out` as IEEE fp16 ([N2, D], 2 bytes per element), the operand the tensor cores stream (written
 *   together with `out` by avssl_ntxent_prepare, so the conversion costs no extra pass).
 * workspace: avssl_ntxent_workspace_bytes(), zero-filled once, reusable.
 *
 * avssl_ntxent_prepare assembles `out` from the all_gather result (C4, models/contrastive.py:771-775):
 *   gathered is [world][2][B][D] (every rank's [q ; q2] block, what ncclAllGather delivers),
 *   out[(v*world + w)*B + b] = gathered[w][v][b], out_f16 = rn_fp16(out).  world = 1 just copies.
 */
AVSSL_API size_t avssl_ntxent_workspace_bytes(int N2, int D, int n_loc);
AVSSL_API int avssl_ntxent_prepare(const float* gathered, int world, int B, int D, float* out, void* out_f16,
                         void* stream);
/* Same, reading the gathered rows out of this rank's NVLink exchange buffer (every rank pushed its [2B, D] block of
 * unit rows with avssl_peer_push_rows: C4 without a collective kernel); the launch waits for every rank's flag first.
 * Declared after avssl_peer_xchg below. */
AVSSL_API int avssl_ntxent_rowsum(const float* out, const void* out_f16, const int* rows, int row0_first,
                        int row1_first, int N2, int D, int n_loc,
                        float T, float* z_loc_out, void* workspace, size_t workspace_bytes, int impl,
                        void* stream);
AVSSL_API int avssl_ntxent_grad(const float* out, const void* out_f16, const int* rows, int row0_first,
                      int row1_first, const float* z_all,
                      const float* norm_loc,
                      int N2, int D, int n_loc, float T, float grad_scale, float* loss_out,
                      float* dfeat_out, void* workspace, size_t workspace_bytes, int impl, void* stream);

/* ---------------------------------------------------------- K10: Sinkhorn-Knopp
 * Replaces `q = exp(out / eps).t(); sinkhorn(q.t(), iters)[-keep_last:]`
 * (models/contrastive.py:665-671, 872-887) with ONE cooperative launch.
 * scores [Btot, P] are the raw prototype scores; codes_out is [keep_last, P], every
 * row summing to 1.  workspace: avssl_sinkhorn_workspace_bytes(), zero-filled once.
 */
AVSSL_API size_t avssl_sinkhorn_workspace_bytes(int Btot, int P);
AVSSL_API int avssl_sinkhorn(const float* scores, int Btot, int P, float eps, int iters, int keep_last,
                   float* codes_out, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------ K11: SwAV swapped-prediction loss
 * Replaces models/contrastive.py:672-679 (and KLDivLoss, :912-916), forward and
 * backward: loss = sum_{a,v} w[a][v] * sum_r ( - sum_k code[a][r][k] *
 * log softmax(scores[v*bs + r] / T)[k] ).  pair_w_host is the host array
 * w[n_assign][n_crops] (the reference's 1/(bs (n_crops-1) n_assign) for v != crop(a)).
 * dscores_out (optional) = d loss / d scores.
 */
AVSSL_API size_t avssl_swav_ce_workspace_bytes(int n_rows);
AVSSL_API int avssl_swav_ce_fwd_bwd(const float* scores, const float* codes, int n_crops, int n_assign, int bs,
                          int P, float T, const float* pair_w_host, float* loss_out, float* dscores_out,
                          void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------- C3: cross-GPU key exchange over NVLink peer memory
 * Replaces the key all_gather + row select of _batch_unshuffle (models/contrastive.py:216-230;
 * cat_all_gather, utils/distributed.py:5-13) for the ranks of ONE box, without a collective
 * kernel: every rank stores its [rows_per_rank, D] block straight into every peer's exchange
 * buffer over NVLink and publishes a per-source epoch flag; consumers spin on their local flags.
 * Bit-exact (plain copies).  Two payload slots (epoch parity) make back-to-back steps safe.
 *
 * Set-up (once): each rank calls avssl_peer_alloc() (cudaMalloc + zero-fill + CUDA IPC export),
 * the caller moves the AVSSL_IPC_HANDLE_BYTES-byte handles between the processes (e.g. a
 * torch.distributed all_gather), every rank opens the others with avssl_peer_open() and fills an
 * avssl_peer_xchg with base[r] = rank r's buffer as mapped HERE (base[rank] = its own).
 * Every rank must issue the same sequence of pushes (one per step).
 */
#define AVSSL_MAX_PEERS 16
#define AVSSL_IPC_HANDLE_BYTES 64
typedef struct {
  void* base[AVSSL_MAX_PEERS]; /* exchange buffers of all ranks, in this process's address space */
  int world, rank;
  int rows_per_rank, D;        /* every rank pushes [rows_per_rank, D] fp32 per step */
  uint32_t timeout_ms;         /* bound of the consumer's spin on a peer's flag; 0 = wait forever.  On expiry
                                  AVSSL_DEVFLAG_PEER_TIMEOUT is or-ed into the status word and the kernel
                                  carries on (that step's keys are garbage; the host must stop). */
  uint32_t reserved_;
} avssl_peer_xchg;

AVSSL_API size_t avssl_peer_xchg_bytes(int world, int rows_per_rank, int D);
AVSSL_API int avssl_peer_alloc(size_t bytes, void** dev_ptr_out, void* ipc_handle_out_host);
AVSSL_API int avssl_peer_open(const void* ipc_handle_host, void** dev_ptr_out);
AVSSL_API int avssl_peer_close(void* dev_ptr);
AVSSL_API int avssl_peer_free(void* dev_ptr);
/* push (stand-alone launch): one CTA per destination rank for blocks up to 64 KiB, several per destination beyond. */
AVSSL_API int avssl_peer_push_rows(const avssl_peer_xchg* x, const float* rows, void* stream);
/* push with Normalize fused in (models/contrastive.py:923-934 applied to the key features, :350):
 * what travels is feat / max(||feat||, eps), bit-identical to avssl_l2norm_fwd followed by
 * avssl_peer_push_rows.  y_local_out (optional, [rows_per_rank, D]) receives this rank's rows. */
AVSSL_API int avssl_l2norm_push_rows(const avssl_peer_xchg* x, const float* feat, float eps, float* y_local_out,
                           void* stream);
/* wait for the current epoch from every rank, then out[i] = gathered[row_idx[i]] (row_idx NULL:
 * this rank's own block, n_out <= rows_per_rank).  gathered is [world * rows_per_rank, D] in rank
 * order = what cat_all_gather returns.  Out-of-range indices set AVSSL_DEVFLAG_BAD_INDEX. */
AVSSL_API int avssl_peer_wait_gather(const avssl_peer_xchg* x, const int64_t* row_idx, int n_out, float* out,
                           uint32_t* status_dev, void* stream);
/* C1 -- the shuffle-BN exchange of the key encoder's INPUT rows (models/contrastive.py:174-214: cat_all_gather of
 * the clip, then x[perm.view(W,-1)[rank]]) as a scatter over NVLink peer stores, ONE launch: every row is written by
 * its owner straight into its final position in the destination rank's buffer, once; the same kernel then waits for
 * every rank's rows of this epoch and copies this rank's rows_per_rank received rows to `out`.  Uses an
 * avssl_peer_xchg descriptor whose buffers were allocated with avssl_peer_scatter_bytes(rows_per_rank, row_bytes)
 * and D = row_bytes / 4 (rows are opaque bytes; row_bytes % 16 == 0).  dest_pos_dev[j] = argsort(perm)[rank *
 * rows_per_rank + j] (int64, device): the position of local row j in the rank-major shuffled batch.
 * dest_pos_host (optional, same values on the host): with at most 256 rows per rank the positions travel in the
 * kernel parameters and the launch depends on no host-to-device copy (dest_pos_dev may then be NULL).  Bit-exact;
 * every rank of the exchange must make the call (same sequence on every rank). */
AVSSL_API size_t avssl_peer_scatter_bytes(int rows_per_rank, int64_t row_bytes);
AVSSL_API int avssl_peer_scatter_exchange(const avssl_peer_xchg* x, const void* rows, const int64_t* dest_pos_dev,
                                const int64_t* dest_pos_host, void* out, uint32_t* status_dev, void* stream);
/* K1 with the push fused into the same launch: `world` extra CTAs at the front of the EMA grid
 * push `rows` while the rest stream the parameters (north_star: the key exchange overlapped with
 * the EMA kernel).  Other arguments as avssl_ema_multi_tensor. */
AVSSL_API int avssl_ema_multi_tensor_push(const avssl_ema_chunk* table_dev, int64_t n_chunks, float m,
                                float one_minus_m, int64_t* iter_dev, int first_iter, int bump_iter,
                                uint32_t* done_counter_dev, const avssl_peer_xchg* x, const float* rows,
                                void* stream);
/* K2+K3+K4 with the wait fused: like avssl_moco_infonce_fwd_bwd_enqueue with n_keys = 1, but the
 * key rows are taken from the exchange buffer (row i = gathered[row_idx ? row_idx[i] : rank*B + i])
 * after the merge CTAs have waited for the current epoch -- the sweep over the queue never waits.
 * tcgen05 kernels only (AVSSL_ERR_UNSUPPORTED otherwise: use avssl_peer_wait_gather first).
 * ptr_dev may be NULL (no enqueue).
 * Queue consistency across ranks (C9: the reference enqueues local keys and lets DDP's buffer
 * broadcast, models/build.py:76-83, overwrite every queue with rank 0's): the rows written to the
 * queue are gathered[enq_row_idx[e]], e < n_enq, and ptr advances by n_enq (K % n_enq == 0).
 * enq_row_idx NULL: this rank's own block (n_enq must be B).  With enq_row_idx = rank 0's rows on
 * every rank the queues stay bit-identical without any broadcast; with all world*B rows it is
 * canonical MoCo.
 * push_feat (optional, [rows_per_rank, D]): this rank's RAW key-encoder output.  One extra CTA of the launch then
 * performs the push of this step itself -- Normalize (models/contrastive.py:350) and the stores into every rank's
 * buffer -- while the others sweep the queue, so no launch sits between the key encoder and the head; without it the
 * caller must have pushed (avssl_peer_push_rows / avssl_l2norm_push_rows / the EMA-fused push) before this call. */
AVSSL_API int avssl_moco_infonce_fwd_bwd_enqueue_peer(const float* feat_q, const avssl_peer_xchg* x,
                                            const int64_t* row_idx, const int64_t* enq_row_idx, int n_enq,
                                            const float* push_feat, float* queue, int64_t* ptr_dev,
                                            uint32_t* status_dev, int B, int D, int K, float T, float* q_out,
                                            float* loss_out, float* dfeat_out, float* row_lse_out,
                                            float* logits_out, void* workspace, size_t workspace_bytes,
                                            int impl, void* stream);

AVSSL_API int avssl_ntxent_prepare_peer(const avssl_peer_xchg* x, uint32_t* status_dev, int B, int D, float* out,
                              void* out_f16, void* stream);

/* K2+K3+K4 with the un-shuffle (and optionally the key Normalize) folded in, for keys that live on THIS
 * device (one GPU, or after an NCCL gather): key_rows is [n_key_rows, D]; query row i meets
 * key_rows[row_idx ? row_idx[i] : i] (idx_restore[rank] of models/contrastive.py:209-230), the queue receives
 * key_rows[enq_row_idx ? enq_row_idx[e] : e], e < n_enq.  keys_raw != 0: key_rows holds the key encoder's raw
 * output and x / ||x|| (Normalize, :350, :923-934) is applied where the rows are read, bit-identical to
 * avssl_l2norm_fwd.  tcgen05 kernels only.  ptr_dev may be NULL (no enqueue). */
AVSSL_API int avssl_moco_infonce_fwd_bwd_enqueue_indexed(const float* feat_q, const float* key_rows, int n_key_rows,
                                               int keys_raw, const int64_t* row_idx, const int64_t* enq_row_idx,
                                               int n_enq, float* queue, int64_t* ptr_dev, uint32_t* status_dev,
                                               int B, int D, int K, float T, float* q_out, float* loss_out,
                                               float* dfeat_out, float* row_lse_out, float* logits_out,
                                               void* workspace, size_t workspace_bytes, int impl, void* stream);

/* K2+K3 in two launches.  The sweep of q against the queue -- 99.9% of the head's work, and all of logits[:, 1:] --
 * depends only on the query encoder's output and the queue, not on the keys: in the reference's own order
 * (models/contrastive.py:462 query encoder, then :478 compute_key_feat = :314 momentum update, :338 shuffle, key
 * encoder, :356 un-shuffle) it can start before the key path does.  avssl_moco_infonce_sweep() launches it alone (any
 * stream); any avssl_moco_infonce_fwd_bwd* entry called afterwards with AVSSL_HEAD_SWEPT(sweep_ctas) or-ed into `impl`
 * -- same B, D, K, n_keys, T, logits_out, workspace and sweep_ctas, on a stream ordered behind the sweep -- then only
 * merges the partials with the key term (loss, gradient, logits[:, 0], enqueue, fused exchange push / wait).
 * sweep_ctas: 0 = one CTA per SM; a smaller count leaves SMs (the kernel owns all registers of the SMs it runs on) to
 * a bandwidth-bound kernel running beside it, e.g. the momentum update.  tcgen05 kernels only
 * (AVSSL_ERR_UNSUPPORTED otherwise).  The sweep call is two launches: the sweep, then the per-CTA partials of every
 * query row reduced to one, so that the call behind the key path reads one partial per row.  logits and q_out are
 * bit-identical to the single-launch form; loss, lse and gradient differ by the (fixed, deterministic) order in
 * which the partials are merged. */
#define AVSSL_HEAD_SWEPT(sweep_ctas) (0x100 | ((int)(sweep_ctas) << 16))
AVSSL_API int avssl_moco_infonce_sweep(const float* feat_q, const float* queue, int B, int D, int K, float T, int n_keys,
                             float* logits_out, void* workspace, size_t workspace_bytes, int impl, int sweep_ctas,
                             void* stream);

/* ------------------------------------------- multi-tensor L2 norm
 * Replaces get_grad_norm_ (models/optimizer.py:375-397; one torch.norm per parameter, a stack and a
 * .cpu() sync every step from utils/solver.py:109-111) and the torch.norm pairs of LARS.step
 * (models/optimizer.py:351-352):
 *     per_tensor[t] = ||x_t||_2 ,   total = || (per_tensor) ||_2 = sqrt(sum_t ||x_t||^2)
 * table_dev: the chunk table of avssl_ema_plan_fill() built with the tensors as `online`
 *   (`hist` is not read; pass the same pointers); first_chunk_dev: int32[n_tensors + 1], the first
 *   chunk of every tensor (prefix sum of ceil(numel / avssl_ema_chunk_elems())).
 * per_tensor_norm_out may be NULL.  An empty list gives total = 0 (:380-381).
 * workspace: avssl_multi_l2norm_workspace_bytes(), zero-filled once, reusable.
 */
AVSSL_API size_t avssl_multi_l2norm_workspace_bytes(int64_t n_chunks, int n_tensors);
AVSSL_API int avssl_multi_l2norm(const avssl_ema_chunk* table_dev, int64_t n_chunks, const int32_t* first_chunk_dev,
                       int n_tensors, float* per_tensor_norm_out, float* total_norm_out, void* workspace,
                       size_t workspace_bytes, void* stream);

/* ------------------------------------------- projection tail
 * The last Linear of the projection MLP (models/head_helper.py:52-58, `MLPHead.projection[-1]`: y = x W^T + b,
 * W [Dout, Kin] as nn.Linear stores it) with the head's Normalize (models/contrastive.py:923-934, applied to the
 * backbone output at :462 / :350 / :757) as its epilogue:
 *     q = y / max(||y||, eps)        norm_out[b] = ||y_b||            (eps = 0 reproduces Normalize)
 * y itself is never written.  normalize = 0 gives the plain Linear (q = y).  Exact fp32, deterministic.
 * Backward, one launch: grad_q = dL/dq ->  dy = (grad_q - (grad_q . q) q) / ||y||   (the arithmetic of avssl_l2norm_bwd),
 *     dx = dy W      dW = dy^T x      db = sum_b dy          (any of the three outputs may be NULL)
 * Limits: Dout <= avssl_linear_l2norm_max_dout() (256), Kin % 4 == 0, 16-byte aligned tensors; AVSSL_ERR_UNSUPPORTED
 * otherwise (the caller keeps nn.Linear + Normalize for such a layer).
 */
AVSSL_API int avssl_linear_l2norm_max_dout(void);
AVSSL_API int avssl_linear_l2norm_fwd(const float* x, const float* W, const float* bias, int B, int Kin, int Dout, float eps,
                            int normalize, float* q_out, float* norm_out, void* stream);
AVSSL_API int avssl_linear_l2norm_bwd(const float* x, const float* W, const float* q, const float* norm, const float* grad_q,
                            int B, int Kin, int Dout, float eps, int normalize, float* dx_out, float* dW_out,
                            float* db_out, void* stream);

/* ------------------------------------------- kNN evaluation top-k
 * eval_knn (models/contrastive.py:232-241): dist = q bank^T, then dist.topk(knn_k, dim=1, largest=True, sorted=True).
 * The similarities come from the head's tcgen05 mainloop -- avssl_moco_infonce_sweep(q, bank, ..., T = 1) writes them
 * as logits[:, 1:], computed on q / ||q|| -- and avssl_topk_rows() is the top-k behind it: exact, every row of `dist`
 * fetched from HBM once, sorted descending, ties towards the smaller index, deterministic.
 *   dist: [N, ld] fp32, the M candidates of a row start at dist + row * ld (pass logits + 1, ld = M + 1);
 *   q_scale_rows: NULL, or the [N, D] queries the similarities were normalised by: yd is multiplied by ||q_row||;
 *   yd_out [N, k] fp32, yi_out [N, k] int64 (the dtype torch.topk returns).
 * k <= min(M, 1024), otherwise AVSSL_ERR_UNSUPPORTED (avssl_topk_rows_workspace_bytes() returns 0 for such a shape);
 * M is not limited (the rows are streamed).  workspace: no initialisation needed.
 */
AVSSL_API size_t avssl_topk_rows_workspace_bytes(int N, int M, int k);
AVSSL_API int avssl_topk_rows(const float* dist, int64_t ld, int N, int M, int k, const float* q_scale_rows, int D,
                    float* yd_out, int64_t* yi_out, void* workspace, size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* AVSSL_B200_H_ */
