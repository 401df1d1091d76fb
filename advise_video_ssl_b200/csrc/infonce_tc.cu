// K2+K3 on the 5th-generation tensor cores: TMA -> shared memory -> tcgen05.mma
// (kind::tf32) with both accumulators in TMEM, online softmax in the epilogue warps.
//
// Replaces models/contrastive.py:462,486-500 + models/losses.py:20-25 (see infonce.cuh).
// One pass over the queue.  Per 64-row queue tile:
//     S   = q . tile^T            (M=128 x N=64,  K=D)      A = q   in TMEM, B = tile, K-major
//     P   = 2^(S*log2e/T - m)     (softmax warps: tcgen05.ld -> exp2 -> tcgen05.st, P overwrites S)
//     acc += P . tile             (M=128 x N=D,   K=64)     A = P   in TMEM, B = tile, MN-major
// so loss statistics (m, l) AND the gradient term sum_j p_ij queue_j come out of the
// same sweep.  tf32 operands that are MN-major (the second GEMM reads the tile with the
// queue-row index as K) must use the 128B-swizzle-with-32B-atoms shared-memory layout,
// while the K-major operand of the first GEMM needs the plain 128B swizzle, so every tile
// exists twice in shared memory: the "S tile" ring feeds the first GEMM, the "V tile" ring
// the second.  An S slot is recycled as soon as its S GEMM retires, a V slot when its PV
// GEMM retires.
//
// fp32-grade accuracy on tf32 tensor cores (kThreeTerm): the tensor core reads an fp32
// operand as tf32 by ignoring its low 13 mantissa bits, so with k_hi = trunc(k) (what the
// hardware sees in the RAW tile), k_lo = k - k_hi (exact), q_hi = rn_tf32(q), q_lo = q - q_hi:
//     S = q_hi.k_hi  (kind::tf32, raw tile)  +  [q_lo | q] . [k | k_lo]   (kind::f16, bf16)
// The two correction terms are ~2^-11 of S, so bf16 operands (2^-9) keep them to ~2^-20, and
// they ride in ONE bf16 MMA pass with the contraction length doubled (A = [q_lo | q] packed
// in TMEM, B = the "correction tile" [bf16(k) | bf16(k_lo)] that four helper warps write
// into the second half of the S slot).  Shared-memory bandwidth (128 B/clk/SM, shared by
// TMA writes, the helper warps and the tensor core's operand reads) bounds the sweep, so
// every byte is touched as few times as possible: the tile is landed ONCE by TMA, the
// helper warps read it once, write the correction tile and -- from registers, once PV(t-2)
// has released the slot -- the V tile rounded to nearest tf32 (unbiased, unlike truncation)
// in the 32B-atom layout.  Without the split (kThreeTerm = false) TMA lands both layouts and
// the hardware truncates: tf32-grade logits.
//
// Measured on B200 (tools/microbench): one thread issues a tcgen05.mma every ~55 cycles at
// best and an M=128, N=64, K=8 tf32 MMA occupies the tensor pipe for 32 cycles, so the issue
// loop is kept warp-uniform (descriptors in uniform registers, one elected lane) and fully
// unrolled.
//
// Warp roles (512 threads, 1 CTA / SM):  0 TMA producer | 1 MMA issuer + TMEM allocator |
// 2-9 helper warps (correction + V tile) | 10-13 softmax + epilogue (thread = query row; warp
// w owns TMEM lanes 32*(w%4)..) | 14-15 logits writers (with the softmax warps that own no rows).
// TMEM columns: [0,D) q_hi | [D,2D) [q_lo | q] as packed bf16 | [2D,2D+128) S/P double buffer |
// [2D+128,3D+128) acc.
//
// The launch is cooperative (grid <= number of SMs, all CTAs co-resident): after the sweep
// every CTA writes its partial (m, l, acc), the grid meets at a barrier, and the CTAs then
// merge the partials one query row each (infonce_combine.cuh), take the mean loss and --
// optionally -- perform the queue ring write of K4 (models/contrastive.py:263-292), which
// is safe there because no CTA reads the queue any more.  One launch per step instead of
// three (split, combine, enqueue).
#include "infonce.cuh"
#include "tc_trace.cuh"
#include "infonce_combine.cuh"
#include "sm100_ptx.cuh"

namespace avssl {

namespace {

constexpr int kBlockJ = kTcTileRows;  // 64 queue rows per tile
constexpr int kM = 128;               // query rows per CTA
constexpr int kTcThreads = 512;
constexpr int kGroupThreads = 128;    // softmax warps
constexpr int kHelperThreads = 256;   // helper warps 2..9
constexpr float kRescaleThreshold = 8.f;  // log2 units: P stays below 2^8
constexpr int kMaxSlots = 3;
constexpr int kStageRows = 64;            // query rows per CTA whose logits go through the staging tiles
constexpr int kStagePitch = kBlockJ + 1;  // floats; odd pitch: row- and column-wise accesses conflict-free
constexpr int kDrainWarps = 4;            // warps 10, 11 (softmax warps without rows), 14, 15
constexpr int kLogitsWarps = 4;           // helper warps 4, 5, 8, 9: the ones that may read TMEM lanes 0..63

template <int D, bool kThreeTerm>
struct TcCfg {
  static constexpr int kKB = D / 32;                    // 128-byte k-blocks per row
  static constexpr int kBoxBytes = kBlockJ * 128;       // one TMA box: 64 rows x 128 B
  static constexpr int kTileBytes = kKB * kBoxBytes;    // 32 KiB at D = 128
  static constexpr int kSSlotBytes = kThreeTerm ? 2 * kTileBytes : kTileBytes;  // hi [, lo]
  static constexpr int kSSlots = kThreeTerm ? 2 : 3;
  static constexpr int kVSlots = kThreeTerm ? 2 : 3;
  static constexpr int kSRingBytes = kSSlots * kSSlotBytes;
  static constexpr int kVRingBytes = kVSlots * kTileBytes;
  static constexpr int kStageBytes = kStageRows * kStagePitch * 4;  // one logits staging tile [64][65]
  static constexpr int kScratchBytes = 2 * kStageBytes;            // double-buffered
  static constexpr int kColQhi = 0, kColQlo = D, kColS = 2 * D, kColAcc = 2 * D + 2 * kBlockJ;
  static constexpr int kTmemCols = 512;
  static_assert(3 * D + 2 * kBlockJ <= 512, "TMEM budget");
  static_assert(kSRingBytes >= (int)sizeof(CombineSmem<kTcThreads>), "merge scratch aliases the S ring");
  static constexpr size_t kSmemBytes = 1024 /*align slack*/ + (size_t)kSRingBytes + kVRingBytes + kScratchBytes + 1024;
};

struct TcBarriers {
  uint64_t s_full[kMaxSlots], s_op[kMaxSlots], s_free[kMaxSlots];
  uint64_t v_full[kMaxSlots], v_op[kMaxSlots], v_free[kMaxSlots];
  uint64_t s_ready[2], p_ready[2], pv_done[2];
  uint64_t q_ready, acc_done;
  uint64_t stage_full[2], stage_free[2];
  uint64_t s_read[2];        // staged CTAs: the logits warps have read S(t) out of TMEM, P may overwrite it
  uint32_t tmem_base;
  float lscale[kStageRows];  // staged CTAs: 1 / (T ||f_r||) of the CTA's rows, for the logits warps
};
static_assert(sizeof(TcBarriers) <= 1024, "barrier block");

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// Everything behind the sweep: one query row per CTA -- its partials merged, the key term added, loss terms, lse, q, the
// gradient through the normalisation, logits column 0 -- then the queue ring write of K4 and the mean over the rows by
// the CTA that arrives last.  Called by all threads of every CTA: in the cooperative kernel after the grid barrier,
// or as the body of infonce_finish_kernel (two-launch form).  `cta` of `n_ctas`.
template <int kThreads>
__device__ __forceinline__ void infonce_tail(const InfoNceParams& p, CombineSmem<kThreads>& csm, int cta, unsigned n_ctas) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int D = p.D;
  // ------------------------------------------- merge: one query row per CTA, then the mean
  if (tid == 0) TC_TRACE(14, 3);
  const float* key_base = p.keys[0];
  const long long n_key_rows = p.use_peer ? (long long)p.peer.world * p.peer.rows_per_rank : (long long)p.n_key_rows;
  const long long own_base = p.use_peer ? (long long)p.peer.rank * p.peer.rows_per_rank : 0ll;
  // the first row indices of this CTA and the ring pointer in ONE round trip (everything here is L2-cold in the
  // two-launch form: the momentum update has streamed through L2 since the data was written)
  long long krow_first = own_base + cta, erow_first = own_base + cta;
  if (p.peer_row_idx && cta < p.B) krow_first = __ldg(p.peer_row_idx + cta);
  if (p.enq_row_idx && p.enq_ptr && cta < p.n_enq) erow_first = __ldg(p.enq_row_idx + cta);
  long long enq_ptr = 0;
  bool enq_ok = false;
  if (p.enq_ptr) {
    enq_ptr = *reinterpret_cast<volatile long long*>(p.enq_ptr);  // advanced only after every CTA arrived below
    enq_ok = enq_ptr >= 0 && enq_ptr + p.n_enq <= p.K;           // models/contrastive.py:285
  }
  if (p.use_peer && (cta < p.B || (enq_ok && cta < p.n_enq))) {  // C3: the keys come from the peer exchange buffer (peer.cuh)
    if (warp == 0) {  // long since landed: the sweep took ~15 us
      const int slot = peer_wait_all_warp(p.peer, p.enq_status);
      if (lane == 0) csm.peer_slot = slot;
    }
    __syncthreads();
    key_base = peer_payload(p.peer.base[p.peer.rank], csm.peer_slot, p.peer);
  }
  for (int i = cta; i < p.B; i += (int)n_ctas) {
    long long krow = i == cta ? krow_first : (p.peer_row_idx ? p.peer_row_idx[i] : own_base + i);  // un-shuffle by index (:216-230)
    if (krow < 0 || krow >= n_key_rows) {  // uniform over the CTA
      if (tid == 0 && p.enq_status) atomicOr(p.enq_status, AVSSL_DEVFLAG_BAD_INDEX);
      krow = 0;
    }
    infonce_combine_row<kThreads, 1>(p, i, csm, key_base + (size_t)krow * D);
  }
  if (enq_ok) {
    // K4 (+ C9): queue[ptr + e] = the e-th row of the enqueue list -- keys[0][e], or key rows picked by enq_row_idx
    // (rank 0's block of the gathered buffer on every rank keeps the queues of all ranks identical, as the
    // reference's DDP buffer broadcast does; all world*B rows = canonical MoCo).  No CTA reads the queue after
    // the grid barrier.
    for (int e = cta; e < p.n_enq; e += (int)n_ctas) {
      const long long krow = e == cta ? erow_first : (p.enq_row_idx ? p.enq_row_idx[e] : own_base + e);
      if (krow < 0 || krow >= n_key_rows) {
        if (tid == 0 && p.enq_status) atomicOr(p.enq_status, AVSSL_DEVFLAG_BAD_INDEX);
        continue;
      }
      const float* src_row = key_base + (size_t)krow * D;
      float4* dst = reinterpret_cast<float4*>(p.queue_rw + (size_t)(enq_ptr + e) * D);
      if (p.keys_raw) {  // Normalize on the way in: x / ||x||, the bits l2norm_fwd_kernel writes
        __syncthreads();
        if (warp == 0) {
          const float knrm = warp_row_norm(src_row, D, lane);
          if (lane == 0) csm.bcast[1] = knrm;
        }
        __syncthreads();
        const float knrm = csm.bcast[1];
        for (int c4 = tid; c4 < D / 4; c4 += kThreads) {
          const float4 v = __ldcg(reinterpret_cast<const float4*>(src_row) + c4);
          dst[c4] = make_float4(v.x / knrm, v.y / knrm, v.z / knrm, v.w / knrm);
        }
      } else {
        for (int c4 = tid; c4 < D / 4; c4 += kThreads) dst[c4] = __ldcg(reinterpret_cast<const float4*>(src_row) + c4);
      }
    }
  }
  if (tid == 0) TC_TRACE(14, 4);
  const bool last_cta = infonce_finish<kThreads>(p, n_ctas, csm);
  if (tid == 0) TC_TRACE(14, 5);
  if (last_cta) {
    if (tid == 0) {
      p.counter[1] = 0u;  // every CTA is past the barrier: it incremented counter[0] afterwards
      p.counter[2] = 0u;  // (infonce_finish_kernel: and past its wait for the push CTAs)
      if (p.enq_ptr) {
        if (enq_ok) {
          long long np = enq_ptr + p.n_enq;
          if (np == p.K) np = 0;  // wrap only when landing exactly on K (:290-291)
          *p.enq_ptr = np;
        } else if (p.enq_status) {
          atomicOr(p.enq_status, AVSSL_DEVFLAG_QUEUE_OVERRUN);
        }
      }
    }
  }
}

template <int D, bool kThreeTerm>
__global__ void __launch_bounds__(kTcThreads, 1)
infonce_tc_kernel(const InfoNceParams p, const __grid_constant__ CUtensorMap tmap,
                  const __grid_constant__ CUtensorMap tmap_v) {
  using C = TcCfg<D, kThreeTerm>;
  extern __shared__ uint8_t smem_raw[];
  // swizzled operands need 1024-byte aligned tiles.  The offset is added to the __shared__ array itself
  // (no integer round trip), so the compiler keeps the shared address space and emits LDS/STS instead
  // of generic loads and stores.
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* s_ring = smem;
  uint8_t* v_ring = smem + C::kSRingBytes;
  float* stage = reinterpret_cast<float*>(smem + C::kSRingBytes + C::kVRingBytes);
  TcBarriers* bar = reinterpret_cast<TcBarriers*>(smem + C::kSRingBytes + C::kVRingBytes + C::kScratchBytes);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) TC_TRACE(14, 0);
  const int split = blockIdx.x;
  const int i_base = blockIdx.y * kM;
  const int j_begin = split * p.rows_per_split;
  const int j_end = min(p.K, j_begin + p.rows_per_split);
  const int n_tiles = (j_end - j_begin + kBlockJ - 1) / kBlockJ;
  // The query rows travel global -> registers (coalesced 128-bit loads by the four softmax
  // warps, requested before the set-up barrier) -> shared memory (the V ring + logits staging
  // area, which nobody else touches before q is in TMEM) -> one row per thread.  The 16-byte
  // row padding keeps the per-thread row reads free of bank conflicts.
  constexpr int kQPitch = D * 4 + 16;
  constexpr int kRowF4 = D / 4;                              // float4 per query row
  constexpr int kQHalf = (kM / 2) * kRowF4 / kGroupThreads;  // float4 per thread for 64 rows
  static_assert(kM * kQPitch <= C::kVRingBytes + C::kScratchBytes, "q staging aliases the V ring + logits staging tile");
  uint8_t* q_stage = v_ring;
  const int q_total = min(kM, p.B - i_base) * kRowF4;  // float4 to stage
  const float4* q_gsrc = reinterpret_cast<const float4*>(p.feat_q + (size_t)i_base * D);
  float4 q_buf[kQHalf];
  if (warp >= 10) {
#pragma unroll
    for (int n = 0; n < kQHalf; ++n) {
      const int idx = (tid - 320) + n * kGroupThreads;
      q_buf[n] = idx < q_total ? __ldg(q_gsrc + idx) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }

  // The launch may carry one EXTRA column of CTAs (blockIdx.x == n_splits) that owns no queue split: its first
  // CTA normalises this rank's raw key rows and stores them into every rank's exchange buffer (C3 push, the
  // Normalize of models/contrastive.py:350 fused in) while the others sweep the queue; the merge below waits
  // for every rank's rows after the sweep.  The step then has no launch between the key encoder and the head.
  const bool exch_cta = p.push_feat != nullptr && blockIdx.x == (unsigned)p.n_splits;
  if (exch_cta) {
    __shared__ unsigned long long s_epoch;
    if (blockIdx.y == 0) peer_push_all_cta(p.peer, p.push_feat, p.push_eps, &s_epoch);
  } else {
  // ------------------------------------------------------------------ one-time setup
  if (warp == 0 && lane == 0) {
    ptx::tma_prefetch_desc(&tmap);
    ptx::tma_prefetch_desc(&tmap_v);
    for (int s = 0; s < kMaxSlots; ++s) {
      ptx::mbar_init(&bar->s_full[s], 1);
      ptx::mbar_init(&bar->s_op[s], kHelperThreads);
      ptx::mbar_init(&bar->s_free[s], 1);
      ptx::mbar_init(&bar->v_full[s], 1);
      ptx::mbar_init(&bar->v_op[s], kHelperThreads);
      ptx::mbar_init(&bar->v_free[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      ptx::mbar_init(&bar->s_ready[b], 1);
      ptx::mbar_init(&bar->p_ready[b], kGroupThreads);
      ptx::mbar_init(&bar->pv_done[b], 1);
    }
    ptx::mbar_init(&bar->q_ready, kGroupThreads);
    ptx::mbar_init(&bar->acc_done, 1);
    for (int b = 0; b < 2; ++b) {
      ptx::mbar_init(&bar->stage_full[b], kLogitsWarps * 32);  // the helper warps that read S and stage logits / T
      ptx::mbar_init(&bar->stage_free[b], kDrainWarps * 32);   // the warps that store the staged logits
      ptx::mbar_init(&bar->s_read[b], kLogitsWarps * 32);
    }
    ptx::mbar_fence_init();
  }
  if (warp == 1) ptx::tmem_alloc(&bar->tmem_base, C::kTmemCols);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = bar->tmem_base;
  if (tid == 0) TC_TRACE(14, 1);

  // Logits of a CTA tile with at most 64 query rows (the MoCo case) go through two padded staging
  // tiles.  The two softmax warps that own the rows are the critical path of the sweep (per tile: 64
  // exp2, max, sum and the P store per thread), so they do NOT touch the logits: four of the helper
  // warps -- 4, 5, 8, 9, the ones whose TMEM sub-partition holds lanes 0..63 -- read S(t) themselves
  // (32 columns each), scale it and write the staging tile, and kDrainWarps other warps (the two
  // softmax warps whose rows do not exist + warps 14, 15) store it as whole 128-byte row segments.
  // (Round 1 staged from the softmax threads: 1300 of their 2100 cycles per tile.)  Larger CTA
  // tiles store directly from the softmax threads.
  const bool staged_cta = p.logits_out != nullptr && (p.B - i_base) <= kStageRows;
  auto drain_tile = [&](int t, int dw) {  // dw = 0..kDrainWarps-1; logits / T of tile t (models/contrastive.py:498)
    const int b = t & 1;
    ptx::mbar_wait(&bar->stage_full[b], (t >> 1) & 1);
    if (dw == 0 && lane == 0) TC_TRACE(11, 1 + t);
    const int j0 = j_begin + t * kBlockJ;
    const int valid = min(kBlockJ, j_end - j0);
    const int rows_staged = p.B - i_base;
    const float* st = stage + b * (kStageRows * kStagePitch) + lane;
    const size_t row_pitch = (size_t)(p.K + 1);
    for (int k = 0; k < p.n_keys; ++k) {
      float* dst = p.logits_out + ((size_t)k * p.B + i_base) * row_pitch + 1 + j0 + lane;
#pragma unroll 4
      for (int rr = dw; rr < rows_staged; rr += kDrainWarps) {
        const float v0 = st[rr * kStagePitch], v1 = st[rr * kStagePitch + 32];
        if (lane < valid) dst[(size_t)rr * row_pitch] = v0;
        if (32 + lane < valid) dst[(size_t)rr * row_pitch + 32] = v1;
      }
    }
    if (dw == 0 && lane == 0) TC_TRACE(11, 32 + t);
    ptx::mbar_arrive(&bar->stage_free[b]);
  };

  if (warp == 0) {
    // ================================================================ TMA producer
    // The S tile of tile t+1 is requested before the V tile of tile t: an S slot is
    // released when S(t-1) retires, a V slot only when PV(t-2) retires (later), and the
    // hi/lo split of the next S tile must not wait behind the V ring.
    if (lane == 0) {
      auto load_s = [&](int t) {
        const int ss = t % C::kSSlots;
        if (t >= C::kSSlots) ptx::mbar_wait(&bar->s_free[ss], ((t / C::kSSlots) - 1) & 1);
        TC_TRACE(0, t);
        ptx::mbar_arrive_expect_tx(&bar->s_full[ss], C::kTileBytes);
        uint8_t* dst = s_ring + (size_t)ss * C::kSSlotBytes;
#pragma unroll
        for (int kb = 0; kb < C::kKB; ++kb)
          ptx::tma_load_2d(dst + kb * C::kBoxBytes, &tmap, &bar->s_full[ss], kb * 32, j_begin + t * kBlockJ);
      };
      auto load_v = [&](int t) {
        const int vs = t % C::kVSlots;
        if (t >= C::kVSlots) ptx::mbar_wait(&bar->v_free[vs], ((t / C::kVSlots) - 1) & 1);
        ptx::mbar_arrive_expect_tx(&bar->v_full[vs], C::kTileBytes);
        uint8_t* dst = v_ring + (size_t)vs * C::kTileBytes;
#pragma unroll
        for (int kb = 0; kb < C::kKB; ++kb)
          ptx::tma_load_2d(dst + kb * C::kBoxBytes, &tmap_v, &bar->v_full[vs], kb * 32, j_begin + t * kBlockJ);
      };
      // The S ring holds two tiles, so a TMA load is in flight for at most one tile time: enough when the queue comes
      // out of L2 or an idle HBM, not when the sweep shares HBM with a bandwidth-bound kernel (the momentum update
      // of the two-launch form: loaded latency of several microseconds).  The split's rows are contiguous in the
      // queue: pull them into L2 kPrefetchTiles tiles ahead, which costs no shared memory.
      constexpr int kPrefetchTiles = 6;
      auto prefetch = [&](int t) {
        const int j0 = j_begin + t * kBlockJ;
        if (t < n_tiles && j0 < j_end)
          ptx::bulk_prefetch_l2(p.queue + (size_t)j0 * D, (uint32_t)(min(kBlockJ, j_end - j0) * D * 4));
      };
      for (int t = 2; t < 2 + kPrefetchTiles; ++t) prefetch(t);
      if (n_tiles > 0) load_s(0);
      for (int t = 0; t < n_tiles; ++t) {
        prefetch(t + 2 + kPrefetchTiles);
        if (t + 1 < n_tiles) load_s(t + 1);
        if (!kThreeTerm) {
          if (t == 0) ptx::mbar_wait(&bar->q_ready, 0);  // the V ring doubles as the q staging area
          load_v(t);
        }
      }
    }
  } else if (warp == 1) {
    // ================================================================== MMA issuer
    // The whole warp runs this (warp-uniform) code; one elected lane issues the tcgen05
    // instructions, so descriptors and TMEM addresses stay in uniform registers.
    constexpr uint32_t idesc_s = ptx::umma_idesc_tf32(kM, kBlockJ, 0, 0);  // B = tile, K-major
    constexpr uint32_t idesc_pv = ptx::umma_idesc_tf32(kM, D, 0, 1);       // B = tile, MN-major
    constexpr uint32_t idesc_c = ptx::umma_idesc_bf16(kM, kBlockJ, 0, 0);  // B = correction tile, K-major
    const uint32_t tm = __shfl_sync(0xffffffffu, tmem, 0);
    const uint32_t s_ring0 = __shfl_sync(0xffffffffu, ptx::smem_u32(s_ring), 0);
    const uint32_t v_ring0 = __shfl_sync(0xffffffffu, ptx::smem_u32(v_ring), 0);
    ptx::mbar_wait_relaxed(&bar->q_ready, 0);
    ptx::tc_fence_after();
    TC_TRACE(13, 0);
    auto issue_pv = [&](int t) {
      const int vs = t % C::kVSlots, b = t & 1;
      ptx::mbar_wait(kThreeTerm ? &bar->v_op[vs] : &bar->v_full[vs], (t / C::kVSlots) & 1);
      ptx::mbar_wait(&bar->p_ready[b], (t >> 1) & 1);
      ptx::tc_fence_after();
      TC_TRACE(5, t);
      // MN-major, 32B-atom swizzle: 8 queue rows per k-step (1024 B) = two 4-row atoms 512 B apart
      // (SBO); the 32-float column blocks (one TMA box each) are kBoxBytes apart (LBO)
      const uint64_t bd0 = ptx::umma_smem_desc(v_ring0 + vs * C::kTileBytes, C::kBoxBytes, 512, ptx::kUmmaSwizzle128BBase32B);
      const uint32_t a0 = tm + C::kColS + b * kBlockJ;
      if (ptx::elect_one()) {
#pragma unroll
        for (int ks = 0; ks < kBlockJ / 8; ++ks)
          ptx::mma_tf32_ts(tm + C::kColAcc, a0 + ks * 8, bd0 + (uint64_t)(ks * 1024 >> 4), idesc_pv,
                           (t > 0 || ks > 0) ? 1u : 0u);
        TC_TRACE(13, 1 + t);
        ptx::tc_commit(&bar->v_free[vs]);  // V tile consumed -> TMA may refill the slot
        ptx::tc_commit(&bar->pv_done[b]);
        TC_TRACE(12, 1 + t);
      }
      __syncwarp();
    };
    for (int t = 0; t < n_tiles; ++t) {
      const int ss = t % C::kSSlots, b = t & 1;
      ptx::mbar_wait(kThreeTerm ? &bar->s_op[ss] : &bar->s_full[ss], (t / C::kSSlots) & 1);
      ptx::tc_fence_after();
      TC_TRACE(3, t);
      // K-major: k-block ks/4 (one TMA box), 32 bytes per k-step inside the 128-byte row
      const uint64_t hi0 = ptx::umma_smem_desc(s_ring0 + ss * C::kSSlotBytes, 16, 1024, ptx::kUmmaSwizzle128B);
      const uint64_t lo0 = hi0 + (uint64_t)(C::kTileBytes >> 4);
      const uint32_t d_s = tm + C::kColS + b * kBlockJ;
      if (ptx::elect_one()) {
        // Smallest contributions first: the tensor core's fp32 accumulation is not round-to-nearest,
        // so adding the ~2^-11 correction terms onto the finished q_hi.k_hi sum costs ~2x the error
        // of summing them first (measured: 3.8e-5 vs 1.7e-5 max logit error at T = 0.07).
        if (kThreeTerm) {
          // bf16 correction pass [q_lo | q].[k | k_lo]: 2D bf16 per row = the same D/32 boxes of 128 B,
          // 16 k per MMA = 32 B per step
#pragma unroll
          for (int ks = 0; ks < D / 8; ++ks)
            ptx::mma_f16_ts(d_s, tm + C::kColQlo + ks * 8, lo0 + (uint64_t)(((ks >> 2) * C::kBoxBytes + (ks & 3) * 32) >> 4),
                            idesc_c, ks > 0 ? 1u : 0u);
        }
#pragma unroll
        for (int ks = 0; ks < D / 8; ++ks)
          ptx::mma_tf32_ts(d_s, tm + C::kColQhi + ks * 8, hi0 + (uint64_t)(((ks >> 2) * C::kBoxBytes + (ks & 3) * 32) >> 4),
                           idesc_s, (kThreeTerm || ks > 0) ? 1u : 0u);
        ptx::tc_commit(&bar->s_ready[b]);
        ptx::tc_commit(&bar->s_free[ss]);  // raw + correction tile consumed -> TMA / helper warps may refill the slot
      }
      __syncwarp();
      TC_TRACE(4, t);
      if (t > 0) issue_pv(t - 1);  // softmax(t-1) overlapped the S(t) MMAs
    }
    if (n_tiles > 0) issue_pv(n_tiles - 1);
    if (ptx::elect_one()) ptx::tc_commit(&bar->acc_done);
    __syncwarp();
  } else if (warp < 10) {
    // ============================ helper warps 2..9 (3-term only): one read of the raw tile ->
    // correction tile [bf16(k) | bf16(k_lo)] and, from registers, the V tile rn_tf32(k)
    // + (staged CTAs, warps 4, 5, 8, 9) the logits of tile u: S(u) out of TMEM, scaled, into the staging tile
    const bool logits_duty = staged_cta && (warp == 4 || warp == 5 || warp == 8 || warp == 9);
    auto logits_tile = [&](int u) {
      const int b = u & 1, lsub = warp & 3, lhc = warp >> 3;  // sub-partition 0 / 1, column half 0 / 1
      ptx::mbar_wait(&bar->s_ready[b], (u >> 1) & 1);
      ptx::tc_fence_after();
      uint32_t sv[32];
      ptx::tmem_ld32(tmem + ((uint32_t)(lsub * 32) << 16) + C::kColS + b * kBlockJ + lhc * 32, sv);
      ptx::tc_wait_ld();
      ptx::tc_fence_before();
      ptx::mbar_arrive(&bar->s_read[b]);  // the softmax warps may overwrite S(u) with P(u)
      if (u >= 2) ptx::mbar_wait(&bar->stage_free[b], ((u >> 1) - 1) & 1);  // tile u-2 has left this buffer
      const int r = lsub * 32 + lane;
      if (r < p.B - i_base) {
        const float sc = bar->lscale[r];
        float* st_row = stage + b * (kStageRows * kStagePitch) + r * kStagePitch + lhc * 32;
#pragma unroll
        for (int c = 0; c < 32; ++c) st_row[c] = __uint_as_float(sv[c]) * sc;
      }
      ptx::mbar_arrive(&bar->stage_full[b]);
    };
    if (!kThreeTerm && logits_duty)
      for (int u = 0; u < n_tiles; ++u) logits_tile(u);
    if (kThreeTerm) {
      // st in [0,256) indexes the float4 this thread owns: a warp takes rows {w, w+4, w+8, w+12} of a
      // 16-row group (8 lanes per row) so that its 8-byte correction-tile stores land in both 64-byte
      // halves of the swizzled rows (no bank conflicts); whole rows keep the 128-bit raw loads and V
      // stores conflict-free too.
      const int hw = warp - 2;
      const int st = (((hw & 3) + 4 * (lane >> 3) + 16 * (hw >> 2)) << 3) | (lane & 7);
      constexpr int kPer = C::kTileBytes / 16 / kHelperThreads;  // float4 per thread and tile (D / 16)
      // Thread st owns the float4 e = st + 256 n of the raw tile (n < kPer).  In the S layout (128B
      // swizzle) e sits in box e >> 9 = n >> 1, row j = (st >> 3) + 32 (n & 1), physical 16-byte chunk
      // st & 7 = logical chunk c ^ (j & 7) -- so c, j & 7 and every swizzle term are per-thread
      // constants and the addresses below are "thread base + compile-time offset".
      const int jl = st >> 3, jc = jl & 7, c = (st & 7) ^ jc;
      // correction tile: element kappa of row j at byte 2 kappa = kbv * 64 + c * 8 (kbv = box of the fp32
      // column block, + D/32 for the k_lo half): box kbv >> 1, logical chunk (kbv & 1) * 4 + (c >> 1)
      const int corr_t = jl * 128 + (((c >> 1) ^ (jc & 3)) << 4) + (c & 1) * 8;
      const int corr_b[2] = {corr_t + ((jc & 4) << 4), corr_t + (((jc & 4) ^ 4) << 4)};
      // V layout (Swizzle<2,5,2>): 32-byte chunk (c >> 1) ^ (j & 3), same half
      const int v_t = (jl << 3) | ((((c >> 1) ^ (jl & 3)) << 1) | (c & 1));
      for (int t = 0; t < n_tiles; ++t) {
        const int ss = t % C::kSSlots, vs = t % C::kVSlots;
        ptx::mbar_wait_relaxed(&bar->s_full[ss], (t / C::kSSlots) & 1);
        if (st == 0) TC_TRACE(1, t);
        const float4* raw = reinterpret_cast<const float4*>(s_ring + (size_t)ss * C::kSSlotBytes) + st;
        uint8_t* corr = s_ring + (size_t)ss * C::kSSlotBytes + C::kTileBytes;  // [bf16(k) | bf16(k_lo)], K-major, 128B swizzle
        float4 h[kPer];
#pragma unroll
        for (int n = 0; n < kPer; ++n) {
          const float4 x = raw[n * kHelperThreads];
          uint2 kk, ll;
          kk.x = ptx::pack_bf16x2(x.x, x.y);
          kk.y = ptx::pack_bf16x2(x.z, x.w);
          // k_lo = k - trunc(k): exact in fp32, what the tf32 pass does not see
          ll.x = ptx::pack_bf16x2(x.x - ptx::trunc_tf32(x.x), x.y - ptx::trunc_tf32(x.y));
          ll.y = ptx::pack_bf16x2(x.z - ptx::trunc_tf32(x.z), x.w - ptx::trunc_tf32(x.w));
          constexpr int kRowOff = 32 * 128;  // 32 rows further per (n & 1)
          const int kb = n >> 1, kbl = kb + C::kKB;
          *reinterpret_cast<uint2*>(corr + (kb >> 1) * C::kBoxBytes + (n & 1) * kRowOff + corr_b[kb & 1]) = kk;
          *reinterpret_cast<uint2*>(corr + (kbl >> 1) * C::kBoxBytes + (n & 1) * kRowOff + corr_b[kbl & 1]) = ll;
          h[n] = make_float4(ptx::round_tf32(x.x), ptx::round_tf32(x.y), ptx::round_tf32(x.z), ptx::round_tf32(x.w));
        }
        ptx::fence_proxy_async_smem();
        if (st == 0) TC_TRACE(2, t);
        ptx::mbar_arrive(&bar->s_op[ss]);
        if (logits_duty && t >= 1) logits_tile(t - 1);  // S(t-1) is ready or about to be: the split runs one tile ahead
        // V tile = rn_tf32(k) in the 32B-atom layout, written from registers once PV(t - kVSlots)
        // has released the slot (and, for the first tiles, once q has left the aliased staging area)
        if (t == 0) ptx::mbar_wait_relaxed(&bar->q_ready, 0);
        if (t >= C::kVSlots) ptx::mbar_wait_relaxed(&bar->v_free[vs], ((t / C::kVSlots) - 1) & 1);
        float4* vt = reinterpret_cast<float4*>(v_ring + (size_t)vs * C::kTileBytes) + v_t;
#pragma unroll
        for (int n = 0; n < kPer; ++n) vt[(n >> 1) * 512 + (n & 1) * 256] = h[n];
        ptx::fence_proxy_async_smem();
        ptx::mbar_arrive(&bar->v_op[vs]);
      }
      if (logits_duty && n_tiles > 0) logits_tile(n_tiles - 1);
    }
  } else if (warp >= 14) {
    // ============================================================ logits writers 14, 15
    if (staged_cta) {
      for (int t = 0; t < n_tiles; ++t) drain_tile(t, warp - 12);
    }
  } else {
    // ==================================================== softmax + epilogue (thread = row)
    const int sub = warp & 3;                   // TMEM sub-partition of this warp
    const int r = sub * 32 + lane;              // row inside the CTA tile
    const int i = i_base + r;                   // global query row
    const bool row_valid = i < p.B;
    const bool warp_valid = (i_base + sub * 32) < p.B;  // warp-uniform
    const uint32_t lane_base = tmem + ((uint32_t)(sub * 32) << 16);

    // ---- A operand of the S GEMM: the RAW query features f, split into (hi, lo), go to TMEM;
    // the l2-normalisation q = f/||f|| is folded into this thread's softmax scale (the thread
    // owns row i, so s_ij = (f_i . k_j) / ||f_i||).  Every thread streams its own 512-byte row
    // with 128-bit loads (32 columns in flight ahead of the tcgen05.st of the current 32).
    float inv_norm = 0.f;
    {
      {
        const int stid = tid - 320;  // 0..127
        auto put = [&](int idx, const float4& v) {
          *reinterpret_cast<float4*>(q_stage + (size_t)(idx / kRowF4) * kQPitch + (idx % kRowF4) * 16) = v;
        };
#pragma unroll
        for (int n = 0; n < kQHalf; ++n) put(stid + n * kGroupThreads, q_buf[n]);
        if (q_total > kQHalf * kGroupThreads) {  // rows 64..127 of the CTA tile (B > 64)
#pragma unroll
          for (int n = 0; n < kQHalf; ++n) {
            const int idx = stid + (kQHalf + n) * kGroupThreads;
            q_buf[n] = idx < q_total ? __ldg(q_gsrc + idx) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
#pragma unroll
          for (int n = 0; n < kQHalf; ++n) put(stid + (kQHalf + n) * kGroupThreads, q_buf[n]);
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");  // the four softmax warps
      }
      if (r == 0) TC_TRACE(15, 1);
      const float4* q_src = reinterpret_cast<const float4*>(q_stage + (size_t)r * kQPitch);
      float4 ss = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int cb = 0; cb < D / 32; ++cb) {
        uint32_t vh[32], vlo[16], vq[16];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          float4 x = make_float4(0.f, 0.f, 0.f, 0.f);  // rows past B feed zeros to the tensor core
          if (row_valid) x = q_src[cb * 8 + k];
          ss.x = fmaf(x.x, x.x, ss.x);
          ss.y = fmaf(x.y, x.y, ss.y);
          ss.z = fmaf(x.z, x.z, ss.z);
          ss.w = fmaf(x.w, x.w, ss.w);
          const float hx = ptx::round_tf32(x.x), hy = ptx::round_tf32(x.y), hz = ptx::round_tf32(x.z),
                      hw = ptx::round_tf32(x.w);
          vh[4 * k + 0] = __float_as_uint(hx);
          vh[4 * k + 1] = __float_as_uint(hy);
          vh[4 * k + 2] = __float_as_uint(hz);
          vh[4 * k + 3] = __float_as_uint(hw);
          vlo[2 * k + 0] = ptx::pack_bf16x2(x.x - hx, x.y - hy);
          vlo[2 * k + 1] = ptx::pack_bf16x2(x.z - hz, x.w - hw);
          vq[2 * k + 0] = ptx::pack_bf16x2(x.x, x.y);
          vq[2 * k + 1] = ptx::pack_bf16x2(x.z, x.w);
        }
        ptx::tmem_st32(lane_base + C::kColQhi + cb * 32, vh);
        if (kThreeTerm) {  // A of the correction pass: k in [0,D) -> q_lo (meets bf16(k)), k in [D,2D) -> q (meets bf16(k_lo))
          ptx::tmem_st16(lane_base + C::kColQlo + cb * 16, vlo);
          ptx::tmem_st16(lane_base + C::kColQlo + D / 2 + cb * 16, vq);
        }
        if (r == 0) TC_TRACE(15, 2 + cb);
      }
      ptx::tc_wait_st();
      if (r == 0) TC_TRACE(15, 6);
      inv_norm = row_valid ? 1.f / sqrtf((ss.x + ss.y) + (ss.z + ss.w)) : 0.f;
      if (r < kStageRows) bar->lscale[r] = p.inv_T * inv_norm;  // for the logits warps (ordered by q_ready -> S(0) -> s_ready)
      ptx::tc_fence_before();
      if (r == 0) TC_TRACE(12, 0);
      ptx::mbar_arrive(&bar->q_ready);
      if (r == 0) TC_TRACE(11, 0);
    }
    const float scale2 = p.inv_T * kLog2e * inv_norm;  // log2-domain logit scale of this row (>= 0)
    const float logit_scale = p.inv_T * inv_norm;
    const bool drainer = staged_cta && sub >= 2;  // warp-uniform: this warp's rows do not exist

    float m_run = -INFINITY, l_run = 0.f;
    for (int t = 0; t < n_tiles; ++t) {
      const int b = t & 1;
      const int j0 = j_begin + t * kBlockJ;
      const int valid = min(kBlockJ, j_end - j0);  // columns of this tile that are real queue rows
      ptx::mbar_wait(&bar->s_ready[b], (t >> 1) & 1);
      ptx::tc_fence_after();
      if (r == 0) TC_TRACE(6, t);
      const uint32_t s_col = lane_base + C::kColS + b * kBlockJ;
      uint32_t sv[2 * 32];
      if (warp_valid) {
        // one TMEM round trip for the whole 64-column row of S
        ptx::tmem_ld32(s_col, *reinterpret_cast<uint32_t(*)[32]>(&sv[0]));
        ptx::tmem_ld32(s_col + 32, *reinterpret_cast<uint32_t(*)[32]>(&sv[32]));
        ptx::tc_wait_ld();
      }
      if (drainer) {
        // rows of this warp do not exist: release PV, then help storing the staged logits of tile t
        ptx::tc_fence_before();
        ptx::mbar_arrive(&bar->p_ready[b]);
        drain_tile(t, sub - 2);
        continue;
      }
      if (p.logits_out) {  // logits / T (models/contrastive.py:498)
        if (staged_cta) {
          // the logits warps read S(t) themselves (see above); nothing to do here
        } else if (row_valid) {
          for (int k = 0; k < p.n_keys; ++k) {
            float* dst = p.logits_out + ((size_t)k * p.B + i) * (size_t)(p.K + 1) + 1 + j0;
#pragma unroll
            for (int c = 0; c < kBlockJ; ++c)
              if (c < valid) dst[c] = __uint_as_float(sv[c]) * logit_scale;
          }
        }
      }
      if (warp_valid) {
        if (valid < kBlockJ) {  // ragged last tile: columns past the queue end never win the max and get P = 0
#pragma unroll
          for (int c = 0; c < kBlockJ; ++c)
            if (c >= valid) sv[c] = 0xff800000u;  // -inf
        }
        float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int c = 0; c < kBlockJ; c += 4) {
#pragma unroll
          for (int u = 0; u < 4; ++u) mx[u] = fmaxf(mx[u], __uint_as_float(sv[c + u]));
        }
        // scale2 >= 0, so the maximum commutes with the scaling (rows past B have scale2 = 0)
        const float tmax = row_valid ? fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3])) * scale2 : 0.f;
        if (r == 0) TC_TRACE(8, t);
        // lazy rescale: keep the reference maximum unless it falls more than 2^8 behind
        const bool need = tmax > m_run + kRescaleThreshold;
        if (t == 0) {
          m_run = tmax;
        } else if (__any_sync(0xffffffffu, need)) {
          const float m_new = need ? tmax : m_run;
          const float alpha = exp2f(m_run - m_new);
          // every PV MMA issued so far must have landed before acc is rewritten
          ptx::mbar_wait(&bar->pv_done[(t - 1) & 1], ((t - 1) >> 1) & 1);
          ptx::tc_fence_after();
#pragma unroll 1
          for (int cb = 0; cb < D / 32; ++cb) {
            uint32_t av[32];
            ptx::tmem_ld32(lane_base + C::kColAcc + cb * 32, av);
            ptx::tc_wait_ld();
#pragma unroll
            for (int c = 0; c < 32; ++c) av[c] = __float_as_uint(__uint_as_float(av[c]) * alpha);
            ptx::tmem_st32(lane_base + C::kColAcc + cb * 32, av);
          }
          ptx::tc_wait_st();
          l_run *= alpha;
          m_run = m_new;
        }
        if (r == 0) TC_TRACE(9, t);
        // P = 2^(s2 - m), rounded to tf32, written over S
        float ps[4] = {0.f, 0.f, 0.f, 0.f};
        const float neg_m = -m_run;
#pragma unroll
        for (int c = 0; c < kBlockJ; c += 4) {
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            float pv;
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(pv) : "f"(fmaf(__uint_as_float(sv[c + u]), scale2, neg_m)));
            ps[u] += pv;
            sv[c + u] = __float_as_uint(ptx::round_tf32(pv));
          }
        }
        if (staged_cta) ptx::mbar_wait(&bar->s_read[b], (t >> 1) & 1);  // S(t) has been read for the logits
        ptx::tmem_st32(s_col, *reinterpret_cast<uint32_t(*)[32]>(&sv[0]));
        ptx::tmem_st32(s_col + 32, *reinterpret_cast<uint32_t(*)[32]>(&sv[32]));
        l_run += (ps[0] + ps[1]) + (ps[2] + ps[3]);
        ptx::tc_wait_st();
      }
      ptx::tc_fence_before();
      if (r == 0) TC_TRACE(7, t);
      ptx::mbar_arrive(&bar->p_ready[b]);
    }

    // ---- epilogue: partial (m, l, acc) of this split
    ptx::mbar_wait(&bar->acc_done, 0);
    ptx::tc_fence_after();
    if (warp_valid) {
      const size_t row = (size_t)split * p.B + (row_valid ? i : 0);
      if (row_valid) {
        p.part_m[row] = m_run;
        p.part_l[row] = l_run;
      }
#pragma unroll 1
      for (int cb = 0; cb < D / 32; ++cb) {
        uint32_t av[32];
        ptx::tmem_ld32(lane_base + C::kColAcc + cb * 32, av);
        ptx::tc_wait_ld();
        if (row_valid) {
          float4* dst = reinterpret_cast<float4*>(p.part_acc + row * D + cb * 32);
#pragma unroll
          for (int c4 = 0; c4 < 8; ++c4)
            __stcg(dst + c4, make_float4(__uint_as_float(av[c4 * 4]), __uint_as_float(av[c4 * 4 + 1]),
                                         __uint_as_float(av[c4 * 4 + 2]), __uint_as_float(av[c4 * 4 + 3])));
        }
      }
    }
  }

  // ------------------------------------------------- teardown of the tensor-core sweep
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem, C::kTmemCols);
  }

  }  // !exch_cta

  if (p.phase == kPhaseSweep) return;  // the partials are complete when the launch is; infonce_finish_kernel follows

  // --------------------------------------------------------------------- grid barrier
  // Cooperative launch: all CTAs are co-resident, so spinning on a global counter is safe.
  const unsigned n_ctas = gridDim.x * gridDim.y;
  if (tid == 0) TC_TRACE(14, 2);
  if (tid == 0) {
    __threadfence();  // this CTA's partials (ordered before by the barrier above) are visible GPU-wide
    atomicAdd(p.counter + 1, 1u);
    while (ld_acquire_u32(p.counter + 1) < n_ctas) __nanosleep(40);
  }
  __syncthreads();

  CombineSmem<kTcThreads>& csm = *reinterpret_cast<CombineSmem<kTcThreads>*>(smem);  // the rings are dead
  infonce_tail<kTcThreads>(p, csm, blockIdx.y * gridDim.x + blockIdx.x, n_ctas);
}

// One CTA per query row: M = max_s m_s, w_s = 2^(m_s - M), L = sum_s w_s l_s, acc = sum_s w_s acc_s -- the first half of
// infonce_combine_row, in a fixed order (deterministic).  Runs under the momentum update in the two-launch form, so
// that the launch behind the key path has one partial per row left to read.
constexpr int kMergeThreads = 256;
__global__ void __launch_bounds__(kMergeThreads)
infonce_merge_partials_kernel(const InfoNceParams p, float* __restrict__ m_out, float* __restrict__ l_out,
                              float* __restrict__ acc_out) {
  __shared__ float s_w[kMaxSplits];
  __shared__ float s_red[32];
  __shared__ __align__(16) float s_acc[kMergeThreads / 32][256];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int i = blockIdx.x, S = p.n_splits, B = p.B, D = p.D;
  float mloc = -INFINITY;
  for (int s = tid; s < S; s += kMergeThreads) mloc = fmaxf(mloc, __ldcg(p.part_m + (size_t)s * B + i));
  mloc = warp_max(mloc);
  if (lane == 0) s_red[warp] = mloc;
  __syncthreads();
  float M = s_red[0];
#pragma unroll
  for (int w = 1; w < kMergeThreads / 32; ++w) M = fmaxf(M, s_red[w]);
  float lloc = 0.f;
  for (int s = tid; s < S; s += kMergeThreads) {
    const float w = fast_ex2(__ldcg(p.part_m + (size_t)s * B + i) - M);
    s_w[s] = w;
    lloc = fmaf(w, __ldcg(p.part_l + (size_t)s * B + i), lloc);
  }
  const float L = block_sum(lloc, s_red);  // (its barriers publish s_w)
  // warp g takes splits g, g + 8, ...; lane l owns columns 4l..4l+3 (and 128 + 4l.. for D > 128)
  const size_t stride = (size_t)B * D;
  for (int c0 = 0; c0 < D; c0 += 128) {
    const int c = c0 + lane * 4;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c < D) {
      const float* src = p.part_acc + (size_t)i * D + c;
#pragma unroll 4
      for (int s = warp; s < S; s += kMergeThreads / 32) {
        const float4 v = __ldcg(reinterpret_cast<const float4*>(src + (size_t)s * stride));
        const float w = s_w[s];
        a.x = fmaf(w, v.x, a.x);
        a.y = fmaf(w, v.y, a.y);
        a.z = fmaf(w, v.z, a.z);
        a.w = fmaf(w, v.w, a.w);
      }
      *reinterpret_cast<float4*>(&s_acc[warp][c]) = a;
    }
  }
  __syncthreads();
  for (int c = tid; c < D; c += kMergeThreads) {
    float a = s_acc[0][c];
#pragma unroll
    for (int g = 1; g < kMergeThreads / 32; ++g) a += s_acc[g][c];
    acc_out[(size_t)i * D + c] = a;
  }
  if (tid == 0) {
    m_out[i] = M;
    l_out[i] = L;
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(sym);
  }
  return fn;
}

// The two tensor maps only depend on (queue pointer, K, D): keep the last pair per thread.
struct TmapCache {
  const float* queue = nullptr;
  int K = 0, D = 0;
  CUtensorMap s, v;
};

template <int D, bool kThreeTerm>
int launch_tc(const InfoNceParams& p, cudaStream_t s) {
  using C = TcCfg<D, kThreeTerm>;
  static thread_local TmapCache cache;
  if (cache.queue != p.queue || cache.K != p.K || cache.D != D) {
    EncodeTiledFn enc = encode_fn();
    AVSSL_REQUIRE(enc, AVSSL_ERR_CUDA, "moco_infonce: cuTensorMapEncodeTiled is not available from the driver");
    const cuuint64_t gdim[2] = {(cuuint64_t)D, (cuuint64_t)p.K};
    const cuuint64_t gstride[1] = {(cuuint64_t)D * sizeof(float)};
    const cuuint32_t box[2] = {32u, (cuuint32_t)kBlockJ};
    const cuuint32_t estride[2] = {1u, 1u};
    CUresult r = enc(&cache.s, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(p.queue), gdim, gstride, box,
                     estride, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    AVSSL_REQUIRE(r == CUDA_SUCCESS, AVSSL_ERR_CUDA, "moco_infonce: cuTensorMapEncodeTiled failed (%d)", (int)r);
    r = enc(&cache.v, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(p.queue), gdim, gstride, box, estride,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    AVSSL_REQUIRE(r == CUDA_SUCCESS, AVSSL_ERR_CUDA, "moco_infonce: cuTensorMapEncodeTiled (32B atoms) failed (%d)", (int)r);
    cache.queue = p.queue;
    cache.K = p.K;
    cache.D = D;
  }
  static unsigned long long configured = 0ull;  // device ordinals already set up
  if (first_use_on_device(configured)) {
    AVSSL_CUDA_OK(cudaFuncSetAttribute(infonce_tc_kernel<D, kThreeTerm>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)C::kSmemBytes));
  }
  dim3 grid(p.n_splits + (p.push_feat ? 1 : 0), (p.B + kM - 1) / kM);
  InfoNceParams pc = p;
  if (p.phase == kPhaseSweep) {  // no grid barrier: a plain launch, any number of CTAs
    infonce_tc_kernel<D, kThreeTerm><<<grid, kTcThreads, C::kSmemBytes, s>>>(pc, cache.s, cache.v);
    const cudaError_t le = cudaGetLastError();
    if (le != cudaSuccess) {
      set_error("launch of infonce_tc_kernel (sweep) failed: %s", cudaGetErrorString(le));
      return AVSSL_ERR_CUDA;
    }
    return AVSSL_OK;
  }
  AVSSL_REQUIRE((int)(grid.x * grid.y) <= sm_count(), AVSSL_ERR_INVALID_ARGUMENT,
                "moco_infonce: %u CTAs cannot be co-resident on %d SMs", grid.x * grid.y, sm_count());
  // cooperative: the kernel contains a grid-wide barrier (all CTAs must be co-resident)
  void* args[] = {&pc, &cache.s, &cache.v};
#ifdef AVSSL_TC_TRACE
  if (getenv("AVSSL_TC_PLAIN_LAUNCH")) {  // developer probe: cost of the cooperative launch itself
    infonce_tc_kernel<D, kThreeTerm><<<grid, kTcThreads, C::kSmemBytes, s>>>(pc, cache.s, cache.v);
    return AVSSL_OK;
  }
#endif
  cudaError_t le = cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(&infonce_tc_kernel<D, kThreeTerm>), grid,
                                               dim3(kTcThreads), args, C::kSmemBytes, s);
  if (le != cudaSuccess) {
    set_error("cooperative launch of infonce_tc_kernel failed: %s", cudaGetErrorString(le));
    return AVSSL_ERR_CUDA;
  }
  return AVSSL_OK;
}

}  // namespace

int launch_infonce_merge_partials(const InfoNceParams& p, float* m_out, float* l_out, float* acc_out, cudaStream_t s) {
  AVSSL_REQUIRE(p.D % 4 == 0 && p.D <= 256 && p.n_splits <= kMaxSplits, AVSSL_ERR_UNSUPPORTED,
                "moco_infonce: cannot merge %d partials of width %d", p.n_splits, p.D);
  infonce_merge_partials_kernel<<<p.B, kMergeThreads, 0, s>>>(p, m_out, l_out, acc_out);
  const cudaError_t le = cudaGetLastError();
  if (le != cudaSuccess) {
    set_error("launch of infonce_merge_partials_kernel failed: %s", cudaGetErrorString(le));
    return AVSSL_ERR_CUDA;
  }
  return AVSSL_OK;
}

// Second launch of the two-launch form: the tail of the cooperative kernel as a small kernel of its own (one CTA per
// query row / enqueued row; with push_feat one more CTA that performs this rank's key push first).  p.part_* hold
// ONE merged partial per row (n_splits == 1).
constexpr int kFinishThreads = 256;
__global__ void __launch_bounds__(kFinishThreads) infonce_finish_kernel(const InfoNceParams p) {
  __shared__ CombineSmem<kFinishThreads> csm;
  // With push_feat the FIRST `world` CTAs perform this rank's key push (C3: Normalize + stores into rank d's buffer +
  // flag, CTA d -> rank d).  The wait in the tail reads the local epoch, which the last push CTA advances: every CTA
  // first waits for the `world` pushes of its own launch (counter[2]; the push CTAs have the lowest block indices, so
  // they are resident before any CTA that spins on them).
  const int n_push = p.push_feat != nullptr ? p.peer.world : 0;
  if ((int)blockIdx.x < n_push) {
    __shared__ unsigned long long s_epoch;
    peer_push_cta<true>(p.peer, p.push_feat, (int)blockIdx.x, &s_epoch, p.push_eps);
    if (threadIdx.x == 0) {  // (thread 0 published the flag and, in the last CTA, advanced the epoch just above)
      __threadfence();
      atomicAdd(p.counter + 2, 1u);
    }
  }
  if (n_push > 0) {
    if (threadIdx.x == 0)
      while (ld_acquire_u32(p.counter + 2) < (unsigned)n_push) __nanosleep(20);
    __syncthreads();
  }
  const int cta = ((int)blockIdx.x - n_push + (int)gridDim.x) % (int)gridDim.x;  // the push CTAs take the last rows, if any
  infonce_tail<kFinishThreads>(p, csm, cta, gridDim.x);
}

int launch_infonce_finish(const InfoNceParams& p, cudaStream_t s) {
  AVSSL_REQUIRE(p.D % 4 == 0 && p.D <= kCombineCols && p.n_splits == 1, AVSSL_ERR_UNSUPPORTED,
                "moco_infonce: finish launch needs one merged partial per row and D <= %d", kCombineCols);
  const int rows = p.enq_ptr && p.n_enq > p.B ? p.n_enq : p.B;
  const int cap = 4 * sm_count();
  const unsigned grid = (unsigned)(rows < cap ? rows : cap) + (p.push_feat ? (unsigned)p.peer.world : 0u);
  infonce_finish_kernel<<<grid, kFinishThreads, 0, s>>>(p);
  const cudaError_t le = cudaGetLastError();
  if (le != cudaSuccess) {
    set_error("launch of infonce_finish_kernel failed: %s", cudaGetErrorString(le));
    return AVSSL_ERR_CUDA;
  }
  return AVSSL_OK;
}

bool infonce_tc_supported(int B, int D, int K) {
  // one CTA per 128 query rows and queue split; the cooperative grid must fit the SMs
  return (D == 32 || D == 64 || D == 96 || D == 128) && K >= 1 && (B + kM - 1) / kM <= sm_count();
}

int launch_infonce_tc(const InfoNceParams& p, int three_term, cudaStream_t s) {
  AVSSL_REQUIRE(infonce_tc_supported(p.B, p.D, p.K), AVSSL_ERR_UNSUPPORTED,
                "moco_infonce: the tcgen05 kernel needs D in {32,64,96,128} (got %d); use AVSSL_IMPL_SIMT", p.D);
  AVSSL_REQUIRE(p.rows_per_split % kBlockJ == 0, AVSSL_ERR_INVALID_ARGUMENT, "moco_infonce: split is not tile aligned");
#define AVSSL_TC_CASE(DD)                                                  \
  case DD:                                                                 \
    return three_term ? launch_tc<DD, true>(p, s) : launch_tc<DD, false>(p, s);
  switch (p.D) {
    AVSSL_TC_CASE(32)
    AVSSL_TC_CASE(64)
    AVSSL_TC_CASE(96)
    AVSSL_TC_CASE(128)
  }
#undef AVSSL_TC_CASE
  return AVSSL_ERR_UNSUPPORTED;
}

}  // namespace avssl
