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
// while the K-major operand of the first GEMM needs the plain 128B swizzle, so TMA lands
// every tile twice (two tensor maps over the same global rows; the second read is an L2
// hit): the "S tile" ring feeds the first GEMM, the "V tile" ring the second.  An S slot is
// recycled as soon as its S GEMM retires, a V slot when its PV GEMM retires.
//
// fp32-grade accuracy on tf32 tensor cores (kThreeTerm): four helper warps split the S
// tile in shared memory into hi = rn_tf32(x) and lo = x - hi, q likewise (in TMEM), and
// S = q_lo.k_hi + q_hi.k_lo + q_hi.k_hi (error ~2^-22, no truncation bias); four more
// warps round the V tile to tf32 (round-to-nearest instead of the hardware's truncation).
// Without the split (kThreeTerm = false) the hardware truncates the queue operand to 10
// mantissa bits: 3x fewer S MMAs, tf32-grade logits.
//
// Measured on B200 (tools/microbench): one thread issues a tcgen05.mma every ~55 cycles at
// best and an M=128, N=64, K=8 tf32 MMA occupies the tensor pipe for 32 cycles, so the issue
// loop is kept warp-uniform (descriptors in uniform registers, one elected lane) and fully
// unrolled.
//
// Warp roles (448 threads, 1 CTA / SM):  0 TMA producer | 1 MMA issuer + TMEM allocator |
// 2-5 S-tile hi/lo split | 6-9 V-tile rounding | 10-13 softmax + epilogue (thread = query row).
// TMEM columns: [0,D) q_hi | [D,2D) q_lo | [2D,2D+128) S/P double buffer | [2D+128,3D+128) acc.
#include "infonce.cuh"
#include "sm100_ptx.cuh"

namespace avssl {

#ifdef AVSSL_TC_TRACE
// developer build only (tools/microbench/tc_trace.cu): per-role timestamps of CTA (0,0)
__device__ long long g_tc_trace[16][64];
#define TC_TRACE(ev, t)                                                                \
  do {                                                                                 \
    if (blockIdx.x == 0 && blockIdx.y == 0 && (t) < 64) g_tc_trace[ev][t] = clock64(); \
  } while (0)
#else
#define TC_TRACE(ev, t) \
  do {                  \
  } while (0)
#endif

namespace {

constexpr int kBlockJ = kTcTileRows;  // 64 queue rows per tile
constexpr int kM = 128;               // query rows per CTA
constexpr int kTcThreads = 448;
constexpr int kGroupThreads = 128;    // split / round / softmax groups
constexpr float kRescaleThreshold = 8.f;  // log2 units: P stays below 2^8
constexpr int kMaxSlots = 3;

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
  static constexpr int kScratchBytes = 4 * 32 * 33 * 4;  // q transpose scratch, one [32][33] per softmax warp
  static constexpr int kColQhi = 0, kColQlo = D, kColS = 2 * D, kColAcc = 2 * D + 2 * kBlockJ;
  static constexpr int kTmemCols = 512;
  static_assert(3 * D + 2 * kBlockJ <= 512, "TMEM budget");
  static constexpr size_t kSmemBytes = 1024 /*align slack*/ + (size_t)kSRingBytes + kVRingBytes + kScratchBytes + 512;
};

struct TcBarriers {
  uint64_t s_full[kMaxSlots], s_op[kMaxSlots], s_free[kMaxSlots];
  uint64_t v_full[kMaxSlots], v_op[kMaxSlots], v_free[kMaxSlots];
  uint64_t s_ready[2], p_ready[2], pv_done[2];
  uint64_t q_ready, acc_done;
  uint32_t tmem_base;
};

template <int D, bool kThreeTerm>
__global__ void __launch_bounds__(kTcThreads, 1)
infonce_tc_kernel(const InfoNceParams p, const __grid_constant__ CUtensorMap tmap,
                  const __grid_constant__ CUtensorMap tmap_v) {
  using C = TcCfg<D, kThreeTerm>;
  extern __shared__ uint8_t smem_raw[];
  // swizzled operands need 1024-byte aligned tiles
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* s_ring = smem;
  uint8_t* v_ring = smem + C::kSRingBytes;
  float* scratch = reinterpret_cast<float*>(smem + C::kSRingBytes + C::kVRingBytes);
  TcBarriers* bar = reinterpret_cast<TcBarriers*>(smem + C::kSRingBytes + C::kVRingBytes + C::kScratchBytes);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int split = blockIdx.x;
  const int i_base = blockIdx.y * kM;
  const int j_begin = split * p.rows_per_split;
  const int j_end = min(p.K, j_begin + p.rows_per_split);
  const int n_tiles = (j_end - j_begin + kBlockJ - 1) / kBlockJ;

  // ------------------------------------------------------------------ one-time setup
  if (warp == 0 && lane == 0) {
    ptx::tma_prefetch_desc(&tmap);
    ptx::tma_prefetch_desc(&tmap_v);
    for (int s = 0; s < kMaxSlots; ++s) {
      ptx::mbar_init(&bar->s_full[s], 1);
      ptx::mbar_init(&bar->s_op[s], kGroupThreads);
      ptx::mbar_init(&bar->s_free[s], 1);
      ptx::mbar_init(&bar->v_full[s], 1);
      ptx::mbar_init(&bar->v_op[s], kGroupThreads);
      ptx::mbar_init(&bar->v_free[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      ptx::mbar_init(&bar->s_ready[b], 1);
      ptx::mbar_init(&bar->p_ready[b], kGroupThreads);
      ptx::mbar_init(&bar->pv_done[b], 1);
    }
    ptx::mbar_init(&bar->q_ready, kGroupThreads);
    ptx::mbar_init(&bar->acc_done, 1);
    ptx::mbar_fence_init();
  }
  if (warp == 1) ptx::tmem_alloc(&bar->tmem_base, C::kTmemCols);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = bar->tmem_base;

  if (warp == 0) {
    // ================================================================ TMA producer
    if (lane == 0) {
      for (int t = 0; t < n_tiles; ++t) {
        const int ss = t % C::kSSlots, vs = t % C::kVSlots;
        const int j0 = j_begin + t * kBlockJ;
        if (t >= C::kSSlots) ptx::mbar_wait(&bar->s_free[ss], ((t / C::kSSlots) - 1) & 1);
        TC_TRACE(0, t);
        ptx::mbar_arrive_expect_tx(&bar->s_full[ss], C::kTileBytes);
        uint8_t* dst = s_ring + (size_t)ss * C::kSSlotBytes;
#pragma unroll
        for (int kb = 0; kb < C::kKB; ++kb) ptx::tma_load_2d(dst + kb * C::kBoxBytes, &tmap, &bar->s_full[ss], kb * 32, j0);
        if (t >= C::kVSlots) ptx::mbar_wait(&bar->v_free[vs], ((t / C::kVSlots) - 1) & 1);
        ptx::mbar_arrive_expect_tx(&bar->v_full[vs], C::kTileBytes);
        dst = v_ring + (size_t)vs * C::kTileBytes;
#pragma unroll
        for (int kb = 0; kb < C::kKB; ++kb) ptx::tma_load_2d(dst + kb * C::kBoxBytes, &tmap_v, &bar->v_full[vs], kb * 32, j0);
      }
    }
  } else if (warp == 1) {
    // ================================================================== MMA issuer
    // The whole warp runs this (warp-uniform) code; one elected lane issues the tcgen05
    // instructions, so descriptors and TMEM addresses stay in uniform registers.
    constexpr uint32_t idesc_s = ptx::umma_idesc_tf32(kM, kBlockJ, 0, 0);  // B = tile, K-major
    constexpr uint32_t idesc_pv = ptx::umma_idesc_tf32(kM, D, 0, 1);       // B = tile, MN-major
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
        ptx::tc_commit(&bar->v_free[vs]);  // V tile consumed -> TMA may refill the slot
        ptx::tc_commit(&bar->pv_done[b]);
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
        // smallest contributions first: q_lo.k_hi, q_hi.k_lo, then q_hi.k_hi
        if (kThreeTerm) {
#pragma unroll
          for (int ks = 0; ks < D / 8; ++ks)
            ptx::mma_tf32_ts(d_s, tm + C::kColQlo + ks * 8, hi0 + (uint64_t)(((ks >> 2) * C::kBoxBytes + (ks & 3) * 32) >> 4),
                             idesc_s, ks > 0 ? 1u : 0u);
#pragma unroll
          for (int ks = 0; ks < D / 8; ++ks)
            ptx::mma_tf32_ts(d_s, tm + C::kColQhi + ks * 8, lo0 + (uint64_t)(((ks >> 2) * C::kBoxBytes + (ks & 3) * 32) >> 4),
                             idesc_s, 1u);
        }
#pragma unroll
        for (int ks = 0; ks < D / 8; ++ks)
          ptx::mma_tf32_ts(d_s, tm + C::kColQhi + ks * 8, hi0 + (uint64_t)(((ks >> 2) * C::kBoxBytes + (ks & 3) * 32) >> 4),
                           idesc_s, (kThreeTerm || ks > 0) ? 1u : 0u);
        ptx::tc_commit(&bar->s_ready[b]);
        ptx::tc_commit(&bar->s_free[ss]);  // S tile consumed -> TMA may refill the slot
      }
      __syncwarp();
      TC_TRACE(4, t);
      if (t > 0) issue_pv(t - 1);  // softmax(t-1) overlapped the S(t) MMAs
    }
    if (n_tiles > 0) issue_pv(n_tiles - 1);
    if (ptx::elect_one()) ptx::tc_commit(&bar->acc_done);
    __syncwarp();
  } else if (warp < 6) {
    // ===================================================== S-tile hi/lo split (3-term only)
    if (kThreeTerm) {
      const int st = tid - 64;  // 0..127
      for (int t = 0; t < n_tiles; ++t) {
        const int ss = t % C::kSSlots;
        ptx::mbar_wait_relaxed(&bar->s_full[ss], (t / C::kSSlots) & 1);
        if (st == 0) TC_TRACE(1, t);
        float4* hi = reinterpret_cast<float4*>(s_ring + (size_t)ss * C::kSSlotBytes);
        float4* lo = reinterpret_cast<float4*>(s_ring + (size_t)ss * C::kSSlotBytes + C::kTileBytes);
#pragma unroll 8
        for (int e = st; e < C::kTileBytes / 16; e += kGroupThreads) {
          const float4 x = hi[e];
          float4 h, l;
          h.x = ptx::round_tf32(x.x);
          h.y = ptx::round_tf32(x.y);
          h.z = ptx::round_tf32(x.z);
          h.w = ptx::round_tf32(x.w);
          l.x = x.x - h.x;
          l.y = x.y - h.y;
          l.z = x.z - h.z;
          l.w = x.w - h.w;
          hi[e] = h;
          lo[e] = l;
        }
        ptx::fence_proxy_async_smem();
        if (st == 0) TC_TRACE(2, t);
        ptx::mbar_arrive(&bar->s_op[ss]);
      }
    }
  } else if (warp < 10) {
    // ======================= V-tile rounding (3-term only): unbiased rn instead of truncation
    if (kThreeTerm) {
      const int st = tid - 192;  // 0..127
      for (int t = 0; t < n_tiles; ++t) {
        const int vs = t % C::kVSlots;
        ptx::mbar_wait_relaxed(&bar->v_full[vs], (t / C::kVSlots) & 1);
        float4* vt = reinterpret_cast<float4*>(v_ring + (size_t)vs * C::kTileBytes);
#pragma unroll 8
        for (int e = st; e < C::kTileBytes / 16; e += kGroupThreads) {
          float4 x = vt[e];
          x.x = ptx::round_tf32(x.x);
          x.y = ptx::round_tf32(x.y);
          x.z = ptx::round_tf32(x.z);
          x.w = ptx::round_tf32(x.w);
          vt[e] = x;
        }
        ptx::fence_proxy_async_smem();
        ptx::mbar_arrive(&bar->v_op[vs]);
      }
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
    // owns row i, so s_ij = (f_i . k_j) / ||f_i||).  Cooperative and coalesced: lanes stride the
    // columns of one row at a time (the summation order of warp_row_norm(), so the norm is
    // bit-identical to the combine kernel's), a [32][33] shared-memory transpose hands every
    // thread its own row for tcgen05.st, and the next 32-column block is loaded while the current
    // one is transposed.  The MMA warp is released before the norms are reduced.
    float inv_norm = 0.f;
    {
      float* sc = scratch + sub * (32 * 33);
      const int row0 = i_base + sub * 32;
      float ssq[32], fv[32], fn[32];
#pragma unroll
      for (int rr = 0; rr < 32; ++rr) ssq[rr] = 0.f;
      if (warp_valid) {
#pragma unroll
        for (int rr = 0; rr < 32; ++rr)  // rows past B re-read the last valid row (branch-free) and are masked below
          fv[rr] = __ldg(p.feat_q + (size_t)min(row0 + rr, p.B - 1) * D + lane);
      }
#pragma unroll
      for (int cb = 0; cb < D / 32; ++cb) {
        uint32_t v[32];
        if (warp_valid) {
          if (cb + 1 < D / 32) {
#pragma unroll
            for (int rr = 0; rr < 32; ++rr)
              fn[rr] = __ldg(p.feat_q + (size_t)min(row0 + rr, p.B - 1) * D + (cb + 1) * 32 + lane);
          }
#pragma unroll
          for (int rr = 0; rr < 32; ++rr) {
            ssq[rr] = fmaf(fv[rr], fv[rr], ssq[rr]);
            sc[rr * 33 + lane] = (row0 + rr) < p.B ? fv[rr] : 0.f;
          }
          __syncwarp();
#pragma unroll
          for (int c = 0; c < 32; ++c) v[c] = __float_as_uint(ptx::round_tf32(sc[lane * 33 + c]));
        } else {
#pragma unroll
          for (int c = 0; c < 32; ++c) v[c] = 0u;
        }
        ptx::tmem_st32(lane_base + C::kColQhi + cb * 32, v);
        if (kThreeTerm) {
          if (warp_valid) {
#pragma unroll
            for (int c = 0; c < 32; ++c) {
              const float x = sc[lane * 33 + c];
              v[c] = __float_as_uint(x - ptx::round_tf32(x));
            }
          }
          ptx::tmem_st32(lane_base + C::kColQlo + cb * 32, v);
        }
        if (warp_valid) {
          __syncwarp();
#pragma unroll
          for (int rr = 0; rr < 32; ++rr) fv[rr] = fn[rr];
        }
      }
      ptx::tc_wait_st();
      ptx::tc_fence_before();
      if (r == 0) TC_TRACE(12, 0);
      ptx::mbar_arrive(&bar->q_ready);
      if (warp_valid) {
        // transpose the per-lane partial sums, then every lane reduces ITS row with the same
        // 16-8-4-2-1 pairing as warp_sum()'s xor butterfly (bit-identical to warp_row_norm)
#pragma unroll
        for (int rr = 0; rr < 32; ++rr) sc[rr * 33 + lane] = ssq[rr];
        __syncwarp();
#pragma unroll
        for (int c = 0; c < 32; ++c) ssq[c] = sc[lane * 33 + c];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
#pragma unroll
          for (int l = 0; l < o; ++l) ssq[l] = ssq[l] + ssq[l + o];
        inv_norm = row_valid ? 1.f / sqrtf(ssq[0]) : 0.f;
      }
      if (r == 0) TC_TRACE(11, 0);
    }
    const float scale2 = p.inv_T * kLog2e * inv_norm;  // log2-domain logit scale of this row
    const float logit_scale = p.inv_T * inv_norm;

    float m_run = -INFINITY, l_run = 0.f;
    for (int t = 0; t < n_tiles; ++t) {
      const int b = t & 1;
      const int j0 = j_begin + t * kBlockJ;
      const int valid = min(kBlockJ, j_end - j0);  // columns of this tile that are real queue rows
      ptx::mbar_wait(&bar->s_ready[b], (t >> 1) & 1);
      ptx::tc_fence_after();
      if (r == 0) TC_TRACE(6, t);
      if (warp_valid) {
        const uint32_t s_col = lane_base + C::kColS + b * kBlockJ;
        // one TMEM round trip for the whole 64-column row of S
        uint32_t sv[2][32];
        ptx::tmem_ld32(s_col, sv[0]);
        ptx::tmem_ld32(s_col + 32, sv[1]);
        ptx::tc_wait_ld();
        if (p.logits_out && row_valid) {  // logits / T (models/contrastive.py:498)
          for (int k = 0; k < p.n_keys; ++k) {
            float* dst = p.logits_out + ((size_t)k * p.B + i) * (size_t)(p.K + 1) + 1 + j0;
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
              for (int c = 0; c < 32; ++c)
                if (h * 32 + c < valid) dst[h * 32 + c] = __uint_as_float(sv[h][c]) * logit_scale;
          }
        }
        float tmax = -INFINITY;
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
          for (int c = 0; c < 32; ++c)
            if (h * 32 + c < valid) tmax = fmaxf(tmax, __uint_as_float(sv[h][c]) * scale2);
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
        float psum = 0.f;
        const float neg_m = -m_run;
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
          for (int c = 0; c < 32; ++c) {
            float pv;
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(pv) : "f"(fmaf(__uint_as_float(sv[h][c]), scale2, neg_m)));
            pv = (h * 32 + c < valid) ? pv : 0.f;
            psum += pv;
            sv[h][c] = __float_as_uint(ptx::round_tf32(pv));
          }
        ptx::tmem_st32(s_col, sv[0]);
        ptx::tmem_st32(s_col + 32, sv[1]);
        l_run += psum;
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
            dst[c4] = make_float4(__uint_as_float(av[c4 * 4]), __uint_as_float(av[c4 * 4 + 1]),
                                  __uint_as_float(av[c4 * 4 + 2]), __uint_as_float(av[c4 * 4 + 3]));
        }
      }
    }
  }

  // -------------------------------------------------------------------- teardown
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem, C::kTmemCols);
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
  static bool configured = false;
  if (!configured) {
    AVSSL_CUDA_OK(cudaFuncSetAttribute(infonce_tc_kernel<D, kThreeTerm>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)C::kSmemBytes));
    configured = true;
  }
  dim3 grid(p.n_splits, (p.B + kM - 1) / kM);
  infonce_tc_kernel<D, kThreeTerm><<<grid, kTcThreads, C::kSmemBytes, s>>>(p, cache.s, cache.v);
  AVSSL_LAUNCH_OK("infonce_tc_kernel");
  return AVSSL_OK;
}

}  // namespace

bool infonce_tc_supported(int B, int D, int K) {
  (void)B;
  return (D == 32 || D == 64 || D == 96 || D == 128) && K >= 1;
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
