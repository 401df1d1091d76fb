// K6 on the 5th-generation tensor cores: SimCLR NT-Xent for THIS rank's rows against all
// gathered columns (models/contrastive.py:770-792; see ntxent.cu for the math and the
// CUDA-core reference kernels, whose partial layout and finalisation kernels are reused).
//
// One CTA = 128 local rows x one range of columns.  Per 64-column tile of `out`:
//     S  = Q . tile^T                       (M=128 x N=64, K=D;   kind::f16, A = Q in TMEM, B = tile K-major)
//     e  = 2^((S - 1) log2e / T)            (softmax warps: tcgen05.ld -> exp2; diagonal masked)
//   pass 1 (rowsum):  z_r += sum_c e_rc
//   pass 2 (grad):    P = e (1/Z_r + 1/Z_c) -> TMEM;  acc += P . tile   (M=128 x N=D, K=64, B = tile MN-major)
//
// Operands are fp16 copies of the unit-length rows (avssl_ntxent_prepare writes them next to the exact
// fp32 rows): on [-1, 1] fp16 carries the same 11 significant bits as tf32, so the accuracy is that of
// round-to-nearest tf32 (measured ~1e-5 loss, ~3e-4 gradient, inside the 1e-3 fp32 tolerance), while
//   * one MMA covers K = 16 instead of 8: half the instructions for the same flops (the tf32 kernel was
//     bound by the ~50-cycle issue interval of its N = 64 MMAs, tensor pipe 22-31 % active),
//   * a tile is half the bytes, and with 16-bit elements the plain 128-byte swizzle serves BOTH operand
//     roles (K-major for S, MN-major for P.V): ONE shared-memory copy per tile instead of two,
//   * Q takes D/2 TMEM columns, so the whole D = 256 accumulator fits beside it (128 + 128 + 256 = 512):
//     pass 2 no longer recomputes S for a second half of D.
// P = e (1/Z_r + 1/Z_c) is the sum of two softmax probabilities, hence <= 2: it is scaled by 2^14 before
// the conversion to fp16 (<= 32768, no overflow; probabilities down to 4e-9 keep full precision) and the
// accumulator is scaled back in the epilogue.  The CUDA-core kernels (AVSSL_IMPL_SIMT) remain the exact-fp32
// reference and serve every D that is not a multiple of 64.
//
// Warp roles (352 threads, 1 CTA / SM): 0 TMA producer | 1 issuer of the S MMAs + TMEM allocator |
// 2-9 softmax + epilogue: thread = row, and the two warps that share a TMEM sub-partition (w, w+4)
// split every tile's columns (and the accumulator read-out) between them | 10 issuer of the P.V MMAs.
// Two issuing warps because ONE thread needs ~54 cycles per tcgen05.mma: with S (16) and P.V (4 long) MMAs of a
// tile issued by the same thread the gradient loop ran at ~1650 cycles per tile against 1024 tensor cycles.
#include <cuda_fp16.h>

#include "ntxent.cuh"
#include "sm100_ptx.cuh"
#include "tc_trace.cuh"

namespace avssl {

namespace {

// Columns of `out` per tile.  An N = 64 kind::f16 MMA costs ~54 cycles whatever it computes (the S MMAs of a 64-column
// tile take 870 cycles for 512 cycles of tensor work); N = 128 MMAs cost ~70 for twice the work, so the row-sum pass
// (TMEM: 3 x 128 columns of S) uses 128-column tiles.  The gradient pass was measured both ways (the code handles
// either): with 128-column tiles TMEM only holds a double buffer of S/P beside the D = 256 accumulator and shared
// memory two 64 KiB tile slots, and the longer softmax per tile (16 exp2 per clock and SM) serialises with the slot
// release: 14.4 us against 13.9 us with 64-column tiles, three S/P buffers and five slots.
constexpr int kBJGrad = 64;
constexpr int kBJSum = 128;
constexpr int kMt = 128;   // local rows per CTA
constexpr int kNtThreads = 352;
constexpr int kSoftmax = 256;
constexpr int kMaxSlots = 6;
constexpr int kHalfCols = 32;  // columns per tcgen05.ld chunk and per 1/Z staging row
constexpr float kLog2eT = 1.4426950408889634f;
constexpr float kPScale = 16384.f, kPUnscale = 1.f / 16384.f;

// UMMA instruction descriptor for kind::f16 with fp16 operands (a/b format 0) and fp32 accumulate
__host__ __device__ constexpr uint32_t umma_idesc_f16(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}

template <int D, bool kGrad>
struct NtCfg {
  static_assert(D % 64 == 0 && D <= 256, "whole 128-byte boxes of fp16");
  static constexpr int kBJ = kGrad ? kBJGrad : kBJSum;  // columns of `out` per tile
  static constexpr int kHalf = kBJ / 2;              // columns of a tile per softmax warp
  static constexpr int kKB = D / 64;                 // 128-byte boxes (64 fp16) per row
  static constexpr int kBoxBytes = kBJ * 128;        // one TMA box: kBJ rows x 128 B
  static constexpr int kTileBytes = kKB * kBoxBytes; // 32 KiB at D = 256
  // Q (the CTA's 128 rows, fp16) lives in SHARED memory as the K-major A operand, landed by TMA like the tiles:
  // no per-thread row loads and no TMEM stores in the prologue (round 2, first version: 128 strided 512-byte row
  // reads per CTA, ~2 us before the first MMA of both passes)
  static constexpr int kQBoxBytes = kMt * 128;          // one TMA box of Q: 128 rows x 128 B
  static constexpr int kQBytes = kKB * kQBoxBytes;      // 64 KiB at D = 256
  // S / P is TRIPLE-buffered in TMEM (3 x 64 columns): S(t) can be issued before P.V(t-2) has even been
  // requested, so the exponentials of tile t-1 have a whole S-issue period of slack (a double buffer forces
  // S(t+1) behind P.V(t-1), i.e. behind softmax(t-1): the chain showed as a 16 % active tensor pipe)
  static constexpr int kBufs = (3 * kBJ + (kGrad ? D : 0)) <= 512 ? 3 : 2;
  static constexpr int kColS = 0, kColAcc = kBufs * kBJ;
  static constexpr int kColsNeeded = kBufs * kBJ + (kGrad ? D : 0);
  static constexpr int kTmemCols = kColsNeeded <= 128 ? 128 : (kColsNeeded <= 256 ? 256 : 512);
  static_assert(kColsNeeded <= 512, "TMEM budget");
  static constexpr int kRingBudget = 232448 /* 227 KiB per CTA */ - 1024 - kQBytes - 8 * kHalf * 4 - 512;
  static constexpr int kSlots = (kRingBudget / kTileBytes) < kMaxSlots ? (kRingBudget / kTileBytes) : kMaxSlots;
  static_assert(kSlots >= 2, "at least a double buffer");
  static constexpr size_t kSmemBytes = 1024 + (size_t)kQBytes + (size_t)kSlots * kTileBytes + 8 * kHalf * 4 + 512;
};

struct NtBarriers {
  uint64_t s_full[kMaxSlots], s_free[kMaxSlots];  // tile landed / both MMAs that read it have retired
  uint64_t s_ready[3], p_ready[3], pv_done[3];
  uint64_t q_ready, acc_done;
  uint32_t tmem_base;
};

template <int D, bool kGrad>
__global__ void __launch_bounds__(kNtThreads, 1)
ntxent_tc_kernel(const NtxArgs a, const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap tmap_q) {
  using C = NtCfg<D, kGrad>;
  constexpr int kSlots = C::kSlots;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);  // keeps the shared address space
  uint8_t* q_smem = smem;
  uint8_t* ring = smem + C::kQBytes;
  float* invz_c = reinterpret_cast<float*>(ring + kSlots * C::kTileBytes);  // [8 softmax warps][kHalf]: 1/Z of the warp's columns
  NtBarriers* bar = reinterpret_cast<NtBarriers*>(ring + kSlots * C::kTileBytes + 8 * C::kHalf * 4);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) TC_TRACE(6, 0);
  const int split = blockIdx.x;
  const int i_base = blockIdx.y * kMt;
  const int j_begin = split * a.cols_per_split;
  const int j_end = min(a.N2, j_begin + a.cols_per_split);
  constexpr int kBJ = C::kBJ;
  const int n_tiles = (j_end - j_begin + kBJ - 1) / kBJ;
  // local rows [0, half) are global rows q_row0.., local rows [half, n_loc) global rows q_row1..: a CTA whose 128 rows
  // stay inside one of the two blocks gets Q by TMA; one that straddles them copies the rows itself (same layout)
  const int half = a.n_loc / 2;
  const bool q_by_tma = (i_base + kMt <= half) || (i_base >= half);

  if (warp == 0 && lane == 0) {
    ptx::tma_prefetch_desc(&tmap);
    ptx::tma_prefetch_desc(&tmap_q);
    for (int s = 0; s < kSlots; ++s) {
      ptx::mbar_init(&bar->s_full[s], 1);
      ptx::mbar_init(&bar->s_free[s], 1);
    }
    for (int b = 0; b < C::kBufs; ++b) {
      ptx::mbar_init(&bar->s_ready[b], 1);
      ptx::mbar_init(&bar->p_ready[b], kSoftmax);
      ptx::mbar_init(&bar->pv_done[b], 1);
    }
    ptx::mbar_init(&bar->q_ready, 1);
    ptx::mbar_init(&bar->acc_done, 1);
    ptx::mbar_fence_init();
  }
  if (warp == 1) ptx::tmem_alloc(&bar->tmem_base, C::kTmemCols);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = bar->tmem_base;
  if (tid == 0) TC_TRACE(6, 1);

  if (warp == 0) {
    // ================================================================ TMA producer
    if (lane == 0) {
      // Q: the CTA's 128 local rows are consecutive global rows (a.q_row0 + i_base ... within one of the two
      // B-row blocks; the host guarantees it), out-of-range rows of the last block are zero-filled by TMA
      if (q_by_tma) {
        ptx::mbar_arrive_expect_tx(&bar->q_ready, C::kQBytes);
        const int qrow = (i_base < half ? a.q_row0 : a.q_row1 - half) + i_base;
#pragma unroll
        for (int kb = 0; kb < C::kKB; ++kb) ptx::tma_load_2d(q_smem + kb * C::kQBoxBytes, &tmap_q, &bar->q_ready, kb * 64, qrow);
      }
      for (int t = 0; t < n_tiles; ++t) {
        const int sl = t % kSlots;
        if (t >= kSlots) ptx::mbar_wait(&bar->s_free[sl], ((t / kSlots) - 1) & 1);
        TC_TRACE(0, t);
        ptx::mbar_arrive_expect_tx(&bar->s_full[sl], C::kTileBytes);
        uint8_t* dst = ring + (size_t)sl * C::kTileBytes;
#pragma unroll
        for (int kb = 0; kb < C::kKB; ++kb)
          ptx::tma_load_2d(dst + kb * C::kBoxBytes, &tmap, &bar->s_full[sl], kb * 64, j_begin + t * kBJ);
      }
    }
  } else if (warp == 1) {
    // ================================================================== MMA issuer
    constexpr uint32_t idesc_s = umma_idesc_f16(kMt, kBJ, 0, 0);  // B = tile, K-major
    const uint32_t tm = __shfl_sync(0xffffffffu, tmem, 0);
    const uint32_t ring0 = __shfl_sync(0xffffffffu, ptx::smem_u32(ring), 0);
    const uint32_t q0 = __shfl_sync(0xffffffffu, ptx::smem_u32(q_smem), 0);
    ptx::mbar_wait(&bar->q_ready, 0);
    ptx::tc_fence_after();
    TC_TRACE(6, 2);
    const uint64_t qd0 = ptx::umma_smem_desc(q0, 16, 1024, ptx::kUmmaSwizzle128B);  // A = Q, K-major, 128-byte swizzle
    for (int t = 0; t < n_tiles; ++t) {
      const int sl = t % kSlots, b = t % C::kBufs;
      ptx::mbar_wait(&bar->s_full[sl], (t / kSlots) & 1);
      // S(t) overwrites the TMEM buffer of tile t-3: its exponentials must have been read (row sums), or the
      // P.V MMAs that read them as operand A must have retired (gradient; they are issued by another warp)
      if (t >= C::kBufs) {
        if (kGrad) ptx::mbar_wait(&bar->pv_done[b], ((t - C::kBufs) / C::kBufs) & 1);
        else ptx::mbar_wait(&bar->p_ready[b], ((t - C::kBufs) / C::kBufs) & 1);
      }
      ptx::tc_fence_after();
      TC_TRACE(1, t);
      // K-major: box ks/4 (64 fp16 = 128 B), 16 elements = 32 bytes per k-step inside the 128-byte row
      const uint64_t sd0 = ptx::umma_smem_desc(ring0 + sl * C::kTileBytes, 16, 1024, ptx::kUmmaSwizzle128B);
      const uint32_t d_s = tm + C::kColS + b * kBJ;
      if (ptx::elect_one()) {
#pragma unroll
        for (int ks = 0; ks < D / 16; ++ks)
          ptx::mma_f16_ss(d_s, qd0 + (uint64_t)(((ks >> 2) * C::kQBoxBytes + (ks & 3) * 32) >> 4),
                          sd0 + (uint64_t)(((ks >> 2) * C::kBoxBytes + (ks & 3) * 32) >> 4), idesc_s, ks > 0 ? 1u : 0u);
        ptx::tc_commit(&bar->s_ready[b]);
        if (!kGrad) ptx::tc_commit(&bar->s_free[sl]);
      }
      __syncwarp();
      TC_TRACE(2, t);
    }
  } else if (warp == 10) {
    // ======================================================= issuer of the P.V MMAs (gradient pass only)
    if (kGrad) {
      constexpr uint32_t idesc_pv = umma_idesc_f16(kMt, D, 0, 1);   // B = the same tile, MN-major
      const uint32_t tm = __shfl_sync(0xffffffffu, tmem, 0);
      const uint32_t ring0 = __shfl_sync(0xffffffffu, ptx::smem_u32(ring), 0);
      for (int t = 0; t < n_tiles; ++t) {
        const int sl = t % kSlots, b = t % C::kBufs;
        ptx::mbar_wait(&bar->p_ready[b], (t / C::kBufs) & 1);
        ptx::tc_fence_after();
        TC_TRACE(5, t);
        // MN-major, 128-byte swizzle, 16-bit elements: one k-step = 16 tile rows = two 8-row atoms 1024 B apart
        // (SBO); the 64-element column blocks of N (one TMA box each) are kBoxBytes apart (LBO)
        const uint64_t bd0 = ptx::umma_smem_desc(ring0 + sl * C::kTileBytes, C::kBoxBytes, 1024, ptx::kUmmaSwizzle128B);
        const uint32_t a0 = tm + C::kColS + b * kBJ;
        if (ptx::elect_one()) {
#pragma unroll
          for (int ks = 0; ks < kBJ / 16; ++ks)  // P of columns 32c..32c+31 sits packed in TMEM columns 32c..32c+15
            ptx::mma_f16_ts(tm + C::kColAcc, a0 + (ks >> 1) * kHalfCols + (ks & 1) * 8, bd0 + (uint64_t)(ks * 2048 >> 4),
                            idesc_pv, (t > 0 || ks > 0) ? 1u : 0u);
          ptx::tc_commit(&bar->s_free[sl]);   // S(t) retired before its exponentials were read: the slot is free
          ptx::tc_commit(&bar->pv_done[b]);   // the S/P buffer may be overwritten
          if (t == n_tiles - 1) ptx::tc_commit(&bar->acc_done);
        }
        __syncwarp();
      }
    }
  } else {
    // ================= softmax + epilogue: thread = local row; warp pair (w, w+4) shares a sub-partition
    const int sw = warp - 2;                    // 0..7
    const int sub = warp & 3;                   // TMEM sub-partition this warp may access
    const int hc = sw >> 2;                     // which 32 columns of every tile this warp handles
    const int r = sub * 32 + lane;
    const int i = i_base + r;
    const bool row_valid = i < a.n_loc;
    const uint32_t lane_base = tmem + ((uint32_t)(sub * 32) << 16);
    const int rid = row_valid ? __ldg(a.rows + i) : -1;  // global row id: the diagonal column of this row
    const float invz_r = (kGrad && row_valid) ? 1.f / __ldg(a.z_all + rid) : 0.f;
    float* iz = invz_c + sw * C::kHalf;         // this warp's staging of 1/Z_c (private: __syncwarp suffices)

    if (!q_by_tma) {
      // the CTA's rows straddle the two blocks: copy them into the K-major 128-byte-swizzled layout TMA would
      // have produced (16-byte chunk c of row r sits at chunk c ^ (r & 7) of its 128-byte line)
      constexpr int kChunksPerRow = D / 8;  // 16-byte chunks of 8 fp16
      const int st = tid - 64;              // 0..255
      for (int e = st; e < kMt * kChunksPerRow; e += kSoftmax) {
        const int rr = e / kChunksPerRow, c = e % kChunksPerRow;
        const int li = i_base + rr;
        uint4 x = make_uint4(0u, 0u, 0u, 0u);
        if (li < a.n_loc) x = __ldg(reinterpret_cast<const uint4*>(a.out_f16 + (size_t)__ldg(a.rows + li) * D) + c);
        *reinterpret_cast<uint4*>(q_smem + (c >> 3) * C::kQBoxBytes + rr * 128 + (((c & 7) ^ (rr & 7)) << 4)) = x;
      }
      ptx::fence_proxy_async_smem();
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (st == 0) ptx::mbar_arrive(&bar->q_ready);
    }
    const float scale2 = a.inv_T * kLog2eT;
    float zs[4] = {0.f, 0.f, 0.f, 0.f};
    // 1/Z of this warp's 32 columns of the NEXT tile is requested one tile ahead (one L2 round trip per tile otherwise)
    constexpr int kZ = C::kHalf / 32;  // values per lane
    float z_next[kZ];
#pragma unroll
    for (int u = 0; u < kZ; ++u) {
      const int j = j_begin + hc * C::kHalf + u * 32 + lane;
      z_next[u] = (kGrad && n_tiles > 0 && j < j_end) ? __ldg(a.z_all + j) : 1.f;
    }

    for (int t = 0; t < n_tiles; ++t) {
      const int b = t % C::kBufs;
      const int j0 = j_begin + t * kBJ + hc * C::kHalf;  // first column of this warp's half tile
      if (kGrad) {
        __syncwarp();
#pragma unroll
        for (int u = 0; u < kZ; ++u) {
          iz[u * 32 + lane] = (j0 + u * 32 + lane < j_end) ? 1.f / z_next[u] : 0.f;
          const int jn = j0 + kBJ + u * 32 + lane;
          z_next[u] = (t + 1 < n_tiles && jn < j_end) ? __ldg(a.z_all + jn) : 1.f;
        }
        __syncwarp();
      }
      ptx::mbar_wait(&bar->s_ready[b], (t / C::kBufs) & 1);
      ptx::tc_fence_after();
      if (sw == 0 && lane == 0) TC_TRACE(3, t);
#pragma unroll
      for (int ch = 0; ch < C::kHalf / kHalfCols; ++ch) {  // 32-column chunks of this warp's half tile (2 in the row-sum pass)
        const int jc = j0 + ch * kHalfCols;
        const uint32_t s_col = lane_base + C::kColS + b * kBJ + hc * C::kHalf + ch * kHalfCols;
        uint32_t sv[kHalfCols];
        ptx::tmem_ld32(s_col, sv);
        ptx::tc_wait_ld();
        const int diag = rid - jc;            // column of this chunk that is the row itself (masked), if in [0, 32)
        const int valid = row_valid ? min(kHalfCols, j_end - jc) : 0;
        uint32_t pk[kHalfCols / 2];
#pragma unroll
        for (int c = 0; c < kHalfCols; c += 2) {
          float e[2];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            // e^{(s - 1)/T}: unit rows give s <= 1, so the exponent is <= 0
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e[u]) : "f"((__uint_as_float(sv[c + u]) - 1.f) * scale2));
            e[u] = (c + u < valid && c + u != diag) ? e[u] : 0.f;
          }
          if (kGrad) {
            const float* izc = iz + ch * kHalfCols;
            const __half2 h = __floats2half2_rn(e[0] * (invz_r + izc[c]) * kPScale, e[1] * (invz_r + izc[c + 1]) * kPScale);
            pk[c >> 1] = *reinterpret_cast<const uint32_t*>(&h);  // low half = even k
          } else {
            zs[c & 3] += e[0];
            zs[(c + 1) & 3] += e[1];
          }
        }
        if (kGrad) {
          ptx::tmem_st16(s_col, pk);  // over the first 16 columns of this warp's own 32 (the pair never overlaps)
          ptx::tc_wait_st();
        }
      }
      ptx::tc_fence_before();
      if (sw == 0 && lane == 0) TC_TRACE(4, t);
      ptx::mbar_arrive(&bar->p_ready[b]);
    }
    if (kGrad) {
      // ---- the gradient partial is complete once the last PV retires; the pair splits the 32-column chunks
      ptx::mbar_wait(&bar->acc_done, 0);
      ptx::tc_fence_after();
      if (sw == 0 && lane == 0) TC_TRACE(6, 3);
      // The accumulator comes out of TMEM one ROW per thread; written like that, every store instruction of a warp
      // touches 32 different 1 KiB-strided rows (half-filled sectors: the epilogue took 9300 of the kernel's 26000
      // cycles).  Each warp therefore transposes its 32 x 32 chunk through a padded scratch in the (now idle) tile
      // ring and stores four whole 128-byte row segments per instruction.
      float* scratch = reinterpret_cast<float*>(ring) + sw * (32 * 33);
      const int g_row = lane >> 3, g_c4 = (lane & 7) * 4;  // store phase: lane -> (row within a group of 4, float4 column)
#pragma unroll 1
      for (int cb = hc; cb < D / 32; cb += 2) {
        uint32_t av[32];
        ptx::tmem_ld32(lane_base + C::kColAcc + cb * 32, av);
        ptx::tc_wait_ld();
        __syncwarp();
#pragma unroll
        for (int c = 0; c < 32; ++c) scratch[lane * 33 + c] = __uint_as_float(av[c]) * kPUnscale;  // conflict-free (pitch 33)
        __syncwarp();
#pragma unroll
        for (int r0 = 0; r0 < 32; r0 += 4) {
          const int rr = r0 + g_row;                 // row of this warp's sub-partition
          const int li = i_base + sub * 32 + rr;     // local row
          if (li < a.n_loc) {
            const float* sp = scratch + rr * 33 + g_c4;
            float4* dst = reinterpret_cast<float4*>(a.part_g + ((size_t)split * a.n_loc + li) * D + cb * 32 + g_c4);
            *dst = make_float4(sp[0], sp[1], sp[2], sp[3]);
          }
        }
      }
      ptx::tc_fence_before();
    } else {
      // the pair's two partial row sums meet in shared memory (the 1/Z staging area is unused in this pass)
      const float z = (zs[0] + zs[1]) + (zs[2] + zs[3]);
      float* zx = invz_c;  // [128] floats: one per row
      if (hc == 1) zx[r] = z;
      asm volatile("bar.sync 1, 256;" ::: "memory");  // the eight softmax warps
      if (hc == 0 && row_valid) a.part_z[(size_t)split * a.n_loc + i] = z + zx[r];
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (tid == 0) TC_TRACE(6, 4);
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem, C::kTmemCols);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn nt_encode_fn() {
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

struct NtTmapCache {
  const void* out = nullptr;
  int N2 = 0, D = 0;
  CUtensorMap s, q;
};

template <int D, bool kGrad>
int launch_nt(const NtxArgs& a, cudaStream_t st) {
  using C = NtCfg<D, kGrad>;
  static thread_local NtTmapCache cache;
  if (cache.out != a.out_f16 || cache.N2 != a.N2 || cache.D != D) {
    EncodeTiledFn enc = nt_encode_fn();
    AVSSL_REQUIRE(enc, AVSSL_ERR_CUDA, "ntxent: cuTensorMapEncodeTiled is not available from the driver");
    const cuuint64_t gdim[2] = {(cuuint64_t)D, (cuuint64_t)a.N2};
    const cuuint64_t gstride[1] = {(cuuint64_t)D * sizeof(uint16_t)};
    const cuuint32_t box[2] = {64u, (cuuint32_t)C::kBJ};  // 64 fp16 = 128 bytes x kBJ rows
    const cuuint32_t estride[2] = {1u, 1u};
    CUresult r = enc(&cache.s, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<uint16_t*>(a.out_f16), gdim, gstride, box,
                     estride, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    AVSSL_REQUIRE(r == CUDA_SUCCESS, AVSSL_ERR_CUDA, "ntxent: cuTensorMapEncodeTiled failed (%d)", (int)r);
    const cuuint32_t qbox[2] = {64u, (cuuint32_t)kMt};  // the CTA's 128 rows, one box per 64 fp16
    r = enc(&cache.q, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<uint16_t*>(a.out_f16), gdim, gstride, qbox, estride,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    AVSSL_REQUIRE(r == CUDA_SUCCESS, AVSSL_ERR_CUDA, "ntxent: cuTensorMapEncodeTiled (Q) failed (%d)", (int)r);
    cache.out = a.out_f16;
    cache.N2 = a.N2;
    cache.D = D;
  }
  static unsigned long long configured = 0ull;  // device ordinals already set up
  if (first_use_on_device(configured)) {
    AVSSL_CUDA_OK(cudaFuncSetAttribute(ntxent_tc_kernel<D, kGrad>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::kSmemBytes));
  }
  dim3 grid(a.n_splits, (a.n_loc + kMt - 1) / kMt);
  ntxent_tc_kernel<D, kGrad><<<grid, kNtThreads, C::kSmemBytes, st>>>(a, cache.s, cache.q);
  AVSSL_LAUNCH_OK("ntxent_tc_kernel");
  return AVSSL_OK;
}

}  // namespace

bool ntxent_tc_supported(int N2, int D, int n_loc) {
  (void)n_loc;
  return (D == 64 || D == 128 || D == 256) && N2 >= 2;
}

// Column split of the tcgen05 kernels: whole 64-column tiles, at most 64 splits (workspace layout),
// about one CTA per SM.
int ntxent_tc_plan(int N2, int n_loc, bool grad, int* n_splits, int* cols_per_split) {
  const int sms = sm_count();
  if (sms <= 0) return -1;
  const int kBJ = grad ? kBJGrad : kBJSum;
  const int row_blocks = (n_loc + kMt - 1) / kMt;
  const int n_tiles = (N2 + kBJ - 1) / kBJ;
  int S = sms / row_blocks;
  if (S < 1) S = 1;
  if (S > n_tiles) S = n_tiles;
  if (S > 64) S = 64;
  const int tps = (n_tiles + S - 1) / S;
  *n_splits = (n_tiles + tps - 1) / tps;
  *cols_per_split = tps * kBJ;
  return 0;
}

int launch_ntxent_tc(const NtxArgs& a, bool grad, cudaStream_t s) {
#define AVSSL_NT_CASE(DD) \
  case DD:                \
    return grad ? launch_nt<DD, true>(a, s) : launch_nt<DD, false>(a, s);
  switch (a.D) {
    AVSSL_NT_CASE(64)
    AVSSL_NT_CASE(128)
    AVSSL_NT_CASE(256)
  }
#undef AVSSL_NT_CASE
  return AVSSL_ERR_UNSUPPORTED;
}

}  // namespace avssl
