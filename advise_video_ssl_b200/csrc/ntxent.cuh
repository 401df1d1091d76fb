// Shared declarations of the SimCLR NT-Xent kernels (K6): CUDA-core reference (ntxent.cu) and
// tcgen05 kernels (ntxent_tc.cu).  Both write the same per-split partials, which the
// finalisation kernels of ntxent.cu merge in a fixed order.
#pragma once
#include "common.cuh"

namespace avssl {

struct NtxArgs {
  const float* out;     // [N2, D] unit rows, global order
  const uint16_t* out_f16;  // the same rows as fp16 (operands of the tcgen05 kernels); may be null for SIMT
  const int* rows;      // [n_loc] global row ids handled here
  int q_row0, q_row1;   // tcgen05 kernels: rows[0 .. n_loc/2) = q_row0 + i, rows[n_loc/2 .. n_loc) = q_row1 + (i - n_loc/2)
  const float* z_all;   // [N2] (pass 2)
  int N2, D, n_loc;
  float inv_T;
  int n_splits, cols_per_split;
  float* part_z;        // [n_splits][n_loc]
  float* part_g;        // [n_splits][n_loc][D]
};

bool ntxent_tc_supported(int N2, int D, int n_loc);
int ntxent_tc_plan(int N2, int n_loc, bool grad, int* n_splits, int* cols_per_split);
int launch_ntxent_tc(const NtxArgs& a, bool grad, cudaStream_t s);

}  // namespace avssl
