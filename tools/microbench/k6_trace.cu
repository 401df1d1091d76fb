// Developer probe: per-role timeline of CTA (0,0) of the tcgen05 NT-Xent kernels (cfg3 per-rank shape).
#define AVSSL_TC_TRACE 1
#include <vector>
#include <cuda_fp16.h>
#include "../../advise_video_ssl_b200/csrc/core.cu"
#include "../../advise_video_ssl_b200/csrc/ntxent_tc.cu"
using namespace avssl;
int main() {
  const int B = 512, W = 8, D = 256, N = B * W, N2 = 2 * N, n_loc = 2 * B;
  std::vector<__half> h((size_t)N2 * D);
  for (size_t i = 0; i < h.size(); ++i) h[i] = __float2half((((i * 2654435761u) % 1000) / 1000.f - 0.5f) * 0.12f);
  __half* out_h; cudaMalloc(&out_h, h.size() * 2); cudaMemcpy(out_h, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
  std::vector<int> rows(n_loc); for (int i = 0; i < B; ++i) { rows[i] = i; rows[B + i] = N + i; }
  int* drows; cudaMalloc(&drows, n_loc * 4); cudaMemcpy(drows, rows.data(), n_loc * 4, cudaMemcpyHostToDevice);
  std::vector<float> z(N2, 1.0f); float* dz; cudaMalloc(&dz, N2 * 4); cudaMemcpy(dz, z.data(), N2 * 4, cudaMemcpyHostToDevice);
  NtxArgs a{}; a.out = nullptr; a.out_f16 = reinterpret_cast<const uint16_t*>(out_h); a.rows = drows; a.q_row0 = 0; a.q_row1 = N;
  a.z_all = dz; a.N2 = N2; a.D = D; a.n_loc = n_loc; a.inv_T = 10.f;
  cudaMalloc(&a.part_z, (size_t)64 * n_loc * 4); cudaMalloc(&a.part_g, (size_t)64 * n_loc * D * 4);
  for (int grad = 0; grad < 2; ++grad) {
    ntxent_tc_plan(N2, n_loc, grad != 0, &a.n_splits, &a.cols_per_split);
    printf("splits %d cols/split %d tiles/CTA %d\n", a.n_splits, a.cols_per_split, a.cols_per_split / (grad ? 64 : 128));
    for (int rep = 0; rep < 3; ++rep) {
      int rc = launch_ntxent_tc(a, grad != 0, 0);
      cudaError_t e = cudaDeviceSynchronize();
      if (rc || e) { printf("rc %d %s %s\n", rc, avssl_last_error(), cudaGetErrorString(e)); return 1; }
    }
    long long tr[16][64]; cudaMemcpyFromSymbol(tr, g_tc_trace, sizeof(tr));
    const long long t0 = tr[6][0];
    printf("pass %d: entry=0 setup_done=%lld q_ready=%lld acc_done_seen=%lld teardown=%lld\n", grad, tr[6][1] - t0, tr[6][2] - t0,
           grad ? tr[6][3] - t0 : -1, tr[6][4] - t0);
    for (int t = 0; t < a.cols_per_split / (grad ? 64 : 128); ++t)
      printf("  tile %d: tma_issue=%lld S_issue=%lld S_issued=%lld S_seen=%lld P_done=%lld PV_issue=%lld\n", t, tr[0][t] - t0,
             tr[1][t] - t0, tr[2][t] - t0, tr[3][t] - t0, tr[4][t] - t0, grad ? tr[5][t] - t0 : -1);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventRecord(e0);
    for (int rep = 0; rep < 20; ++rep) launch_ntxent_tc(a, grad != 0, 0);
    cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("  20 back-to-back launches: %.2f us per launch\n", ms * 50.f);
  }
  return 0;
}
