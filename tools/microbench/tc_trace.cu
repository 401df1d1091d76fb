// Developer probe: per-role timeline of one CTA of the tcgen05 InfoNCE kernel.
#define AVSSL_TC_TRACE 1
#include <vector>
#include "../../advise_video_ssl_b200/csrc/core.cu"
#include "../../advise_video_ssl_b200/csrc/infonce_tc.cu"
using namespace avssl;
int main(int argc, char** argv) {
  const int three = argc > 1 ? atoi(argv[1]) : 1;
  const int want_logits = argc > 2 ? atoi(argv[2]) : 0;
  const int B = 64, D = 128, K = 65536;
  float *f, *q, *pm, *pl, *pa, *o;
  unsigned* cnt;
  cudaMalloc(&f, B * D * 4); cudaMalloc(&q, (size_t)K * D * 4);
  std::vector<float> h((size_t)K * D);
  for (size_t i = 0; i < h.size(); ++i) h[i] = ((i * 2654435761u) % 1000) / 1000.f - 0.5f;
  cudaMemcpy(q, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(f, h.data(), B * D * 4, cudaMemcpyHostToDevice);
  InfoNceParams p{};
  p.feat_q = f; p.queue = q; p.B = B; p.D = D; p.K = K; p.inv_T = 10.f; p.n_keys = 1;
  const int sms = 148, n_tiles = K / 64;
  int S = sms; int tps = (n_tiles + S - 1) / S; S = (n_tiles + tps - 1) / tps;
  p.n_splits = S; p.rows_per_split = tps * 64;
  cudaMalloc(&pm, S * B * 4); cudaMalloc(&pl, S * B * 4); cudaMalloc(&pa, (size_t)S * B * D * 4);
  p.part_m = pm; p.part_l = pl; p.part_acc = pa;
  cudaMalloc(&cnt, 256); cudaMemset(cnt, 0, 256); p.counter = cnt;
  cudaMalloc(&o, (size_t)B * (D * 2 + 8) * 4);
  if (want_logits) { float* lg; cudaMalloc(&lg, (size_t)B * (K + 1) * 4); p.logits_out = lg; }
  p.q_out = o; p.dfeat_out = o + B * D; p.row_loss = o + 2 * B * D; p.loss_out = o + 2 * B * D + B; p.keys[0] = f;
  { int pr = argc > 3 ? atoi(argv[3]) : 0; cudaMemcpyToSymbol(g_tc_probe, &pr, sizeof(int)); }
  for (int rep = 0; rep < 3; ++rep) {
    int rc = launch_infonce_tc(p, three, 0);
    cudaError_t e = cudaDeviceSynchronize();
    if (rc || e) { printf("rc %d %s %s\n", rc, avssl_last_error(), cudaGetErrorString(e)); return 1; }
  }
  long long tr[16][64];
  cudaMemcpyFromSymbol(tr, g_tc_trace, sizeof(tr));
  const char* names[] = {"tma_issue", "split_start", "split_done", "S_issue_start", "S_issued", "PV_start", "S_ready_seen", "P_done"};
  long long t0 = tr[0][0];
  printf("tiles per CTA %d, three_term %d (cycles relative to first TMA issue)\n", tps, three);
  for (int t = 0; t < tps; ++t) {
    printf("tile %d:", t);
    for (int ev = 0; ev < 8; ++ev) printf(" %s=%lld", names[ev], tr[ev][t] - t0);
    printf(" | pass1_done=%lld rescale_done=%lld | PV_mma_issued=%lld PV_committed=%lld | drain %lld..%lld\n", tr[8][t] - t0, tr[9][t] - t0, tr[13][1 + t] - t0, tr[12][1 + t] - t0, tr[11][1 + t] - t0, tr[11][32 + t] - t0);
  }
  printf("phases (cycles rel. first TMA issue): entry=%lld setup_done=%lld sweep_done=%lld barrier_passed=%lld rows_merged=%lld finish=%lld\n",
         tr[14][0] - t0, tr[14][1] - t0, tr[14][2] - t0, tr[14][3] - t0, tr[14][4] - t0, tr[14][5] - t0);
  {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    for (int rep = 0; rep < 50; ++rep) launch_infonce_tc(p, three, 0);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("50 back-to-back launches (queue L2-resident): %.2f us per launch\n", ms * 20.f);
  }
  printf("merge row 0: enter=%lld loads_issued=%lld max_done=%lld L_done=%lld acc_merged=%lld keys_done=%lld end=%lld\n", tr[10][0] - t0,
         tr[10][1] - t0, tr[10][2] - t0, tr[10][3] - t0, tr[10][4] - t0, tr[10][5] - t0, tr[10][6] - t0);
  printf("q prologue: enter=%lld q_full=%lld cb0=%lld cb1=%lld cb2=%lld cb3=%lld wait_st=%lld\n", tr[15][0] - t0, tr[15][1] - t0,
         tr[15][2] - t0, tr[15][3] - t0, tr[15][4] - t0, tr[15][5] - t0, tr[15][6] - t0);
  printf("prologue: normsA_done=%lld q_in_tmem=%lld mma_saw_q=%lld\n", tr[11][0] - t0, tr[12][0] - t0, tr[13][0] - t0);
  return 0;
}
