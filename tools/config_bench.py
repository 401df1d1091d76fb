"""Per-config measurements of the other BASELINE.json configs (GPU box): SimCLR NT-Xent (cfg3),
BYOL EMA + sim loss (cfg4), SwAV Sinkhorn + swapped-prediction CE (cfg5), memory bank (K5).

Prints one JSON line per measurement: CUDA-event medians with a 512 MB L2 flush before every timed
call, the algorithmic bytes / flops of SURVEY.md §8(d) and the resulting roofline fraction.
These are single-GPU kernel timings of the per-rank work; `bench.py` is the headline (cfg2)."""
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from advise_video_ssl_b200 import ops  # noqa: E402

dev = torch.device("cuda")
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
HBM = float(peaks.get("hbm_gbs", 6650.0))
TF = float(peaks.get("bf16_tflops", 1590.0))


# L2 flush before every timed call.  "zero" (what profiles/r1_config_bench.jsonl was measured with) leaves
# up to 126 MB of dirty lines that are written back DURING the timed kernel; "read" streams the buffer
# through L2 instead, so the timed kernel starts with a clean cache (fairer to read-only kernels).
FLUSH = os.environ.get("AVSSL_FLUSH", "zero")


def flush_l2():
    if FLUSH == "read":
        flush.sum(dtype=torch.int64)
    else:
        flush.zero_()


def timed(fn, n=30):
    ts = []
    for _ in range(n):
        flush_l2()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return statistics.median(ts), min(ts)


ev_over = timed(lambda: None)[0]


def report(name, fn, nbytes=None, flops=None, note=""):
    med, mn = timed(fn)
    us = max(med - ev_over, 1e-3)
    out = {"config": name, "us_median": round(med, 2), "us_min": round(mn, 2), "event_overhead_us": round(ev_over, 2),
           "us_net": round(us, 2), "note": note}
    if nbytes is not None:
        out.update({"bound": "hbm", "algorithmic_bytes": int(nbytes), "achieved_gbs": round(nbytes / us / 1e3, 1),
                    "peak_gbs": HBM, "frac": round(nbytes / us / 1e3 / HBM, 4)})
    if flops is not None:
        out.update({"flops": int(flops), "achieved_tflops": round(flops / us / 1e6, 2), "bf16_peak_tflops": TF})
    print(json.dumps(out), flush=True)


# ------------------------------------------------------------------ cfg4: BYOL EMA + sim loss
shapes = json.load(open(os.path.join(ROOT, "tests/golden/slow_r50_param_shapes.json")))
for tag in shapes:
    shp = [tuple(s) for _, s in shapes[tag]["shapes"]]
    online = [torch.randn(s, device=dev) * 0.02 for s in shp]
    hist = [torch.zeros_like(o) for o in online]
    plan = ops.EmaPlan(online, hist)
    it = torch.ones(1, dtype=torch.int64, device=dev)
    report("K1 EMA %s (%d tensors, %d params)" % (tag, len(shp), plan.n_params),
           lambda: plan.run(0.996, it, True, first_iter=False), nbytes=plan.algorithmic_bytes)
    del online, hist, plan
# ---------------------------------------------- get_grad_norm_ over the Slow-R50 gradient list (8f rank 4)
shp = [tuple(s) for _, s in shapes["slow_r50_moco_dim128"]["shapes"]]
grads = [torch.randn(s, device=dev) for s in shp]
nplan = ops.MultiTensorNorm(grads)
report("multi-tensor L2 norm (get_grad_norm_) slow_r50_moco_dim128 (%d tensors, %d elements)" % (len(shp), nplan.n_elems),
       lambda: nplan.run(), nbytes=nplan.algorithmic_bytes)
del grads, nplan
pred = torch.randn(64, 256, device=dev)
key = torch.nn.functional.normalize(torch.randn(64, 256, device=dev))
report("K7 BYOL sim_loss fwd+bwd B=64 D=256 (one pair)", lambda: ops.byol_simloss(pred, key, 1.0), nbytes=3 * 4 * 64 * 256,
       note="latency-bound by construction (196 KB)")

# ------------------------------------------------------------------ cfg3: SimCLR NT-Xent, one rank of 8
IMPL = int(os.environ.get("AVSSL_NTX_IMPL", "0"))  # 0 auto (tcgen05), 1 CUDA cores
for B, W in ((512, 1), (512, 8)):
    D, T = 256, 0.1
    N = B * W
    f1, f2 = torch.randn(B, D, device=dev), torch.randn(B, D, device=dev)
    if W == 1:
        report("K6 NT-Xent single rank: B=%d D=%d (2N=%d rows x 2N cols)" % (B, D, 2 * N),
               lambda: ops.ntxent(f1, f2, T, gather=False, impl=IMPL), flops=3 * 2 * (2 * B) * (2 * N) * D,
               note="rowsum pass + gradient pass (2 GEMM-shaped sweeps + PV)")
    else:
        # per-rank work of the 8-GPU config on one GPU: this rank's 2B rows against 2N gathered columns
        from advise_video_ssl_b200._lib import lib, check
        gathered = torch.nn.functional.normalize(torch.randn(W, 2, B, D, device=dev), dim=-1)  # what the all_gather delivers
        out = torch.empty(2 * N, D, device=dev)
        out_r = torch.empty(2 * N, D, device=dev, dtype=torch.float16)
        rows = torch.cat([torch.arange(0, B, dtype=torch.int32, device=dev),
                          torch.arange(N, N + B, dtype=torch.int32, device=dev)])
        n_loc = 2 * B
        ws = torch.zeros(lib.avssl_ntxent_workspace_bytes(2 * N, D, n_loc), dtype=torch.uint8, device=dev)
        z = torch.empty(n_loc, device=dev)
        zall = torch.ones(2 * N, device=dev)
        nrm = torch.ones(n_loc, device=dev)
        loss = torch.empty(1, device=dev)
        dfe = torch.empty(n_loc, D, device=dev)
        st = torch.cuda.current_stream().cuda_stream

        def rank_work():
            check(lib.avssl_ntxent_prepare(gathered.data_ptr(), W, B, D, out.data_ptr(), out_r.data_ptr(), st), "prepare")
            check(lib.avssl_ntxent_rowsum(out.data_ptr(), out_r.data_ptr(), rows.data_ptr(), 0, N, 2 * N, D, n_loc, T, z.data_ptr(),
                                          ws.data_ptr(), ws.numel(), IMPL, st), "rowsum")
            check(lib.avssl_ntxent_grad(out.data_ptr(), out_r.data_ptr(), rows.data_ptr(), 0, N, zall.data_ptr(), nrm.data_ptr(),
                                        2 * N, D, n_loc, T, float(W), loss.data_ptr(), dfe.data_ptr(), ws.data_ptr(),
                                        ws.numel(), IMPL, st), "grad")
        report("K6 NT-Xent per-rank work of cfg3: 2B=%d local rows x 2N=%d cols, D=%d" % (2 * B, 2 * N, D), rank_work,
               flops=3 * 2 * (2 * B) * (2 * N) * D,
               note="assemble [q_all;q2_all] + tf32 copy, rowsum, grad, finalise (collectives excluded)")

# ------------------------------------------------------------------ cfg5: SwAV
for P in (3000, 1000):
    Bc, crops = 256, 6
    scores = torch.randn(crops * Bc, P, device=dev) * 0.1
    sc2 = scores[:Bc].contiguous()
    report("K10 Sinkhorn P=%d B=%d 3 iters (one call)" % (P, Bc), lambda: ops.sinkhorn(sc2, 0.05, 3),
           nbytes=8 * P * Bc, note="algorithmic = read scores once + write codes once; latency-dominated")
    codes = torch.softmax(torch.randn(2, Bc, P, device=dev), -1)
    report("K11 SwAV swapped CE fwd+bwd P=%d, %d crops x %d" % (P, crops, Bc), lambda: ops.swav_ce(scores, codes, crops, Bc, 0.1),
           nbytes=4 * P * (crops * Bc * 2 + 2 * Bc))

# ------------------------------------------------------------------ K5: memory bank
L, D = 239975, 128
bank = torch.nn.functional.normalize(torch.randn(L, 1, D, device=dev), dim=-1)
n = 512
mem = torch.randn(n, D, device=dev)
ind = torch.randint(0, L, (n,), device=dev)
report("K5 membank update n=%d (8 ranks x 64) D=128, bank %d rows" % (n, L), lambda: ops.membank_update(bank, mem, ind, None, 1.0),
       nbytes=12 * n * D + 8 * n, note="latency-bound (786 KB)")
