"""Measurements of the SURVEY.md §8(f) kernels (GPU box): projection tail (rank 2) and kNN evaluation top-k (rank 4),
each beside the stock sequence it replaces on the same GPU.  One JSON line per measurement: CUDA-event medians, an L2
flush (512 MB read pass) before every timed call.  Usage: python tools/next_bench.py [--quick]"""
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
QUICK = "--quick" in sys.argv


def timed(fn, n=20):
    """Cold: median over n single launches, each behind two read passes over 512 MB (an L2 flush long enough -- ~200 us
    -- for the host to have the launch queued before the GPU gets there)."""
    ts = []
    for _ in range(3):
        fn()
    for _ in range(n):
        flush.sum(dtype=torch.int64)
        flush.sum(dtype=torch.int64)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return statistics.median(ts)


def streamed(fn, n=50):
    """Warm: n launches back to back behind one long flush, per-launch average (launch overhead overlapped, inputs in L2)."""
    fn()
    flush.sum(dtype=torch.int64)
    flush.sum(dtype=torch.int64)
    flush.sum(dtype=torch.int64)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / n


ev_over = timed(lambda: None)


def line(name, ours, stock, **kw):
    """ours / stock: callables.  Cold = one launch behind an L2 flush; warm = 50 launches back to back."""
    out = {"what": name, "us": round(timed(ours) - ev_over, 2), "stock_us": round(timed(stock) - ev_over, 2),
           "us_back_to_back": round(streamed(ours), 2), "stock_us_back_to_back": round(streamed(stock), 2),
           "event_overhead_us": round(ev_over, 2)}
    out.update(kw)
    print(json.dumps(out), flush=True)


from advise_video_ssl_b200._lib import lib, check  # noqa: E402  (direct C-ABI calls: no allocation in the timed region)

ST = lambda: torch.cuda.current_stream().cuda_stream  # noqa: E731

# ------------------------------------------------------------------ projection tail (models/head_helper.py:52-58 + Normalize)
torch.backends.cuda.matmul.allow_tf32 = False
for B, Kin, Dout in ((64, 2048, 128), (128, 2048, 128), (512, 2048, 256)):
    x = torch.randn(B, Kin, device=dev).relu_()
    W = torch.randn(Dout, Kin, device=dev) / Kin ** 0.5
    b = torch.randn(Dout, device=dev) * 0.1
    G = torch.randn(B, Dout, device=dev)
    q, nrm = ops.linear_l2norm_fwd(x, W, b)
    y = torch.empty(B, Dout, device=dev)
    dy, dx, dW, db = torch.empty_like(G), torch.empty_like(x), torch.empty_like(W), torch.empty_like(b)

    def ours_fwd():
        check(lib.avssl_linear_l2norm_fwd(x.data_ptr(), W.data_ptr(), b.data_ptr(), B, Kin, Dout, 0.0, 1, q.data_ptr(),
                                          nrm.data_ptr(), ST()), "fwd")

    def stock_fwd():
        torch.addmm(b, x, W.t(), out=y)
        check(lib.avssl_l2norm_fwd(y.data_ptr(), B, Dout, 0.0, q.data_ptr(), nrm.data_ptr(), ST()), "l2norm")

    def ours_bwd():
        check(lib.avssl_linear_l2norm_bwd(x.data_ptr(), W.data_ptr(), q.data_ptr(), nrm.data_ptr(), G.data_ptr(), B, Kin, Dout,
                                          0.0, 1, dx.data_ptr(), dW.data_ptr(), db.data_ptr(), ST()), "bwd")

    def stock_bwd():
        check(lib.avssl_l2norm_bwd(q.data_ptr(), nrm.data_ptr(), G.data_ptr(), B, Dout, 0.0, dy.data_ptr(), ST()), "l2norm_bwd")
        torch.mm(dy, W, out=dx)
        torch.mm(dy.t(), x, out=dW)
        torch.sum(dy, 0, out=db)

    line("projection tail fwd B=%d Kin=%d Dout=%d" % (B, Kin, Dout), ours_fwd, stock_fwd,
         stock_path="cuBLAS fp32 addmm + avssl_l2norm_fwd", algorithmic_bytes=4 * (B * Kin + Dout * Kin + Dout + B * Dout + B),
         flops=2 * B * Kin * Dout)
    line("projection tail bwd B=%d Kin=%d Dout=%d" % (B, Kin, Dout), ours_bwd, stock_bwd,
         stock_path="avssl_l2norm_bwd + two cuBLAS fp32 GEMMs + column sum",
         algorithmic_bytes=4 * (2 * B * Kin + 2 * Dout * Kin + 2 * B * Dout + Dout), flops=4 * B * Kin * Dout,
         note="module dispatch: fused launch up to B = 128, library GEMMs above (autograd.LinearNormalize)")

# ------------------------------------------------------------------ eval_knn (models/contrastive.py:232-241)
for N, M, D, k in ((64, 239975, 128, 200),) + (() if QUICK else ((256, 239975, 128, 200), (64, 65536, 128, 200))):
    qn = torch.nn.functional.normalize(torch.randn(N, D, device=dev), dim=1)
    bank = torch.nn.functional.normalize(torch.randn(M, D, device=dev), dim=1)
    dist = qn @ bank.t()
    yd = torch.empty(N, k, device=dev)
    yi = torch.empty(N, k, dtype=torch.int64, device=dev)
    ws = torch.empty(int(lib.avssl_topk_rows_workspace_bytes(N, M, k)), dtype=torch.uint8, device=dev)

    def ours_topk():
        check(lib.avssl_topk_rows(dist.data_ptr(), M, N, M, k, None, 0, yd.data_ptr(), yi.data_ptr(), ws.data_ptr(), ws.numel(),
                                  ST()), "topk")

    line("eval_knn N=%d M=%d D=%d k=%d" % (N, M, D, k), lambda: ops.knn_similarity_topk(qn, bank, k),
         lambda: (qn @ bank.t()).topk(k, dim=1, largest=True, sorted=True),
         stock_path="torch matmul (fp32) + torch.topk", algorithmic_bytes=4 * (M * D + N * D) + 12 * N * k,
         note="ours = tcgen05 sweep writing the [N, M+1] similarities + two-pass exact top-k")
    line("  top-k only (similarities given) N=%d M=%d k=%d" % (N, M, k), ours_topk,
         lambda: torch.topk(dist, k, dim=1, largest=True, sorted=True), stock_path="torch.topk",
         algorithmic_bytes=4 * N * M + 12 * N * k)
    del dist, bank

# ------------------------------------------------------------------ K11 SwAV swapped-prediction CE, kernel variants (cfg5)
import numpy as np  # noqa: E402

Bc, crops, P = 256, 6, 3000
sets = [(torch.randn(crops * Bc, P, device=dev) * 0.1, torch.empty(crops * Bc, P, device=dev)) for _ in range(8)]  # 8 x 37 MB > L2
codes = torch.softmax(torch.randn(2, Bc, P, device=dev), -1)
pw = np.ascontiguousarray(ops.swav_pair_weights(crops, 2, Bc), dtype=np.float32)
loss = torch.empty(1, device=dev)
wsc = torch.zeros(int(lib.avssl_swav_ce_workspace_bytes(crops * Bc)), dtype=torch.uint8, device=dev)
nbytes = 4 * P * (crops * Bc * 2 + 2 * Bc)
state = {"i": 0}


def k11():
    sc, d = sets[state["i"] % len(sets)]
    state["i"] += 1
    check(lib.avssl_swav_ce_fwd_bwd(sc.data_ptr(), codes.data_ptr(), crops, 2, Bc, P, 0.1, pw.ctypes.data, loss.data_ptr(),
                                    d.data_ptr(), wsc.data_ptr(), wsc.numel(), ST()), "swav_ce")


res, res_b2b = {}, {}
for name in ("sample", "reg"):
    os.environ["AVSSL_SWAV_CE_KERNEL"] = name
    res[name] = timed(k11) - ev_over
    res_b2b[name] = streamed(k11, n=48)
os.environ.pop("AVSSL_SWAV_CE_KERNEL")
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
hbm = float(peaks.get("hbm_gbs", 6650.0))
print(json.dumps({"what": "K11 SwAV CE fwd+bwd P=3000, 6 crops x 256 (cfg5)", "algorithmic_bytes": nbytes,
                  "us_cold": {k: round(v, 2) for k, v in res.items()},
                  "us_back_to_back_rotating_8_sets": {k: round(v, 2) for k, v in res_b2b.items()},
                  "frac_of_hbm_cold": {k: round(nbytes / v / 1e3 / hbm, 3) for k, v in res.items()},
                  "frac_of_hbm_back_to_back": {k: round(nbytes / v / 1e3 / hbm, 3) for k, v in res_b2b.items()},
                  "peak_gbs": hbm}), flush=True)
