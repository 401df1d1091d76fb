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
    ts = []
    for _ in range(3):
        fn()
    for _ in range(n):
        flush.sum(dtype=torch.int64)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return statistics.median(ts)


ev_over = timed(lambda: None)


def line(name, ours, stock_us, **kw):
    out = {"what": name, "us": round(ours - ev_over, 2), "stock_us": round(stock_us - ev_over, 2),
           "event_overhead_us": round(ev_over, 2)}
    out.update(kw)
    print(json.dumps(out), flush=True)


# ------------------------------------------------------------------ projection tail (models/head_helper.py:52-58 + Normalize)
torch.backends.cuda.matmul.allow_tf32 = False
for B, Kin, Dout in ((64, 2048, 128), (512, 2048, 256)):
    x = torch.randn(B, Kin, device=dev).relu_()
    W = torch.randn(Dout, Kin, device=dev) / Kin ** 0.5
    b = torch.randn(Dout, device=dev) * 0.1
    G = torch.randn(B, Dout, device=dev)
    q, nrm = ops.linear_l2norm_fwd(x, W, b)

    def stock_fwd():
        y = torch.nn.functional.linear(x, W, b)
        return ops.l2norm_fwd(y)

    def stock_bwd():
        dy = ops.l2norm_bwd(q, nrm, G)
        return dy @ W, dy.t() @ x, dy.sum(0)

    nbytes = 4 * (B * Kin + Dout * Kin + Dout + B * Dout + B)
    line("projection tail fwd B=%d Kin=%d Dout=%d" % (B, Kin, Dout), timed(lambda: ops.linear_l2norm_fwd(x, W, b)),
         timed(stock_fwd), stock="F.linear (cuBLAS fp32) + avssl_l2norm_fwd", algorithmic_bytes=nbytes, flops=2 * B * Kin * Dout)
    line("projection tail bwd B=%d Kin=%d Dout=%d" % (B, Kin, Dout), timed(lambda: ops.linear_l2norm_bwd(x, W, q, nrm, G)),
         timed(stock_bwd), stock="avssl_l2norm_bwd + two cuBLAS GEMMs + column sum",
         algorithmic_bytes=4 * (2 * B * Kin + 2 * Dout * Kin + 2 * B * Dout + Dout), flops=4 * B * Kin * Dout)

# ------------------------------------------------------------------ eval_knn (models/contrastive.py:232-241)
for N, M, D, k in ((64, 239975, 128, 200),) + (() if QUICK else ((256, 239975, 128, 200), (64, 65536, 128, 200))):
    qn = torch.nn.functional.normalize(torch.randn(N, D, device=dev), dim=1)
    bank = torch.nn.functional.normalize(torch.randn(M, D, device=dev), dim=1)

    def stock():
        return (qn @ bank.t()).topk(k, dim=1, largest=True, sorted=True)

    dist = qn @ bank.t()
    line("eval_knn N=%d M=%d D=%d k=%d" % (N, M, D, k), timed(lambda: ops.knn_similarity_topk(qn, bank, k)), timed(stock),
         stock="torch matmul (fp32) + torch.topk", algorithmic_bytes=4 * (M * D + N * D) + 12 * N * k,
         note="ours = tcgen05 sweep writing the [N, M+1] similarities + two-pass exact top-k")
    line("  top-k only (similarities given) N=%d M=%d k=%d" % (N, M, k), timed(lambda: ops.topk_rows(dist, k)),
         timed(lambda: dist.topk(k, dim=1, largest=True, sorted=True)), stock="torch.topk", algorithmic_bytes=4 * N * M + 12 * N * k)
    del dist, bank
