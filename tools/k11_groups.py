"""Developer probe (GPU box): K11 sample-major kernel at cfg5 with the crops of a sample split over 1 / 2 / 3 / 6 CTAs
(AVSSL_SWAV_CE_GROUPS), cold (L2 flushed) and back to back over 8 rotating buffer sets.  Prints one JSON line."""
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
from advise_video_ssl_b200 import ops  # noqa: E402
from advise_video_ssl_b200._lib import lib, check  # noqa: E402

dev = torch.device("cuda")
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
Bc, crops, P = 256, 6, 3000
sets = [(torch.randn(crops * Bc, P, device=dev) * 0.1, torch.empty(crops * Bc, P, device=dev)) for _ in range(8)]
codes = torch.softmax(torch.randn(2, Bc, P, device=dev), -1)
pw = np.ascontiguousarray(ops.swav_pair_weights(crops, 2, Bc), dtype=np.float32)
loss = torch.empty(1, device=dev)
ws = torch.zeros(int(lib.avssl_swav_ce_workspace_bytes(crops * Bc)), dtype=torch.uint8, device=dev)
state = {"i": 0}


def k11():
    sc, d = sets[state["i"] % len(sets)]
    state["i"] += 1
    check(lib.avssl_swav_ce_fwd_bwd(sc.data_ptr(), codes.data_ptr(), crops, 2, Bc, P, 0.1, pw.ctypes.data, loss.data_ptr(),
                                    d.data_ptr(), ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream), "swav_ce")


def cold(n=20):
    ts = []
    for _ in range(n):
        flush.sum(dtype=torch.int64)
        flush.sum(dtype=torch.int64)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        k11()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return statistics.median(ts)


def b2b(n=48):
    flush.sum(dtype=torch.int64)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        k11()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / n


out = {}
for g in (1, 2, 3, 6):
    os.environ["AVSSL_SWAV_CE_GROUPS"] = str(g)
    for _ in range(3):
        k11()
    out["groups=%d" % g] = {"cold_us_incl_event_overhead": round(cold(), 2), "back_to_back_us": round(b2b(), 2)}
print(json.dumps({"what": "K11 sample-major, crops of a sample over G CTAs (cfg5)", **out}))
