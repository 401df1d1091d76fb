"""Developer probe (GPU box): CUDA-event timing of each C-ABI call of the MoCo head step
(cfg2 sizes), L2 flushed before every timed call.  Prints medians."""
import json, os, sys, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from advise_video_ssl_b200 import ops, _lib

shapes = [tuple(s) for _, s in json.load(open(os.path.join(ROOT, "tests/golden/slow_r50_param_shapes.json")))["slow_r50_moco_dim128"]["shapes"]]
dev = torch.device("cuda")
online = [torch.randn(s, device=dev) * 0.02 for s in shapes]
hist = [torch.zeros_like(o) for o in online]
plan = ops.EmaPlan(online, hist)
it = torch.ones(1, dtype=torch.int64, device=dev)
ptr = torch.zeros(1, dtype=torch.int64, device=dev)
feat = torch.randn(64, 128, device=dev)
key = torch.nn.functional.normalize(torch.randn(64, 128, device=dev))
queue = torch.nn.functional.normalize(torch.randn(65536, 128, device=dev))
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
out = {}

def timed(fn, n=30, do_flush=True):
    ts = []
    for _ in range(n):
        if do_flush:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return statistics.median(ts), min(ts)

print("ema            : median %.1f us  min %.1f us  (%.0f GB/s at median)" % (*timed(lambda: plan.run(0.999, it, True)), plan.algorithmic_bytes / timed(lambda: plan.run(0.999, it, True))[0] / 1e3))
for name, impl in (("simt", 1), ("tc3x", 2), ("tc1x", 3)):
    for want in (False, True):
        r = ops.moco_infonce(feat, [key], queue, 0.1, want, impl)
        o = {k: v for k, v in r.items() if v is not None}
        f = lambda: ops.moco_infonce(feat, [key], queue, 0.1, want, impl, out=o)
        print("infonce %-5s logits=%d: cold-L2 median %.1f us min %.1f | warm median %.1f us" % (name, want, *timed(f), timed(f, do_flush=False)[0]))
print("enqueue        : median %.1f us  min %.1f us" % timed(lambda: ops.queue_enqueue(queue, ptr, key)))
e = lambda: None
print("empty (event overhead): median %.1f us" % timed(e)[0])
