"""Developer probe (GPU box): CUPTI timeline of the NT-Xent kernel sequence (per-rank work of cfg3) inside a CUDA graph."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from advise_video_ssl_b200._lib import lib, check  # noqa: E402

dev = torch.device("cuda")
B, W, D, T = 512, int(os.environ.get("K6_W", "8")), 256, 0.1
N = B * W
gathered = torch.nn.functional.normalize(torch.randn(W, 2, B, D, device=dev), dim=-1)
out = torch.empty(2 * N, D, device=dev)
out_h = torch.empty(2 * N, D, device=dev, dtype=torch.float16)
rows = torch.cat([torch.arange(0, B, dtype=torch.int32, device=dev), torch.arange(N, N + B, dtype=torch.int32, device=dev)])
n_loc = 2 * B
ws = torch.zeros(lib.avssl_ntxent_workspace_bytes(2 * N, D, n_loc), dtype=torch.uint8, device=dev)
z = torch.empty(n_loc, device=dev)
zall = torch.ones(2 * N, device=dev)
nrm = torch.ones(n_loc, device=dev)
loss = torch.empty(1, device=dev)
dfe = torch.empty(n_loc, D, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def seq():
    st = torch.cuda.current_stream().cuda_stream
    check(lib.avssl_ntxent_prepare(gathered.data_ptr(), W, B, D, out.data_ptr(), out_h.data_ptr(), st), "prepare")
    check(lib.avssl_ntxent_rowsum(out.data_ptr(), out_h.data_ptr(), rows.data_ptr(), 0, N, 2 * N, D, n_loc, T, z.data_ptr(), ws.data_ptr(), ws.numel(), 0, st), "rowsum")
    check(lib.avssl_ntxent_grad(out.data_ptr(), out_h.data_ptr(), rows.data_ptr(), 0, N, zall.data_ptr(), nrm.data_ptr(), 2 * N, D, n_loc, T, float(W), loss.data_ptr(), dfe.data_ptr(), ws.data_ptr(), ws.numel(), 0, st), "grad")


for _ in range(3):
    seq()
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    flush.zero_()
    seq()
g.replay()
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile  # noqa: E402

with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
evs = sorted([e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA], key=lambda e: e.time_range.start)
fl = [i for i, e in enumerate(evs) if "FillFunctor" in e.name]
lo = fl[-1]
t0 = evs[lo].time_range.end
print("%-60s %9s %9s" % ("kernel", "start_us", "dur_us"))
for e in evs[lo + 1:]:
    print("%-60s %9.2f %9.2f" % (e.name[:60], e.time_range.start - t0, e.time_range.end - e.time_range.start))
print("total after the flush: %.2f us" % (evs[-1].time_range.end - t0))
