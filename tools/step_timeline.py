"""Developer probe (GPU box): kernel-level timeline of ONE graph-replayed module step
(contrastive_forward + backward, cfg2 sizes) from CUPTI via torch.profiler: start offset, duration
and stream of every kernel / memcpy, so gaps and overlaps are visible.  Prints a table."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
dev = torch.device("cuda", local)
torch.cuda.set_device(local)
if world > 1:  # torchrun --nproc-per-node N tools/step_timeline.py : rank 0 prints its timeline
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
C = bench.register_stub_backbone()
cfg = bench.head_cfg(world)
if os.environ.get("SHUFFLE", "1") == "0":
    cfg.BN.NORM_TYPE, cfg.BN.NUM_SYNC_DEVICES = "sync_batchnorm", 1
model = C.ContrastiveModel(cfg).to(dev).train()
model.materialize_logits = os.environ.get("LOGITS", "1") == "1"
B, D = bench.B_PER_GPU, bench.DIM
xq = torch.randn(B, D, device=dev).requires_grad_(True)
xk = torch.randn(B, D, device=dev)
index = torch.arange(B, device=dev)
t_in = torch.zeros(B, 2, 1, device=dev)


def step():
    xq.grad = None
    _, _, loss, bwd = C.contrastive_forward(model, cfg, [[xq], [xk]], index, t_in, 0.0)
    if bwd:
        loss.backward()
    return loss


for _ in range(5):
    step()
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    step()
for _ in range(5):
    g.replay()
torch.cuda.synchronize()

from torch.profiler import ProfilerActivity, profile  # noqa: E402

with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(4):
        g.replay()
    torch.cuda.synchronize()
if rank != 0:
    torch.cuda.synchronize()
    os._exit(0)
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
ema = [i for i, e in enumerate(evs) if "ema_multi_tensor" in e.name]
if len(ema) < 3:
    print("no CUPTI kernel records (%d events)" % len(evs))
    sys.exit(0)
lo, hi = ema[1], ema[2]  # the second replayed step, EMA start to next EMA start
# include what runs beside the EMA before it (side-stream work forked ahead)
t0 = evs[lo].time_range.start
first = lo
while first > 0 and evs[first - 1].time_range.start > evs[ema[0]].time_range.end:
    first -= 1
print("%-58s %9s %9s %7s" % ("kernel / copy", "start_us", "dur_us", "stream"))
for e in evs[first:hi]:
    print("%-58s %9.2f %9.2f %7s" % (e.name[:58], (e.time_range.start - t0), (e.time_range.end - e.time_range.start),
                                     getattr(e, "stream", "?")))
print("step (EMA start -> next EMA start): %.2f us" % (evs[hi].time_range.start - t0))
if world > 1:
    sys.stdout.flush()
    os._exit(0)
