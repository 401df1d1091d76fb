"""Developer probe (torchrun, N GPUs): CUPTI timeline of one graph-replayed SimCLR module step (cfg3 shape: 512 clips per
GPU, dim 256), rank 0 prints.  Shows what the two all_gathers and the torch glue cost next to the NT-Xent kernels."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.nn as nn  # noqa: E402
import bench  # noqa: E402

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
dev = torch.device("cuda", local)
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
from advise_video_ssl_b200 import contrastive as C  # noqa: E402


class Identity(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.dummy = nn.Parameter(torch.zeros(4))

    def forward(self, x):
        return x[0] if isinstance(x, (list, tuple)) else x


C._MODEL_TYPES["identity_embed"] = Identity
cfg = bench.head_cfg(world)
cfg.MODEL.ARCH = "identity_embed"
cfg.CONTRASTIVE.TYPE, cfg.CONTRASTIVE.DIM, cfg.CONTRASTIVE.T, cfg.CONTRASTIVE.QUEUE_LEN = "simclr", 256, 0.1, 64
cfg.TRAIN.BATCH_SIZE = 512 * world
model = C.ContrastiveModel(cfg).to(dev).train()
f1 = torch.randn(512, 256, device=dev).requires_grad_(True)
f2 = torch.randn(512, 256, device=dev).requires_grad_(True)
index = torch.arange(512, device=dev)


def step():
    f1.grad, f2.grad = None, None
    _, loss = model([[f1], [f2]], index, None, 0.0)
    loss.backward()


for _ in range(5):
    step()
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    step()
for _ in range(3):
    g.replay()
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile  # noqa: E402

with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(4):
        g.replay()
    torch.cuda.synchronize()
if rank == 0:
    evs = sorted([e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA], key=lambda e: e.time_range.start)
    fin = [i for i, e in enumerate(evs) if "ntxent_finish" in e.name]
    lo, hi = fin[1], fin[2]
    # one step = from the end of a finish kernel's backward tail to the next: print everything between two finish kernels
    t0 = evs[lo].time_range.end
    print("%-64s %9s %9s" % ("kernel / copy (after the previous step's finish kernel)", "start_us", "dur_us"))
    for e in evs[lo + 1:hi + 1]:
        print("%-64s %9.2f %9.2f" % (e.name[:64], e.time_range.start - t0, e.time_range.end - e.time_range.start))
    print("step (finish end -> next finish end): %.2f us" % (evs[hi].time_range.end - t0))
sys.stdout.flush()
torch.cuda.synchronize()
os._exit(0)
