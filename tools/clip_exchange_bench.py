"""C1 at full size (torchrun, N GPUs of one box): the shuffle-BN exchange of one key clip per rank,
[64, 3, 8, 224, 224] fp32 = 308.3 MB (BASELINE configs[1]), three ways:

  reference   cat_all_gather(x)[perm[rank]]     (models/contrastive.py:186-207: N x the bytes, then a gather)
  nccl_a2a    ShufflePlan.shuffled over NCCL    (send-buffer gather, all_to_all_single, reorder)
  nvlink      ops.PeerScatter.exchange          (each row written once, straight into its final position)

All three must agree bit for bit.  CUDA events on the stream, max over ranks; one JSON line from rank 0."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from advise_video_ssl_b200 import ops  # noqa: E402
from advise_video_ssl_b200.shuffle import ShufflePlan  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
B = int(os.environ.get("CLIP_B", "64"))
shape = (B, 3, 8, 224, 224)
x = torch.randn(shape, generator=torch.Generator().manual_seed(rank)).to(dev)
row_bytes = x[0].numel() * 4
gperm = torch.Generator().manual_seed(99)
sc = ops.PeerScatter(B, row_bytes)
status = torch.zeros(1, dtype=torch.int32, device=dev)
REPS = 5


def timed(fn):
    out, ts = None, []
    for r in range(REPS + 1):
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = fn()
        b.record()
        torch.cuda.synchronize()
        if r > 0:
            ts.append(a.elapsed_time(b))
    t = torch.tensor([sorted(ts)[len(ts) // 2]], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return out, float(t.item())


perm = torch.randperm(world * B, generator=gperm)
plan = ShufflePlan(perm.numpy(), world, rank, B, dev)
_ = plan.restore  # upload the index arrays before timing
take = plan.take


def reference():
    big = torch.empty((world * B,) + shape[1:], device=dev)
    dist.all_gather_into_tensor(big, x)
    return big.index_select(0, take)


ref, t_ref = timed(reference)
a2a, t_a2a = timed(lambda: plan.shuffled(x, None, None))
nvl, t_nvl = timed(lambda: plan.shuffled(x, None, sc, status))
ok = bool(torch.equal(ref, a2a) and torch.equal(ref, nvl)) and int(status.item()) == 0
flag = torch.tensor([0 if ok else 1], device=dev)
dist.all_reduce(flag)
if rank == 0:
    mb = B * row_bytes / 1e6
    print(json.dumps({"what": "C1 key-clip shuffle exchange, %.1f MB per rank, %d ranks" % (mb, world), "bit_identical": int(flag.item()) == 0,
                      "reference_all_gather_then_select_ms": round(t_ref, 3), "nccl_all_to_all_ms": round(t_a2a, 3),
                      "nvlink_scatter_ms": round(t_nvl, 3),
                      "nvlink_scatter_GBps_per_rank": round(mb * (world - 1) / world / t_nvl, 1),
                      "speedup_vs_reference": round(t_ref / t_nvl, 2)}), flush=True)
torch.cuda.synchronize()
sc.close()
os._exit(0)
