"""Developer probe (torchrun, N GPUs): where the multi-GPU step time goes.

Runs the cfg2 head step as a replayed CUDA graph in several variants and prints the per-rank and
max-over-ranks microseconds per step of each:
  replica    no exchange at all (N independent copies of the 1-GPU step)
  push_only  keys pushed to the peers by the EMA launch, head reads its local keys (no waiting)
  peer       the product path: push fused into the EMA, wait fused into the head
  nccl       all_gather on a high-priority side stream next to the EMA
"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
from advise_video_ssl_b200 import ops

world = int(os.environ["WORLD_SIZE"]); rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
shapes = [tuple(s) for _, s in json.load(open(os.path.join(ROOT, "tests/golden/slow_r50_param_shapes.json")))["slow_r50_moco_dim128"]["shapes"]]
online = [torch.randn(s, device=dev) * 0.02 for s in shapes]; hist = [torch.zeros_like(o) for o in online]
plan = ops.EmaPlan(online, hist); it = torch.ones(1, dtype=torch.int64, device=dev)
B, D, K = 64, 128, 65536
k = torch.nn.functional.normalize(torch.randn(B, D, device=dev)); f = torch.randn(B, D, device=dev)
queue = torch.nn.functional.normalize(torch.randn(K, D, device=dev))
ptr = torch.zeros(1, dtype=torch.int64, device=dev); status = torch.zeros(1, dtype=torch.int32, device=dev)
gathered = torch.empty(world * B, D, device=dev)
ws = torch.zeros(ops.moco_infonce_workspace_bytes(B, D, K, 1), dtype=torch.uint8, device=dev)
x = ops.PeerExchange(B, D)
comm = torch.cuda.Stream(device=dev, priority=-1)
out = {}
STEPS = int(os.environ.get("PROBE_STEPS", "300"))


def head(keys=None, peer=None):
    r = ops.moco_infonce(f, keys, queue, 0.1, want_logits=True, out=out, enqueue=(ptr, status), workspace=ws, peer=peer)
    if not out:
        out.update(r)


def step(mode):
    main = torch.cuda.current_stream()
    if mode == "replica":
        plan.run(0.999, it, True, first_iter=False)
        head([k])
    elif mode == "push_only":
        plan.run(0.999, it, True, first_iter=False, push=(x, k))
        head([k])
    elif mode == "peer":
        plan.run(0.999, it, True, first_iter=False, push=(x, k))
        head(None, peer=x)
    elif mode == "nccl":
        comm.wait_stream(main)
        with torch.cuda.stream(comm):
            dist.all_gather_into_tensor(gathered, k)
        plan.run(0.999, it, True, first_iter=False)
        main.wait_stream(comm)
        head([gathered[rank * B:(rank + 1) * B]])


def sync_all():
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()


res = {"world": world, "steps": STEPS}
ev = lambda: torch.cuda.Event(enable_timing=True)
for mode in os.environ.get("PROBE_MODES", "replica,push_only,peer,nccl").split(","):
    for _ in range(5):
        step(mode)
    sync_all()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        step(mode)
    for _ in range(5):
        g.replay()
    sync_all()
    a, b = ev(), ev(); a.record()
    for _ in range(STEPS):
        g.replay()
    b.record(); torch.cuda.synchronize()
    us = torch.tensor([a.elapsed_time(b) * 1e3 / STEPS], device=dev)
    allus = torch.empty(world, device=dev)
    dist.all_gather_into_tensor(allus, us)
    res[mode] = {"max": round(float(allus.max()), 2), "per_rank": [round(float(v), 1) for v in allus.tolist()]}
    sync_all()
    del g
if rank == 0:
    print(json.dumps(res), flush=True)
sync_all()
os._exit(0)
