"""Developer probe (torchrun, N GPUs): where the multi-GPU step time goes."""
import json, os, sys, statistics
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
k = torch.nn.functional.normalize(torch.randn(64, 128, device=dev)); f = torch.randn(64, 128, device=dev)
queue = torch.nn.functional.normalize(torch.randn(65536, 128, device=dev))
ptr = torch.zeros(1, dtype=torch.int64, device=dev); status = torch.zeros(1, dtype=torch.int32, device=dev)
gathered = torch.empty(world * 64, 128, device=dev)
ws = torch.zeros(ops.moco_infonce_workspace_bytes(64, 128, 65536, 1), dtype=torch.uint8, device=dev)
out = {}
prio = int(os.environ.get("COMM_PRIO", "-1"))
comm = torch.cuda.Stream(device=dev, priority=prio)
ema_first = os.environ.get("EMA_FIRST", "1") == "1"
ev = lambda: torch.cuda.Event(enable_timing=True)
def step(rec=None):
    main = torch.cuda.current_stream()
    if rec: rec["t0"].record()
    if ema_first:
        plan.run(0.999, it, True, first_iter=False)
        if rec: rec["ema_end"].record()
    comm.wait_stream(main) if not ema_first else None
    with torch.cuda.stream(comm):
        if rec: rec["g0"].record()
        dist.all_gather_into_tensor(gathered, k)
        if rec: rec["g1"].record()
    if not ema_first:
        plan.run(0.999, it, True, first_iter=False)
        if rec: rec["ema_end"].record()
    main.wait_stream(comm)
    kk = gathered[rank * 64:(rank + 1) * 64]
    r = ops.moco_infonce(f, [kk], queue, 0.1, want_logits=True, out=out, enqueue=(ptr, status), workspace=ws)
    if not out: out.update(r)
    if rec: rec["end"].record()
for _ in range(10): step()
torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
tl = []
for _ in range(60):
    rec = {n: ev() for n in ("t0", "ema_end", "g0", "g1", "end")}
    step(rec)
    tl.append(rec)
torch.cuda.synchronize()
med = lambda a, b: statistics.median(r[a].elapsed_time(r[b]) * 1e3 for r in tl[10:])
res = {"world": world, "ema_first": ema_first, "comm_prio": prio, "ema_end": med("t0", "ema_end"), "gather_start": med("t0", "g0"), "gather_end": med("t0", "g1"),
       "step_end": med("t0", "end")}
# back-to-back steps (no per-step sync)
torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
a, b = ev(), ev(); a.record()
for _ in range(100): step()
b.record(); torch.cuda.synchronize()
res["eager_us_per_step"] = a.elapsed_time(b) * 10
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g): step()
for _ in range(3): g.replay()
torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
a.record()
for _ in range(100): g.replay()
b.record(); torch.cuda.synchronize()
res["graph_us_per_step"] = a.elapsed_time(b) * 10
if rank == 0: print(json.dumps(res), flush=True)
dist.barrier(); torch.cuda.synchronize()
os._exit(0)
