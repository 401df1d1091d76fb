"""Multi-GPU check of the NVLink peer-memory key exchange (torchrun, N >= 2 GPUs of one box):
every step pushes fresh rows and compares what each rank reads back -- the whole gathered tensor
and an idx_restore-style row selection -- bit for bit with NCCL's all_gather, also through the
EMA-fused push and the InfoNCE-fused wait, eager and as a replayed CUDA graph."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
import torch.nn.functional as F
from advise_video_ssl_b200 import ops

world = int(os.environ["WORLD_SIZE"]); rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
B, D, K = 64, 128, 65536
x = ops.PeerExchange(B, D)
g = torch.Generator().manual_seed(100 + rank)
gp = torch.Generator().manual_seed(7)  # same permutation on every rank
bad = 0
for step in range(20):
    rows = F.normalize(torch.randn(B, D, generator=g)).to(dev)
    ref = torch.empty(world * B, D, device=dev)
    dist.all_gather_into_tensor(ref, rows)
    x.push(rows)
    got = x.wait_gather_all()
    perm = torch.randperm(world * B, generator=gp).view(world, B)[rank].to(dev)
    sel = x.wait_gather(perm)
    bad += int(not torch.equal(got, ref)) + int(not torch.equal(sel, ref[perm]))

# EMA-fused push + InfoNCE-fused wait against the NCCL path
online = [torch.randn(50000, device=dev), torch.randn(4096 * 7 + 5, device=dev)]
hist_a = [torch.randn_like(o) for o in online]; hist_b = [h.clone() for h in hist_a]
it_a = torch.ones(1, dtype=torch.int64, device=dev); it_b = it_a.clone()
plan_a, plan_b = ops.EmaPlan(online, hist_a), ops.EmaPlan(online, hist_b)
queue_a = F.normalize(torch.randn(K, D, generator=torch.Generator().manual_seed(1))).to(dev); queue_b = queue_a.clone()
ptr_a = torch.zeros(1, dtype=torch.int64, device=dev); ptr_b = ptr_a.clone()
st = torch.zeros(1, dtype=torch.int32, device=dev)
for step in range(6):
    f = torch.randn(B, D, generator=g).to(dev)
    rows = F.normalize(torch.randn(B, D, generator=g)).to(dev)
    perm = torch.randperm(world * B, generator=gp).view(world, B)[rank].to(dev)
    ref = torch.empty(world * B, D, device=dev)
    dist.all_gather_into_tensor(ref, rows)
    plan_a.run(0.99, it_a, bump_iter=True, first_iter=False)
    a = ops.moco_infonce(f, [ref[perm].contiguous()], queue_a, 0.1, enqueue=(ptr_a, st))
    plan_b.run(0.99, it_b, bump_iter=True, first_iter=False, push=(x, rows))
    b = ops.moco_infonce(f, None, queue_b, 0.1, enqueue=(ptr_b, st), peer=x, peer_row_idx=perm)
    for n in ("loss", "dfeat", "logits", "lse"):
        bad += int(not torch.equal(a[n], b[n]))
    bad += int(not torch.equal(queue_a, queue_b)) + int(not torch.equal(ptr_a, ptr_b))
    bad += sum(int(not torch.equal(p, q)) for p, q in zip(hist_a, hist_b))

# the module's _batch_unshuffle: peer path against the NCCL path (models/contrastive.py:216-230)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import make_cfg, register_backbones
C = register_backbones()
cfg = make_cfg(CONTRASTIVE__TYPE="moco", CONTRASTIVE__DIM=D, CONTRASTIVE__QUEUE_LEN=1024, NUM_GPUS=world)
model = C.ContrastiveModel(cfg).to(dev).train()
for step in range(4):
    y = F.normalize(torch.randn(B, D, generator=g)).to(dev)
    restore = torch.argsort(torch.randperm(world * B, generator=gp)).view(world, B).to(dev)
    model.enable_peer_exchange(False)
    ref = model._batch_unshuffle(y, restore)
    model.enable_peer_exchange(True)
    got = model._batch_unshuffle(y, restore)
    bad += int(not torch.equal(ref, got))
bad += int(model.check_device_status() != 0)

# the same step as a replayed graph (static inputs): results must keep matching the eager ones
f = torch.randn(B, D, generator=g).to(dev); rows = F.normalize(torch.randn(B, D, generator=g)).to(dev)
out = {}
ws = torch.zeros(ops.moco_infonce_workspace_bytes(B, D, K, 1), dtype=torch.uint8, device=dev)
def step_fn():
    plan_b.run(0.99, it_b, bump_iter=True, first_iter=False, push=(x, rows))
    r = ops.moco_infonce(f, None, queue_b, 0.1, enqueue=(ptr_b, st), peer=x, out=out, workspace=ws)
    if not out: out.update(r)
for _ in range(3): step_fn()
torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
gr = torch.cuda.CUDAGraph()
with torch.cuda.graph(gr): step_fn()
ev = lambda: torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
t0, t1 = ev(), ev(); t0.record()
for _ in range(200): gr.replay()
t1.record(); torch.cuda.synchronize()
us = t0.elapsed_time(t1) * 1e3 / 200
ref_loss = ops.moco_infonce(f, [rows], queue_a, 0.1)["loss"]  # queue_a is 203 enqueues behind: only finite-ness is checked
bad += int(not torch.isfinite(out["loss"]).item()) + int(st.item() != 0)
tot = torch.tensor([bad, int(us)], device=dev); dist.all_reduce(tot, op=dist.ReduceOp.MAX)
if rank == 0:
    print(json.dumps({"world": world, "mismatches": int(tot[0].item()), "graph_us_per_step_small_ema": float(tot[1].item())}), flush=True)
torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
os._exit(0 if int(tot[0].item()) == 0 else 1)
