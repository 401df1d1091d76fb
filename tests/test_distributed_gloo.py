"""CPU-only, world_size 2 over gloo: the collective helpers and the all-to-all
shuffle exchange (C1) against the oracle's emulation of the reference semantics
(models/contrastive.py:174-230, utils/distributed.py:109-155)."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, fn_name, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
        if p not in sys.path:
            sys.path.insert(0, p)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        globals()[fn_name](rank, world)
        ret[rank] = "ok"
    except Exception as e:  # surface the failure in the parent
        import traceback
        ret[rank] = traceback.format_exc()
    finally:
        dist.destroy_process_group()


def _run(fn_name, world=2):
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, fn_name, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert not p.is_alive(), "worker hung"
    for r in range(world):
        assert ret.get(r) == "ok", ret.get(r)


# ----------------------------------------------------------------------- worker bodies
def _body_helpers(rank, world):
    from advise_video_ssl_b200 import distributed as du
    assert du.get_rank() == rank and du.get_world_size() == world
    assert du.get_local_rank() == rank and du.get_local_size() == world
    x = torch.arange(6, dtype=torch.float32).view(3, 2) + 100 * rank
    g = du.cat_all_gather(x)
    assert torch.equal(g, torch.cat([torch.arange(6, dtype=torch.float32).view(3, 2) + 100 * r for r in range(world)]))
    a, b = du.all_gather([x, torch.tensor([rank, rank + 10])])
    assert torch.equal(a, g) and b.tolist() == [0, 10, 1, 11]
    t = torch.full((4,), float(rank + 1))
    du.all_reduce([t], average=True)
    assert torch.allclose(t, torch.full((4,), 1.5))
    t = torch.full((4,), float(rank + 1))
    du.all_reduce([t], average=False)
    assert torch.allclose(t, torch.full((4,), 3.0))


def _body_allgather_grad(rank, world):
    """AllGatherWithGradient backward == SUM over ranks of the full gradient, sliced
    (utils/distributed.py:142-155) -> local grad = world * true gradient when every
    rank computes the same loss."""
    from advise_video_ssl_b200 import distributed as du
    from oracle import contrastive_oracle as O
    torch.manual_seed(0)
    full = torch.randn(world * 3, 4)
    x = full[rank * 3:(rank + 1) * 3].clone().requires_grad_(True)
    g = du.AllGatherWithGradient.apply(x)
    assert torch.equal(g.detach(), full)
    w = torch.arange(g.numel(), dtype=torch.float32).view_as(g) * (rank + 1)  # rank-dependent upstream grad
    (g * w).sum().backward()
    per_rank = [torch.arange(g.numel(), dtype=torch.float32).view_as(g) * (r + 1) for r in range(world)]
    ref = O.allgather_with_gradient_bwd(per_rank, rank)
    assert torch.equal(x.grad, ref)


def _body_shuffle(rank, world):
    """_batch_shuffle / _batch_unshuffle of the module (all-to-all exchange) against
    the oracle's emulation of the reference's gather-then-select."""
    from helpers import make_cfg, register_backbones
    from oracle import contrastive_oracle as O
    C = register_backbones()
    cfg = make_cfg(CONTRASTIVE__TYPE="moco", CONTRASTIVE__DIM=8, CONTRASTIVE__QUEUE_LEN=32, NUM_GPUS=world)
    model = C.ContrastiveModel(cfg).train()
    B = 5
    gen = torch.Generator().manual_seed(123)
    parts = [torch.randn(B, 3, 2, generator=gen) for _ in range(world)]
    crops = [torch.randn(B, 7, generator=gen) for _ in range(world)]
    torch.manual_seed(77 + rank)  # only rank 0's draw may matter (broadcast src=0, :199-201)
    (x, xc), restore = model._batch_shuffle([parts[rank], crops[rank]])
    torch.manual_seed(77)
    perm = torch.randperm(world * B)
    ref_x, ref_restore = O.shuffle_emulated(parts, perm)
    ref_c, _ = O.shuffle_emulated(crops, perm)
    assert torch.equal(x, ref_x[rank]) and torch.equal(xc, ref_c[rank])
    assert torch.equal(restore, ref_restore) and restore.dtype == torch.int64
    # keys computed on the shuffled batch come back in the original order
    y = x.flatten(1).sum(1, keepdim=True)
    back = model._batch_unshuffle(y, restore)
    assert torch.equal(back, parts[rank].flatten(1).sum(1, keepdim=True))
    # across ranks the head stays one launch behind the key encoder (its sweep hides the key exchange):
    # the automatic sizing of a sweep beside the momentum update declines
    assert model.overlap_sweep and model._sweep_ctas_beside_ema(64, torch.device("cpu")) is None


def _body_shuffle_local_group(rank, world):
    """LOCAL_SHUFFLE_BN with a real per-node group smaller than WORLD (the reference's SLURM path):
    the permutation comes from GLOBAL rank 0 (broadcast over WORLD, :199-201) but rows are exchanged
    inside each node's group only; shuffle and un-shuffle must use the same group."""
    from helpers import make_cfg, register_backbones
    from advise_video_ssl_b200 import distributed as du
    from oracle import contrastive_oracle as O
    C = register_backbones()
    local = 2
    groups = [dist.new_group(list(range(n * local, (n + 1) * local))) for n in range(world // local)]
    du._LOCAL_PROCESS_GROUP = groups[rank // local]
    try:
        cfg = make_cfg(CONTRASTIVE__TYPE="moco", CONTRASTIVE__DIM=8, CONTRASTIVE__QUEUE_LEN=32, NUM_GPUS=local,
                       NUM_SHARDS=world // local)
        model = C.ContrastiveModel(cfg).train()
        assert du.get_local_size() == local and du.get_local_rank() == rank % local
        B = 6
        gen = torch.Generator().manual_seed(321)
        parts = [torch.randn(B, 4, generator=gen) for _ in range(world)]
        torch.manual_seed(5 + rank)
        (x,), restore = model._batch_shuffle([parts[rank]])
        torch.manual_seed(5)  # global rank 0's draw
        perm = torch.randperm(local * B)
        node = rank // local
        ref_x, ref_restore = O.shuffle_emulated(parts[node * local:(node + 1) * local], perm)
        assert torch.equal(x, ref_x[rank % local]) and torch.equal(restore, ref_restore)
        y = x * 2 + 1
        assert torch.equal(model._batch_unshuffle(y, restore), parts[rank] * 2 + 1)
    finally:
        du._LOCAL_PROCESS_GROUP = None


def _body_dist_sinkhorn(rank, world):
    """ops.sinkhorn_distributed is the multi-node cold path; it needs CUDA tensors, so on
    gloo/CPU only the oracle's rank emulation is checked for consistency with the
    single-rank result (the all_reduce structure of :889-910)."""
    from oracle import contrastive_oracle as O
    torch.manual_seed(5)
    Q = torch.exp((torch.rand(8, 12) * 2 - 1) / 0.05)  # [B, P]
    whole = O.sinkhorn(Q, 3)
    halves = O.distributed_sinkhorn_emulated([Q[:4].t().clone(), Q[4:].t().clone()], 3)
    assert torch.allclose(torch.cat(halves), whole, rtol=1e-4, atol=1e-7)


@pytest.mark.parametrize("fn", ["_body_helpers", "_body_allgather_grad", "_body_shuffle", "_body_dist_sinkhorn"])
def test_world2(fn):
    _run(fn)


def test_world4_local_groups_of_2():
    _run("_body_shuffle_local_group", world=4)
