"""Multi-rank GPU parity worker (one process per GPU, launched by torchrun / tests/test_gpu_multirank.py).

Every rank runs the module through its public API on cuda:LOCAL_RANK over NCCL and compares with the CPU
oracle's single-process emulation of the reference's multi-GPU semantics:

  shuffle     _batch_shuffle / _batch_unshuffle with a random permutation  (models/contrastive.py:174-230)
  moco        3 training steps, shuffle BN on: per-rank loss / gradient, and C9 - the queue that the
              reference's DDP buffer broadcast (models/build.py:76-83) leaves on every rank - bit-identical
              across ranks, for queue_mode "reference", "canonical" and over both exchange transports
  simclr      row-sharded NT-Xent + AllGatherWithGradient against the single-process loss over the gathered
              batch, gradients with the x world factor                      (:770-792, utils/distributed.py:131-155)
  bank        Memory.update after the all_gather of (mem, ind, time)        (:989-1036, C8)

Rank 0 writes a JSON summary to --out; any mismatch makes every rank exit non-zero.
"""
import argparse
import json
import os
import sys
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def gather_cpu(t):
    """[world, ...] on the host."""
    out = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t.contiguous())
    return torch.stack([o.cpu() for o in out])


def check_shuffle(rank, world, dev, C, O, make_cfg, report):
    cfg = make_cfg(CONTRASTIVE__TYPE="moco", CONTRASTIVE__DIM=64, CONTRASTIVE__QUEUE_LEN=1024, NUM_GPUS=world)
    model = C.ContrastiveModel(cfg).to(dev).train()
    B = 48
    gen = torch.Generator().manual_seed(123)
    parts = [torch.randn(B, 3, 4, 5, generator=gen) for _ in range(world)]
    crops = [torch.randn(B, 7, generator=gen) for _ in range(world)]
    for trial, peer in enumerate((True, False)):
        model.enable_peer_exchange(peer)
        torch.manual_seed(77 + rank + 100 * trial)  # only rank 0's draw may matter (C2)
        (x, xc), restore = model._batch_shuffle([parts[rank].to(dev), crops[rank].to(dev)])
        torch.manual_seed(77 + 100 * trial)
        perm = torch.randperm(world * B)
        ref_x, ref_restore = O.shuffle_emulated(parts, perm)
        ref_c, _ = O.shuffle_emulated(crops, perm)
        assert torch.equal(x.cpu(), ref_x[rank]) and torch.equal(xc.cpu(), ref_c[rank]), "shuffled rows differ"
        assert torch.equal(restore.cpu(), ref_restore) and restore.dtype == torch.int64
        y = torch.randn(B, 64, generator=torch.Generator().manual_seed(5 + rank)).to(dev)
        ys = [torch.randn(B, 64, generator=torch.Generator().manual_seed(5 + r)) for r in range(world)]
        back = model._batch_unshuffle(y, restore)
        assert torch.equal(back.cpu(), O.unshuffle_emulated(ys, ref_restore)[rank]), "un-shuffled rows differ (peer=%s)" % peer
    assert model.check_device_status() == 0
    report["shuffle"] = {"rows_per_rank": B, "transports": ["nvlink_peer", "nccl"], "bit_exact": True}


def check_moco(rank, world, dev, C, O, make_cfg, report):
    from helpers import rel_err
    B, D, K, T, m = 32, 128, 2048, 0.1, 0.9
    out = {}
    # sweep: None = the single-launch head; 3 = the two-launch head (sweep on its own stream with 3 CTAs, the key
    # push over NVLink by the first CTAs of the finish launch)
    for mode, peer, sweep in (("reference", True, None), ("reference", False, None), ("canonical", True, None),
                              ("canonical", False, None), ("local", True, None), ("reference", True, 3),
                              ("canonical", True, 3), ("reference", False, 3), ("local", True, 0)):
        cfg = make_cfg(CONTRASTIVE__TYPE="moco", CONTRASTIVE__T=T, CONTRASTIVE__DIM=D, CONTRASTIVE__QUEUE_LEN=K,
                       CONTRASTIVE__MOMENTUM=m, NUM_GPUS=world)
        cfg.CONTRASTIVE.QUEUE_MODE = mode
        torch.manual_seed(0)  # identical replicas, as DDP would make them
        model = C.ContrastiveModel(cfg).to(dev).train()
        model.enable_peer_exchange(peer)
        model.sweep_ctas = sweep
        assert model._batch_shuffle_on
        W_on = model.backbone.proj.weight.detach().cpu().clone()
        W_hi = model.backbone_hist.proj.weight.detach().cpu().clone()
        queue = model.queue_x.cpu().clone()
        ptr = 0
        worst = {"loss": 0.0, "grad": 0.0}
        for step in range(3):
            gen = torch.Generator().manual_seed(1000 * step + 17)
            xq = [torch.randn(B, D, generator=gen) for _ in range(world)]
            xk = [torch.randn(B, D, generator=gen) for _ in range(world)]
            torch.manual_seed(50 + step + rank)
            model.zero_grad()
            logits, loss = model([[xq[rank].to(dev)], [xk[rank].to(dev)]], torch.arange(B, device=dev),
                                 torch.zeros(B, 2, 1, device=dev), 0.0)
            loss.backward()
            # ---- emulation of the reference on W ranks (shuffle BN does not change the keys of a BN-free stub)
            W_hi = O.ema_update([W_on], [W_hi], m, step)[0]
            keys = [O.l2_normalize(x @ W_hi.t()) for x in xk]
            f = (xq[rank] @ W_on.t()).requires_grad_(True)
            _, ref_logits, ref_loss = O.moco_head(f, [keys[rank]], queue, T)
            ref_loss.backward()
            worst["loss"] = max(worst["loss"], rel_err(loss.detach(), ref_loss.detach()))
            gW = model.backbone.proj.weight.grad.cpu()
            worst["grad"] = max(worst["grad"], rel_err(gW, f.grad.t() @ xq[rank]))
            assert (logits.cpu() - ref_logits.detach()).abs().max().item() < 3e-5
            if mode == "reference":    # every rank enqueues its own keys, then rank 0's buffers overwrite all others
                ptr = O.enqueue(queue, ptr, [keys[0]], K)
            elif mode == "canonical":  # all W*B keys in rank order
                ptr = O.enqueue(queue, ptr, [torch.cat(keys)], K)
            else:                      # the raw per-rank behaviour (queues diverge without DDP's broadcast)
                ptr = O.enqueue(queue, ptr, [keys[rank]], K)
            qs = gather_cpu(model.queue_x)
            ps = gather_cpu(model.ptr)
            if mode != "local":
                for r in range(1, world):
                    assert torch.equal(qs[r], qs[0]), "queues differ between ranks (%s, step %d)" % (mode, step)
            assert all(int(p) == ptr for p in ps.view(-1)), "ptr %s != %d" % (ps.view(-1).tolist(), ptr)
            # equal to the emulation: untouched rows bit for bit, freshly normalised rows to 1 ulp
            assert (qs[rank] - queue).abs().max().item() < 2e-7
        assert worst["loss"] < 2e-5 and worst["grad"] < 5e-4, worst
        assert model.check_device_status() == 0
        out["%s/%s%s" % (mode, "peer" if peer else "nccl", "" if sweep is None else "/two_launch")] = dict(
            worst, queues_identical=mode != "local", ptr=ptr)
        model.enable_peer_exchange(False)
    report["moco"] = out


def check_simclr(rank, world, dev, C, O, make_cfg, report):
    from helpers import rel_err
    from advise_video_ssl_b200 import _lib
    out = {}
    for B, D, T in ((64, 128, 0.1), (96, 256, 0.2)):
        cfg = make_cfg(CONTRASTIVE__TYPE="simclr", CONTRASTIVE__T=T, CONTRASTIVE__DIM=D, CONTRASTIVE__QUEUE_LEN=64,
                       TRAIN__BATCH_SIZE=B * world, NUM_GPUS=world, MODEL__ARCH="identity")
        model = C.ContrastiveModel(cfg).to(dev).train()
        gen = torch.Generator().manual_seed(9)
        f1 = [torch.randn(B, D, generator=gen) for _ in range(world)]
        f2 = [torch.randn(B, D, generator=gen) for _ in range(world)]
        # single-process reference: every rank evaluates the SAME loss over the gathered batch; the gradient that
        # reaches a rank's local rows is the sum over ranks of d loss / d rows = world x (utils/distributed.py:142-155)
        a = torch.cat(f1).requires_grad_(True)
        b = torch.cat(f2).requires_grad_(True)
        ref = O.ntxent(O.l2_normalize(a), O.l2_normalize(b), T)
        ref.backward()
        for peer in (True, False):  # both gathers over NVLink peer stores / over NCCL
            model.enable_peer_exchange(peer)
            for impl, lt, gt in ((_lib.IMPL_SIMT, 1e-5, 1e-4), (_lib.IMPL_AUTO, 2e-4, 1e-3)):
                model.ntxent_impl = impl
                x1 = f1[rank].to(dev).requires_grad_(True)
                x2 = f2[rank].to(dev).requires_grad_(True)
                _, loss = model([[x1], [x2]], torch.arange(B, device=dev), None, 0.0)
                loss.backward()
                sl = slice(rank * B, (rank + 1) * B)
                e = (rel_err(loss.detach(), ref.detach()), rel_err(x1.grad, world * a.grad[sl]), rel_err(x2.grad, world * b.grad[sl]))
                assert e[0] < lt and e[1] < gt and e[2] < gt, (B, D, impl, peer, e)
                out["B%d_D%d_impl%d_%s" % (B, D, impl, "peer" if peer else "nccl")] = {"loss": e[0], "grad": max(e[1:])}
            assert bool(model._peer_xchgs) == peer
        assert model.check_device_status() == 0
        model.enable_peer_exchange(False)
    report["simclr"] = out


def check_bank(rank, world, dev, C, O, make_cfg, report):
    cfg = make_cfg(NUM_GPUS=world)
    L, D, n = 500, 64, 24
    torch.manual_seed(3)
    bank = C.Memory(L, 1, D, cfg).to(dev)
    ref = bank.memory.cpu().clone()
    gen = torch.Generator().manual_seed(44)
    mem = [torch.randn(n, D, generator=gen) for _ in range(world)]
    ind = [torch.randint(0, L, (n,), generator=gen) for _ in range(world)]  # duplicates across ranks: last wins, rank order
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    bank.update(mem[rank].to(dev), 0.5, ind[rank].to(dev), torch.zeros(n, dtype=torch.int64, device=dev), status=status)
    O.membank_update(ref, torch.cat(mem), 0.5, torch.cat(ind), torch.zeros(world * n, dtype=torch.int64))
    got = gather_cpu(bank.memory)
    for r in range(world):
        assert torch.equal(got[r], got[0]), "banks differ between ranks"
    assert (got[rank] - ref).abs().max().item() < 2e-7 and int(status.item()) == 0
    report["bank"] = {"rows": world * n, "identical_across_ranks": True}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--only", default="shuffle,moco,simclr,bank")
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from helpers import make_cfg, register_backbones
    from oracle import contrastive_oracle as O
    C = register_backbones()
    report, failed = {"world": world}, None
    try:
        for name in args.only.split(","):
            globals()["check_" + name](rank, world, dev, C, O, make_cfg, report)
            torch.cuda.synchronize()
            dist.barrier()
    except Exception:  # noqa: BLE001
        failed = traceback.format_exc()
        sys.stderr.write("[rank %d] %s\n" % (rank, failed))
    flag = torch.tensor([1 if failed else 0], device=dev)
    dist.all_reduce(flag)
    report["ok"] = int(flag.item()) == 0
    if rank == 0:
        line = json.dumps(report)
        print(line, flush=True)
        if args.out:
            with open(args.out, "w") as f:
                f.write(line + "\n")
    torch.cuda.synchronize()
    sys.stdout.flush()
    sys.stderr.flush()
    os._exit(0 if report["ok"] else 1)  # no NCCL teardown: peers may already be gone when a check failed


if __name__ == "__main__":
    main()
