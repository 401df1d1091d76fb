"""Generate the committed golden vectors by running the UNMODIFIED reference
(`/root/reference/models/contrastive.py`) on seeded synthetic inputs.

Run in the authoring container only (needs /root/reference):

    python tests/golden/make_golden.py

Outputs `tests/golden/*.npz`.  `tests/test_oracle_golden.py` pins `oracle/`
against them (no GPU); the `-m gpu` tests compare the CUDA path against both.
Recipes that tests re-create from seeds (large tensors are not committed) live
in `tests/golden/recipes.py` and are imported here, so generator and tests share
one definition.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402
import recipes  # noqa: E402

rc = ref_shim.load_reference()
ref_shim.register_stub_backbone(rc)
torch.set_num_threads(1)  # deterministic reduction order for the committed vectors


def _np(t):
    return t.detach().cpu().numpy().copy()


class Tap:
    """Records backbone outputs (with grads) of the reference model."""

    def __init__(self, module):
        self.outs = []
        module.register_forward_hook(self._hook)

    def _hook(self, mod, inp, out):
        lst = out if isinstance(out, list) else [out]
        for o in lst:
            if o.requires_grad:
                o.retain_grad()
        self.outs.append(lst)

    def pop(self):
        o, self.outs = self.outs, []
        return o


def save(name, **arrs):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrs)
    print("wrote", path, "%.1f KB" % (os.path.getsize(path) / 1024))


# ------------------------------------------------------------------ MoCo (small, 3 steps)
def gen_moco_small(shuffle_on):
    B, D, K, T, m, steps = 16, 128, 1024, 0.1, 0.99, 3
    cfg = ref_shim.make_cfg(CONTRASTIVE__TYPE="moco", CONTRASTIVE__T=T, CONTRASTIVE__DIM=D,
                            CONTRASTIVE__QUEUE_LEN=K, CONTRASTIVE__MOMENTUM=m)
    torch.manual_seed(11)
    model = rc.ContrastiveModel(cfg).train()
    model._batch_shuffle_on = shuffle_on
    tap = Tap(model.backbone)
    out = {"B": B, "D": D, "K": K, "T": T, "m": m, "steps": steps,
           "queue0": _np(model.queue_x), "W0": _np(model.backbone.proj.weight),
           "Whist0": _np(model.backbone_hist.proj.weight)}
    g = torch.Generator().manual_seed(5)
    for s in range(steps):
        xq = torch.randn(B, D, generator=g)
        xk = torch.randn(B, D, generator=g)
        idx = torch.arange(B)
        t = torch.zeros(B, 2, 1)
        model.zero_grad()
        W_before = _np(model.backbone.proj.weight)
        logits, loss = model([[xq], [xk]], idx, t, 0.0)
        loss.backward()
        (fq,) = tap.pop()[0]
        out["xq%d" % s] = _np(xq)
        out["xk%d" % s] = _np(xk)
        out["W%d" % s] = W_before
        out["featq%d" % s] = _np(fq)
        out["dfeatq%d" % s] = _np(fq.grad)
        out["logits%d" % s] = _np(logits)
        out["loss%d" % s] = _np(loss)
        out["dW%d" % s] = _np(model.backbone.proj.weight.grad)
        out["Whist_after%d" % s] = _np(model.backbone_hist.proj.weight)
        out["queue_after%d" % s] = _np(model.queue_x)
        out["ptr_after%d" % s] = _np(model.ptr)
        out["iter_after%d" % s] = _np(model.iter)
        with torch.no_grad():
            model.backbone.proj.weight -= 0.5 * model.backbone.proj.weight.grad
    save("moco_small_shuffle" if shuffle_on else "moco_small", **out)


# --------------------------------------------------------------- MoCo cfg1 (full size)
def gen_moco_cfg1():
    r = recipes.moco_cfg1()
    B, D, K, T = r["B"], r["D"], r["K"], r["T"]
    cfg = ref_shim.make_cfg(CONTRASTIVE__TYPE="moco", CONTRASTIVE__T=T, CONTRASTIVE__DIM=D,
                            CONTRASTIVE__QUEUE_LEN=K, CONTRASTIVE__MOMENTUM=r["m"])
    torch.manual_seed(0)
    model = rc.ContrastiveModel(cfg).train()
    model._batch_shuffle_on = False
    with torch.no_grad():
        model.queue_x.copy_(r["queue"])
        model.backbone.proj.weight.copy_(r["W"])
        model.backbone_hist.proj.weight.copy_(r["W"])  # keys = l2(xk @ W^T) after EMA(iter 0)
    tap = Tap(model.backbone)
    logits, loss = model([[r["xq"]], [r["xk"]]], torch.arange(B), torch.zeros(B, 2, 1), 0.0)
    loss.backward()
    (fq,) = tap.pop()[0]
    lg = logits.detach().double()
    save("moco_cfg1", loss=_np(loss), featq=_np(fq), dfeatq=_np(fq.grad),
         logits_head=_np(logits[:, :16]), logits_tail=_np(logits[:, -16:]),
         logits_rowsum=_np(lg.sum(1)), lse=_np(torch.logsumexp(lg, 1)),
         queue_rows_after=_np(model.queue_x[:B]), ptr_after=_np(model.ptr),
         Whist_after=_np(model.backbone_hist.proj.weight),
         queue_checksum=np.float64(r["queue"].double().sum().item()))


# ------------------------------------------------- MoCo multi-key + multi-view queue wrap
def gen_moco_multikey():
    B, D, K, T, m, steps = 8, 64, 64, 0.07, 0.9, 5
    cfg = ref_shim.make_cfg(CONTRASTIVE__TYPE="moco", CONTRASTIVE__T=T, CONTRASTIVE__DIM=D,
                            CONTRASTIVE__QUEUE_LEN=K, CONTRASTIVE__MOMENTUM=m,
                            CONTRASTIVE__MOCO_MULTI_VIEW_QUEUE=True)
    torch.manual_seed(3)
    model = rc.ContrastiveModel(cfg).train()
    model._batch_shuffle_on = False
    tap = Tap(model.backbone)
    out = {"B": B, "D": D, "K": K, "T": T, "m": m, "steps": steps,
           "queue0": _np(model.queue_x), "W0": _np(model.backbone.proj.weight)}
    g = torch.Generator().manual_seed(6)
    for s in range(steps):
        xs = [torch.randn(B, D, generator=g) for _ in range(3)]
        model.zero_grad()
        logits, loss = model([[x] for x in xs], torch.arange(B), torch.zeros(B, 3, 1), 0.0)
        loss.backward()
        (fq,) = tap.pop()[0]
        for i, x in enumerate(xs):
            out["x%d_%d" % (s, i)] = _np(x)
        out["featq%d" % s] = _np(fq)
        out["dfeatq%d" % s] = _np(fq.grad)
        out["logits%d" % s] = _np(logits)
        out["loss%d" % s] = _np(loss)
        out["queue_after%d" % s] = _np(model.queue_x)
        out["ptr_after%d" % s] = _np(model.ptr)
        out["Whist_after%d" % s] = _np(model.backbone_hist.proj.weight)
    save("moco_multikey", **out)


# ------------------------------------------------------------------------------- BYOL
def gen_byol():
    B, D, T, m, steps = 16, 128, 0.5, 0.996, 2
    cfg = ref_shim.make_cfg(CONTRASTIVE__TYPE="byol", CONTRASTIVE__T=T, CONTRASTIVE__DIM=D,
                            CONTRASTIVE__QUEUE_LEN=256, CONTRASTIVE__MOMENTUM=m,
                            CONTRASTIVE__PREDICTOR_DEPTHS=[1])
    torch.manual_seed(21)
    model = rc.ContrastiveModel(cfg).train()
    tap = Tap(model.backbone)
    names = [n for n, _ in model.backbone.named_parameters()]
    out = {"B": B, "D": D, "T": T, "m": m, "steps": steps, "K": 256,
           "param_names": np.array(names)}
    for n, p in model.backbone.named_parameters():
        out["online0/" + n] = _np(p)
    for n, p in model.backbone_hist.named_parameters():
        out["hist0/" + n] = _np(p)
    g = torch.Generator().manual_seed(8)
    for s in range(steps):
        x1 = torch.randn(B, D, generator=g)
        x2 = torch.randn(B, D, generator=g)
        model.zero_grad()
        for n, p in model.backbone.named_parameters():
            out["online%d/%s" % (s, n)] = _np(p)
        logits, loss = model([[x1], [x2]], torch.arange(B), None, 0.0)
        loss.backward()
        o = tap.pop()
        (f1, p1), (f2, p2) = o[0], o[1]
        out["x1_%d" % s], out["x2_%d" % s] = _np(x1), _np(x2)
        out["pred1_%d" % s], out["pred2_%d" % s] = _np(p1), _np(p2)
        out["dpred1_%d" % s], out["dpred2_%d" % s] = _np(p1.grad), _np(p2.grad)
        out["loss%d" % s] = _np(loss)
        out["logits_shape%d" % s] = np.array(logits.shape)
        out["logits_col0_%d" % s] = _np(logits[:, 0])
        out["logits_abs_rest_sum%d" % s] = _np(logits[:, 1:].abs().sum())
        for n, p in model.backbone_hist.named_parameters():
            out["hist_after%d/%s" % (s, n)] = _np(p)
        with torch.no_grad():
            for p in model.backbone.parameters():
                p -= 0.1 * p.grad
    # keys of the last step, recomputed from the stored hist weights by the tests
    save("byol", **out)


# ----------------------------------------------------------------------------- SimCLR
def gen_simclr():
    B, D, T = 16, 64, 0.1
    cfg = ref_shim.make_cfg(CONTRASTIVE__TYPE="simclr", CONTRASTIVE__T=T, CONTRASTIVE__DIM=D,
                            CONTRASTIVE__QUEUE_LEN=32, TRAIN__BATCH_SIZE=B)
    torch.manual_seed(31)
    model = rc.ContrastiveModel(cfg).train()
    tap = Tap(model.backbone)
    g = torch.Generator().manual_seed(9)
    x1, x2 = torch.randn(B, D, generator=g), torch.randn(B, D, generator=g)
    logits, loss = model([[x1], [x2]], torch.arange(B), None, 0.0)
    loss.backward()
    o = tap.pop()
    (f1,), (f2,) = o[0], o[1]
    save("simclr", B=B, D=D, T=T, feat1=_np(f1), feat2=_np(f2), dfeat1=_np(f1.grad),
         dfeat2=_np(f2.grad), loss=_np(loss), logits_shape=np.array(logits.shape),
         logits_col0=_np(logits[:, 0]))


# ------------------------------------------------------------------------------- SwAV
def gen_swav():
    B, D, T, n_crops = 8, 128, 0.1, 4
    cfg = ref_shim.make_cfg(CONTRASTIVE__TYPE="swav", CONTRASTIVE__T=T, CONTRASTIVE__DIM=D,
                            CONTRASTIVE__QUEUE_LEN=64)
    torch.manual_seed(41)
    model = rc.ContrastiveModel(cfg).train()
    tap = Tap(model.backbone)
    g = torch.Generator().manual_seed(10)
    xs = [torch.randn(B, D, generator=g) for _ in range(n_crops)]
    W0 = _np(model.swav_prototypes.weight)
    logits, loss = model([[x] for x in xs], torch.arange(B), None, 0.0)
    loss.backward()
    o = tap.pop()
    out = dict(B=B, D=D, T=T, n_crops=n_crops, W0=W0, loss=_np(loss),
               W_after=_np(model.swav_prototypes.weight),
               dW=_np(model.swav_prototypes.weight.grad),
               logits_shape=np.array(logits.shape))
    for i, (f,) in enumerate(o):
        out["feat%d" % i] = _np(f)
        out["dfeat%d" % i] = _np(f.grad)
    save("swav", **out)


def gen_swav_queue():
    B, D, T, n_crops, L, steps = 8, 64, 0.1, 3, 24, 5
    cfg = ref_shim.make_cfg(CONTRASTIVE__TYPE="swav", CONTRASTIVE__T=T, CONTRASTIVE__DIM=D,
                            CONTRASTIVE__QUEUE_LEN=64, CONTRASTIVE__SWAV_QEUE_LEN=L)
    torch.manual_seed(42)
    model = rc.ContrastiveModel(cfg).train()
    tap = Tap(model.backbone)
    g = torch.Generator().manual_seed(12)
    out = dict(B=B, D=D, T=T, n_crops=n_crops, L=L, steps=steps,
               W0=_np(model.swav_prototypes.weight))
    for s in range(steps):
        xs = [torch.randn(B, D, generator=g) for _ in range(n_crops)]
        model.zero_grad()
        out["Wpre%d" % s] = _np(model.swav_prototypes.weight)
        logits, loss = model([[x] for x in xs], torch.arange(B), None, 15.0)
        loss.backward()
        o = tap.pop()
        for i, (f,) in enumerate(o):
            out["feat%d_%d" % (s, i)] = _np(f)
            out["dfeat%d_%d" % (s, i)] = _np(f.grad)
        out["loss%d" % s] = _np(loss)
        out["dW%d" % s] = _np(model.swav_prototypes.weight.grad)
        out["queue_after%d" % s] = _np(model.queue_swav)
        out["use_queue%d" % s] = np.array(bool(model.swav_use_the_queue))
    save("swav_queue", **out)


def gen_sinkhorn():
    torch.manual_seed(51)
    cfg = ref_shim.make_cfg(CONTRASTIVE__TYPE="swav", CONTRASTIVE__DIM=32, CONTRASTIVE__QUEUE_LEN=64)
    model = rc.ContrastiveModel(cfg)
    out = {}
    for name, (B, P) in {"a": (8, 1000), "b": (40, 300), "c": (5, 7)}.items():
        scores = torch.rand(B, P) * 2 - 1
        Q = torch.exp(scores / 0.05)
        out["scores_" + name] = _np(scores)
        out["code_" + name] = _np(model.sinkhorn(Q.clone(), 3))
    save("sinkhorn", **out)


# ------------------------------------------------------------------- memory banks / mem
def gen_membank():
    cfg = ref_shim.make_cfg()
    out = {}
    # Memory (2d), non-interp, duplicates, momentum 0.5 and 1.0 (kNN path)
    torch.manual_seed(61)
    for tag, mom in (("half", 0.5), ("one", 1.0)):
        mem = rc.Memory(40, 3, 32, cfg)
        bank0 = _np(mem.memory)
        upd = torch.randn(12, 32)
        ind = torch.tensor([3, 7, 3, 9, 11, 7, 0, 39, 3, 5, 6, 9])
        tim = torch.tensor([0, 1, 0, 2, 1, 1, 0, 2, 1, 0, 0, 2])
        mem.update(upd, momentum=mom, ind=ind, time=tim, interp=False)
        out.update({"m2d_%s_bank0" % tag: bank0, "m2d_%s_upd" % tag: _np(upd),
                    "m2d_%s_ind" % tag: _np(ind), "m2d_%s_time" % tag: _np(tim),
                    "m2d_%s_bank1" % tag: _np(mem.memory)})
    # Memory (2d) interp update + get
    mem = rc.Memory(20, 4, 16, cfg)
    bank0 = _np(mem.memory)
    upd = torch.randn(6, 16)
    ind = torch.tensor([1, 5, 9, 5, 0, 19])
    tim = torch.tensor([0.25, 1.5, 3.0, 2.75, 0.0, 1.0])
    got = mem.get(ind, tim, interp=True)
    mem.update(upd, momentum=0.7, ind=ind, time=tim, interp=True)
    out.update(mi_bank0=bank0, mi_upd=_np(upd), mi_ind=_np(ind), mi_time=_np(tim),
               mi_got=_np(got), mi_bank1=_np(mem.memory))
    # Memory1D
    mem = rc.Memory1D(30, 1, 24, cfg)
    bank0 = _np(mem.memory)
    upd = torch.randn(7, 24)
    ind = torch.tensor([2, 4, 29, 4, 0, 13, 2])
    mem.update(upd, momentum=0.3, ind=ind, time=torch.zeros_like(ind))
    out.update(m1d_bank0=bank0, m1d_upd=_np(upd), m1d_ind=_np(ind), m1d_bank1=_np(mem.memory))
    save("membank", **out)


def gen_mem_mode():
    B, D, K, L, T, m = 8, 32, 16, 50, 0.07, 0.5
    out = dict(B=B, D=D, K=K, L=L, T=T, m=m)
    for tag, mem_type, interp in (("1d", "1d", False), ("2di", "2d", True)):
        cfg = ref_shim.make_cfg(CONTRASTIVE__TYPE="mem", CONTRASTIVE__T=T, CONTRASTIVE__DIM=D,
                                CONTRASTIVE__QUEUE_LEN=K, CONTRASTIVE__LENGTH=L,
                                CONTRASTIVE__MOMENTUM=m, CONTRASTIVE__MEM_TYPE=mem_type,
                                CONTRASTIVE__INTERP_MEMORY=interp, CONTRASTIVE__KNN_ON=True)
        torch.manual_seed(71)
        model = rc.ContrastiveModel(cfg).train()
        tap = Tap(model.backbone)
        g = torch.Generator().manual_seed(13)
        x = torch.randn(B, D, generator=g)
        index = torch.tensor([4, 9, 4, 0, 49, 17, 23, 9])
        out[tag + "_bank0"] = _np(model.memory.memory)
        out[tag + "_knn0"] = _np(model.knn_mem.memory)
        torch.manual_seed(72)  # governs the negatives drawn inside forward (:390-397)
        prod, zero, flag = model([x], index, torch.zeros(B), 0.0)
        (fq,) = tap.pop()[0]
        out[tag + "_x"] = _np(x)
        out[tag + "_index"] = _np(index)
        out[tag + "_featq"] = _np(fq)
        out[tag + "_prod"] = _np(prod)
        out[tag + "_bank1"] = _np(model.memory.memory)
        out[tag + "_knn1"] = _np(model.knn_mem.memory)
        out[tag + "_ret"] = np.array([float(zero), float(flag)])
    save("mem_mode", **out)


# ---------------------------------------------------------------- EMA, many odd tensors
def gen_ema():
    import torch.nn as nn

    class Odd(nn.Module):
        def __init__(self, cfg):
            super().__init__()
            self.a = nn.Parameter(torch.randn(7, 3, 5))
            self.b = nn.Parameter(torch.randn(1))
            self.c = nn.Parameter(torch.randn(1031))
            self.d = nn.Parameter(torch.randn(64, 130))
            self.e = nn.Parameter(torch.randn(4096 + 3))
            self.bn = nn.BatchNorm1d(6)

        def forward(self, x):
            return x[0] if isinstance(x, (list, tuple)) else x

    rc._MODEL_TYPES["odd"] = Odd
    cfg = ref_shim.make_cfg(CONTRASTIVE__TYPE="moco", CONTRASTIVE__DIM=16, CONTRASTIVE__QUEUE_LEN=32,
                            CONTRASTIVE__MOMENTUM=0.9, CONTRASTIVE__MOMENTUM_ANNEALING=True,
                            MODEL__ARCH="odd", SOLVER__MAX_EPOCH=10)
    torch.manual_seed(81)
    model = rc.ContrastiveModel(cfg).train()
    names = [n for n, _ in model.backbone.named_parameters()]
    out = {"names": np.array(names), "m0": 0.9, "max_epoch": 10}
    for n, p in model.backbone_hist.named_parameters():
        out["hist_init/" + n] = _np(p)
    epochs = [0.0, 0.37, 4.2, 9.99]
    out["epochs"] = np.array(epochs)
    g = torch.Generator().manual_seed(14)
    for s, ep in enumerate(epochs):
        with torch.no_grad():
            for p in model.backbone.parameters():
                p.copy_(torch.randn(p.shape, generator=g))
        for n, p in model.backbone.named_parameters():
            out["online%d/%s" % (s, n)] = _np(p)
        model.momentum_anneal_cosine(ep)
        out["mmt%d" % s] = np.float64(model.mmt)
        model._update_history()
        model.iter += 1
        for n, p in model.backbone_hist.named_parameters():
            out["hist%d/%s" % (s, n)] = _np(p)
    save("ema", **out)


if __name__ == "__main__":
    gen_moco_small(False)
    gen_moco_small(True)
    gen_moco_cfg1()
    gen_moco_multikey()
    gen_byol()
    gen_simclr()
    gen_swav()
    gen_swav_queue()
    gen_sinkhorn()
    gen_membank()
    gen_mem_mode()
    gen_ema()
