"""Pins `oracle/contrastive_oracle.py` against the golden vectors produced by the
unmodified reference (tests/golden/make_golden.py).  CPU only.

Bit-exact (`torch.equal`) wherever the restatement performs the same torch ops in
the same order with one thread, which is nearly everywhere; a stated tolerance
otherwise.
"""
import math

import pytest
import torch
import torch.nn.functional as F

import recipes
from oracle import contrastive_oracle as O


@pytest.fixture(autouse=True)
def _one_thread():
    n = torch.get_num_threads()
    torch.set_num_threads(1)
    yield
    torch.set_num_threads(n)


def _moco_replay(g, n_keys, multi_view):
    B, K, T, m, steps = (int(g.scalar("B")), int(g.scalar("K")), g.scalar("T"), g.scalar("m"),
                         int(g.scalar("steps")))
    queue = g["queue0"].clone()
    hist = [g["Whist0"].clone()] if "Whist0" in g.keys() else [torch.zeros_like(g["W0"])]
    ptr = 0
    for s in range(steps):
        W = g["W%d" % s] if ("W%d" % s) in g.keys() else g["W0"]
        hist = O.ema_update([W], hist, m, s)
        assert torch.equal(hist[0], g["Whist_after%d" % s])
        if n_keys == 1:
            xq, xks = g["xq%d" % s], [g["xk%d" % s]]
        else:
            xq = g["x%d_0" % s]
            xks = [g["x%d_%d" % (s, i)] for i in range(1, n_keys + 1)]
        fq = F.linear(xq, W).requires_grad_(True)
        assert torch.equal(fq.detach(), g["featq%d" % s])
        keys = [O.l2_normalize(F.linear(xk, hist[0])) for xk in xks]
        q, logits, loss = O.moco_head(fq, keys, queue, T)
        loss.backward()
        assert torch.equal(logits.detach(), g["logits%d" % s])
        assert torch.equal(loss.detach(), g["loss%d" % s])
        assert torch.equal(fq.grad, g["dfeatq%d" % s])
        # independent closed form agrees with autograd (fp64 vs fp32 reference)
        cl, cdf, _ = O.moco_head_closed_form(fq.detach(), keys, queue, T)
        assert abs(cl.item() - loss.item()) < 2e-6 * abs(loss.item())
        assert (cdf.float() - fq.grad).abs().max() < 1e-6 * fq.grad.abs().max() + 1e-9
        ptr = O.enqueue(queue, ptr, keys, K, multi_view=multi_view)
        assert torch.equal(queue, g["queue_after%d" % s])
        assert ptr == int(g["ptr_after%d" % s][0])


def test_moco_small(golden):
    _moco_replay(golden("moco_small"), 1, False)


def test_moco_small_shuffle_matches_unshuffled(golden):
    # world_size 1 shuffle is a permutation followed by its inverse around a
    # per-sample backbone: the reference's own outputs do not change.
    a, b = golden("moco_small"), golden("moco_small_shuffle")
    for s in range(3):
        assert torch.equal(a["loss%d" % s], b["loss%d" % s])
        assert torch.equal(a["queue_after%d" % s], b["queue_after%d" % s])


def test_moco_multikey_wrap(golden):
    g = golden("moco_multikey")
    _moco_replay(g, 2, True)
    ptrs = [int(g["ptr_after%d" % s][0]) for s in range(5)]
    assert ptrs == [16, 32, 48, 0, 16]


def test_moco_cfg1(golden):
    g = golden("moco_cfg1")
    r = recipes.moco_cfg1()
    assert abs(r["queue"].double().sum().item() - g.scalar("queue_checksum")) < 1e-9
    hist = O.ema_update([r["W"]], [r["W"].clone()], r["m"], 0)
    assert torch.equal(hist[0], g["Whist_after"])
    fq = F.linear(r["xq"], r["W"]).requires_grad_(True)
    assert torch.equal(fq.detach(), g["featq"])
    keys = [O.l2_normalize(F.linear(r["xk"], hist[0]))]
    queue = r["queue"].clone()
    q, logits, loss = O.moco_head(fq, keys, queue, r["T"])
    loss.backward()
    assert torch.equal(loss.detach(), g["loss"])
    assert torch.equal(fq.grad, g["dfeatq"])
    assert torch.equal(logits[:, :16].detach(), g["logits_head"])
    assert torch.equal(logits[:, -16:].detach(), g["logits_tail"])
    assert torch.allclose(torch.logsumexp(logits.detach().double(), 1), g["lse"], rtol=0, atol=1e-12)
    ptr = O.enqueue(queue, 0, keys, r["K"])
    assert ptr == int(g["ptr_after"][0]) == 64
    assert torch.equal(queue[:64], g["queue_rows_after"])


def _byol_params(g, prefix):
    return [g[prefix + "/" + n] for n in g["param_names"]]


def test_byol(golden):
    g = golden("byol")
    T, m = g.scalar("T"), g.scalar("m")
    names = list(g["param_names"])
    hist = _byol_params(g, "hist0")
    for s in range(int(g.scalar("steps"))):
        online = _byol_params(g, "online%d" % s)
        hist = O.ema_update(online, hist, m, s)
        for n, h in zip(names, hist):  # includes the unused predictor (Q16)
            assert torch.equal(h, g["hist_after%d/%s" % (s, n)])
        hp = dict(zip(names, hist))
        keys = [O.l2_normalize(F.linear(g["x%d_%d" % (i, s)], hp["proj.weight"])) for i in (1, 2)]
        p1 = g["pred1_%d" % s].clone().requires_grad_(True)
        p2 = g["pred2_%d" % s].clone().requires_grad_(True)
        loss = O.byol_pair_loss(p1, p2, keys[0], keys[1], T)
        loss.backward()
        assert torch.equal(loss.detach(), g["loss%d" % s])
        assert torch.equal(p1.grad, g["dpred1_%d" % s])
        assert torch.equal(p2.grad, g["dpred2_%d" % s])
        shp = tuple(int(v) for v in g["logits_shape%d" % s])
        d = O.dummy_logits(shp[0], shp[1] - 1)
        assert torch.equal(d[:, 0], g["logits_col0_%d" % s]) and d[:, 1:].abs().sum() == 0
        assert g["logits_abs_rest_sum%d" % s].item() == 0


def test_simclr(golden):
    g = golden("simclr")
    T = g.scalar("T")
    f1 = g["feat1"].clone().requires_grad_(True)
    f2 = g["feat2"].clone().requires_grad_(True)
    loss = O.ntxent(O.l2_normalize(f1), O.l2_normalize(f2), T)
    loss.backward()
    assert torch.equal(loss.detach(), g["loss"])
    assert torch.equal(f1.grad, g["dfeat1"]) and torch.equal(f2.grad, g["dfeat2"])
    # closed form (fp64) through the normalisation
    q1, q2 = O.l2_normalize(g["feat1"].double()), O.l2_normalize(g["feat2"].double())
    cl, G, _ = O.ntxent_closed_form(q1, q2, T)
    assert abs(cl.item() - loss.item()) < 1e-6 * abs(loss.item())
    f = torch.cat([g["feat1"], g["feat2"]]).double()
    q = torch.cat([q1, q2])
    df = (G - (G * q).sum(1, keepdim=True) * q) / f.norm(dim=1, keepdim=True)
    ref = torch.cat([f1.grad, f2.grad]).double()
    assert (df - ref).abs().max() < 1e-5 * ref.abs().max()


def test_sinkhorn(golden):
    g = golden("sinkhorn")
    for name in "abc":
        Q = torch.exp(g["scores_" + name] / 0.05)
        code = O.sinkhorn(Q, 3)
        assert torch.equal(code, g["code_" + name])
        assert torch.allclose(code.sum(1), torch.ones(code.shape[0]), atol=1e-5)


def test_swav(golden):
    g = golden("swav")
    B, n_crops, T = int(g.scalar("B")), int(g.scalar("n_crops")), g.scalar("T")
    W = O.swav_renorm_prototypes(g["W0"])
    assert torch.equal(W, g["W_after"])
    W = W.requires_grad_(True)
    feats = [g["feat%d" % i].clone().requires_grad_(True) for i in range(n_crops)]
    outs = [O.swav_scores(f, W)[1] for f in feats]
    loss, codes, _ = O.swav_loss(torch.cat(outs, 0), B, n_crops, T)
    loss.backward()
    assert torch.equal(loss.detach(), g["loss"])
    assert torch.equal(W.grad, g["dW"])
    for i, f in enumerate(feats):
        assert torch.equal(f.grad, g["dfeat%d" % i])


def test_swav_queue(golden):
    g = golden("swav_queue")
    B, n_crops, T, L = int(g.scalar("B")), int(g.scalar("n_crops")), g.scalar("T"), int(g.scalar("L"))
    D = int(g.scalar("D"))
    queue = torch.zeros(2, L, D)
    use = False
    seen_use = []
    for s in range(int(g.scalar("steps"))):
        W = O.swav_renorm_prototypes(g["Wpre%d" % s]).requires_grad_(True)
        feats = [g["feat%d_%d" % (s, i)].clone().requires_grad_(True) for i in range(n_crops)]
        eo = [O.swav_scores(f, W) for f in feats]
        emb = torch.cat([e for e, _ in eo], 0).detach()
        out = torch.cat([o for _, o in eo], 0)
        # the reference decides per assign-crop i, sticky (:651-654)
        loss = 0
        newq = queue.clone()
        for i, crop in enumerate((0, 1)):
            with torch.no_grad():
                o_i = out[B * crop:B * (crop + 1)]
                if use or not torch.all(newq[i, -1, :] == 0):
                    use = True
                    o_i = torch.cat((torch.mm(newq[i], W.detach().t()), o_i))
                newq[i] = O.swav_queue_push(newq[i], emb[crop * B:(crop + 1) * B])
                code = O.sinkhorn(torch.exp(o_i / 0.05), 3)[-B:]
            sub = 0
            for v in [c for c in range(n_crops) if c != crop]:
                p = torch.softmax(out[B * v:B * (v + 1)] / T, dim=1)
                sub = sub - torch.mean(torch.sum(code * torch.log(p), dim=1))
            loss = loss + sub / (n_crops - 1)
        loss = loss / 2
        loss.backward()
        queue = newq
        seen_use.append(use)
        assert torch.equal(queue, g["queue_after%d" % s])
        assert bool(g["use_queue%d" % s]) == use
        assert torch.equal(loss.detach(), g["loss%d" % s])
        assert torch.equal(W.grad, g["dW%d" % s])
        for i, f in enumerate(feats):
            assert torch.equal(f.grad, g["dfeat%d_%d" % (s, i)])
    assert seen_use[-1] and not seen_use[0]


def test_membank(golden):
    g = golden("membank")
    for tag, mom in (("half", 0.5), ("one", 1.0)):
        bank = g["m2d_%s_bank0" % tag].clone()
        O.membank_update(bank, g["m2d_%s_upd" % tag], mom, g["m2d_%s_ind" % tag], g["m2d_%s_time" % tag])
        assert torch.equal(bank, g["m2d_%s_bank1" % tag])
    bank = g["mi_bank0"].clone()
    got = O.membank_get(bank, g["mi_ind"], g["mi_time"], interp=True)
    assert torch.equal(got, g["mi_got"])
    O.membank_update(bank, g["mi_upd"], 0.7, g["mi_ind"], g["mi_time"], interp=True)
    assert torch.equal(bank, g["mi_bank1"])
    bank = g["m1d_bank0"].clone()
    O.memory1d_update(bank, g["m1d_upd"], 0.3, g["m1d_ind"])
    assert torch.equal(bank, g["m1d_bank1"])


def test_mem_mode(golden):
    g = golden("mem_mode")
    B, K, L = (int(g.scalar(k)) for k in ("B", "K", "L"))
    T, m = g.scalar("T"), g.scalar("m")
    for tag, one_d, interp in (("1d", True, False), ("2di", False, True)):
        q = O.l2_normalize(g[tag + "_featq"])
        torch.manual_seed(72)
        clip_ind = torch.randint(0, L, size=(B, K + 1))  # models/contrastive.py:390-397
        clip_ind.select(1, 0).copy_(g[tag + "_index"])
        if one_d:
            time_ind = torch.zeros(size=(B, K + 1), dtype=int)
        else:
            time_ind = torch.empty(B, K + 1).uniform_(0, 0)
        bank = g[tag + "_bank0"].clone()
        if one_d:
            prod = O.mem_mode_prod(q, bank, clip_ind, time_ind, T, True)
        else:
            k = O.membank_get(bank, clip_ind, time_ind, interp=True)
            prod = torch.div(torch.einsum("nc,nkc->nk", q, k), T)
        assert torch.equal(prod, g[tag + "_prod"])
        tzero = torch.zeros(B)
        if one_d:
            O.memory1d_update(bank, q, m, g[tag + "_index"])
        else:
            O.membank_update(bank, q, m, g[tag + "_index"], tzero, interp=True)
        assert torch.equal(bank, g[tag + "_bank1"])
        knn = g[tag + "_knn0"].clone()
        O.membank_update(knn, q, 1.0, g[tag + "_index"], torch.zeros_like(g[tag + "_index"]))
        assert torch.equal(knn, g[tag + "_knn1"])
        assert list(g[tag + "_ret"]) == [0.0, 1.0]


def test_ema_annealed(golden):
    g = golden("ema")
    names = list(g["names"])
    hist = [g["hist_init/" + n] for n in names]
    for s, ep in enumerate(g["epochs"].tolist()):
        m = O.momentum_cosine(g.scalar("m0"), ep, int(g.scalar("max_epoch")))
        assert m == g.scalar("mmt%d" % s)
        online = [g["online%d/%s" % (s, n)] for n in names]
        hist = O.ema_update(online, hist, m, s)
        for n, h in zip(names, hist):
            assert torch.equal(h, g["hist%d/%s" % (s, n)])
    assert math.isclose(O.momentum_cosine(0.9, 10, 10), 1.0)


def test_shuffle_roundtrip():
    torch.manual_seed(0)
    W, B = 4, 6
    parts = [torch.randn(B, 5) for _ in range(W)]
    perm = torch.randperm(W * B)
    sh, restore = O.shuffle_emulated(parts, perm)
    allx = torch.cat(parts)
    for r in range(W):
        assert torch.equal(sh[r], allx[perm.view(W, -1)[r]])
    back = O.unshuffle_emulated(sh, restore)
    for r in range(W):
        assert torch.equal(back[r], parts[r])


def test_distributed_sinkhorn_matches_local_when_one_rank():
    torch.manual_seed(1)
    Q = torch.exp((torch.rand(12, 40) * 2 - 1) / 0.05)
    a = O.sinkhorn(Q, 3)
    (b,) = O.distributed_sinkhorn_emulated([Q.t().clone()], 3)
    assert torch.allclose(a, b, rtol=1e-5, atol=1e-8)


# ------------------------------------------- TemporalModel path
def _temporal_names(g):
    return [k[len("hist_init/"):] for k in g.keys() if k.startswith("hist_init/")]


def test_temporal_ema(golden):
    """temporal.npz comes from the reference's own TemporalModel._update_history
    (tests/golden/make_golden_temporal.py): three steps, the first one initialising."""
    g = golden("temporal")
    names = _temporal_names(g)
    hist = [g["hist_init/" + n] for n in names]
    for s in range(int(g.scalar("n_steps"))):
        online = [g["online%d/%s" % (s, n)] for n in names]
        hist = O.temporal_ema_update(online, hist, g.scalar("m"), s == 0)
        for h, n in zip(hist, names):
            assert torch.equal(h, g["hist%d/%s" % (s, n)]), (s, n)


def test_temporal_contrast_loss(golden):
    g = golden("temporal")
    T = g.scalar("T")
    s = int(g.scalar("n_steps")) - 1
    lin = torch.nn.functional.linear
    feats = [g["feat%d" % i].clone().requires_grad_(True) for i in range(2)]
    keys = [g["key1"], g["key0"]]  # keys[::-1], :356
    Wp, bp = g["online%d/proj/weight" % s], g["online%d/proj/bias" % s]
    Wq, bq = g["online0/pred/weight"], g["online0/pred/bias"]
    Wh, bh = g["hist%d/proj/weight" % s], g["hist%d/proj/bias" % s]
    qs = [lin(lin(f, Wp, bp), Wq, bq) for f in feats]
    ks = [lin(k, Wh, bh) for k in keys]
    loss = O.temporal_contrast_loss(qs, ks, T)
    loss.backward()
    assert torch.equal(loss.detach(), g["loss"])
    for i in range(2):
        assert torch.equal(feats[i].grad, g["dfeat%d" % i])


def test_grad_norm(golden):
    """gradnorm.npz comes from the reference's own get_grad_norm_ (make_golden_gradnorm.py)."""
    g = golden("gradnorm")
    grads = [g["grad%d" % i] for i in range(int(g.scalar("n")))]
    assert torch.equal(O.grad_norm(grads), g["total"])
    assert O.grad_norm([]).item() == g["empty"].item() == 0.0


def test_projection_tail(golden):
    """projtail.npz comes from the reference's own MLPHead + Normalize (make_golden_projtail.py): the oracle's
    restatement of the tail (last Linear + Normalize) on the tapped input reproduces q bit for bit."""
    g = golden("projtail")
    for case in ("a", "b"):
        layers = int(g[case + "_cfg"][4])
        last = "projection.%d" % (3 * (layers - 1) if int(g[case + "_cfg"][5]) else 2 * (layers - 1))
        W = g["%s_sd_%s.weight" % (case, last)]
        key_b = "%s_sd_%s.bias" % (case, last)
        b = g[key_b] if key_b in g.keys() else None
        assert torch.equal(torch.nn.functional.linear(g[case + "_tail_x"], W, b), g[case + "_tail_y"])
        assert torch.equal(O.projection_tail(g[case + "_tail_x"], W, b), g[case + "_q"])


def test_knn_topk(golden):
    """knn.npz comes from the reference's own eval_knn and the eval branch of forward (make_golden_knn.py)."""
    g = golden("knn")
    for k in (200, 5):
        yd, yi = O.knn_topk(g["q"], g["bank"], k)
        assert torch.equal(yd, g["yd%d" % k]) and torch.equal(yi, g["yi%d" % k])
    q = O.l2_normalize(torch.nn.functional.linear(g["x"], g["W"]))
    yd, yi = O.knn_topk(q, g["bank"], 200)
    assert torch.equal(yd, g["fwd_yd"]) and torch.equal(yi, g["fwd_yi"])


# ------------------------------------------------------------------ oracle self-consistency (size-independent properties)
@pytest.mark.parametrize("B,D,K,nk,T", [(5, 16, 40, 1, 0.1), (9, 32, 100, 3, 0.07), (1, 4, 1, 2, 0.5)])
def test_moco_closed_form_agrees_with_autograd(B, D, K, nk, T):
    """The independent fp64 closed form (what the GPU tests use at sizes autograd would be slow for) reproduces the
    autograd gradient of the restated forward (models/contrastive.py:462-500, losses.py:20-25)."""
    g = torch.Generator().manual_seed(B + K)
    f = torch.randn(B, D, generator=g, dtype=torch.float64).requires_grad_(True)
    keys = [O.l2_normalize(torch.randn(B, D, generator=g, dtype=torch.float64)) for _ in range(nk)]
    queue = O.l2_normalize(torch.randn(K, D, generator=g, dtype=torch.float64))
    _, logits, loss = O.moco_head(f, keys, queue, T)
    loss.backward()
    cl, cdf, _ = O.moco_head_closed_form(f.detach(), keys, queue, T)
    assert abs(cl.item() - loss.item()) < 1e-12 * max(1.0, abs(loss.item()))
    assert (cdf - f.grad).abs().max().item() < 1e-12
    assert tuple(logits.shape) == (nk * B, K + 1)


@pytest.mark.parametrize("N,D,T", [(4, 8, 0.5), (11, 16, 0.1)])
def test_ntxent_closed_form_agrees_with_autograd(N, D, T):
    g = torch.Generator().manual_seed(N)
    a = O.l2_normalize(torch.randn(N, D, generator=g, dtype=torch.float64)).requires_grad_(True)
    b = O.l2_normalize(torch.randn(N, D, generator=g, dtype=torch.float64)).requires_grad_(True)
    loss = O.ntxent(a, b, T)
    loss.backward()
    cl, G, _ = O.ntxent_closed_form(a.detach(), b.detach(), T)
    assert abs(cl.item() - loss.item()) < 1e-12
    assert (G - torch.cat([a.grad, b.grad])).abs().max().item() < 1e-12


def test_enqueue_ring_properties():
    """:263-292: the pointer stays a multiple of the batch inside [0, K), wraps exactly at K, and after K / n pushes
    every row of the queue has been rewritten, oldest first."""
    K, n, D = 24, 4, 3
    queue = torch.zeros(K, D)
    ptr = 0
    for step in range(2 * K // n + 1):
        ptr = O.enqueue(queue, ptr, [torch.full((n, D), float(step + 1))], K)
        assert 0 <= ptr < K and ptr % n == 0 and ptr == ((step + 1) * n) % K
    assert torch.equal(queue[:n], torch.full((n, D), float(2 * K // n + 1)))      # the newest block
    assert torch.equal(queue[n:2 * n], torch.full((n, D), float(K // n + 2)))     # the oldest surviving block
    with pytest.raises(AssertionError):
        O.enqueue(queue, 0, [torch.zeros(5, D)], K)                               # K % n != 0 (:284)


def test_ema_limits():
    """:158-172: m = 1 keeps the history bits, m = 0 copies the online weights, iteration 0 copies first (Q11)."""
    on = [torch.randn(7, 3), torch.randn(5)]
    hi = [torch.randn(7, 3), torch.randn(5)]
    assert all(torch.equal(a, b) for a, b in zip(O.ema_update(on, [h.clone() for h in hi], 1.0, 3), hi))
    assert all(torch.equal(a, b) for a, b in zip(O.ema_update(on, [h.clone() for h in hi], 0.0, 3), on))
    first = O.ema_update(on, [h.clone() for h in hi], 0.9, 0)
    assert all(torch.equal(a, o * (1.0 - 0.9) + o * 0.9) for a, o in zip(first, on))
