"""-m gpu parity tests of the drop-in `ContrastiveModel` (all five modes, the banks,
the driver functions) against the golden vectors produced by the unmodified
reference and against the CPU oracle.  Everything goes through the C-ABI library.

Tolerances: north_star asks 1e-3 relative for fp32 loss and gradients and bit-exact
queue pointers / bank indices.  The CUDA-core kernels meet much tighter bounds,
written at each assert; EMA and integer state are compared with torch.equal.
"""
import numpy as np
import pytest
import torch

from helpers import Tap, make_cfg, register_backbones, rel_err, rel_err_above_floor
from oracle import contrastive_oracle as O

pytestmark = pytest.mark.gpu
LOSS_RTOL = 2e-6
# MoCo goes through the default (AUTO) InfoNCE kernel = tcgen05 3xTF32: logits/loss are
# fp32-grade, the gradient term sum_j p_ij queue_j carries one unbiased tf32 rounding of P and
# of the queue (<= 2.5e-4 relative, measured 2e-4 worst case); north_star allows 1e-3.
GRAD_RTOL = 5e-4


def _model(cfg, seed=0):
    C = register_backbones()
    torch.manual_seed(seed)
    m = C.ContrastiveModel(cfg).cuda().train()
    return C, m


def _set(p, v):
    with torch.no_grad():
        p.copy_(v)


# ------------------------------------------------------------------------------ MoCo
@pytest.mark.parametrize("sweep_ctas", [None, 0, 3])
@pytest.mark.parametrize("shuffle", [False, True])
def test_moco_small_golden(golden, shuffle, sweep_ctas):
    """sweep_ctas None: the single-launch head (the update of this tiny encoder is too short to hide a sweep under);
    0 / 3: the two-launch head -- sweep on its own stream before the momentum update, finish behind the key path."""
    g = golden("moco_small")
    B, D, K, T, m_ = int(g.scalar("B")), int(g.scalar("D")), int(g.scalar("K")), g.scalar("T"), g.scalar("m")
    cfg = make_cfg(CONTRASTIVE__TYPE="moco", CONTRASTIVE__T=T, CONTRASTIVE__DIM=D, CONTRASTIVE__QUEUE_LEN=K,
                   CONTRASTIVE__MOMENTUM=m_)
    C, model = _model(cfg)
    assert model._batch_shuffle_on  # reference default at 1 GPU without sync BN
    model._batch_shuffle_on = shuffle
    assert model.sweep_ctas is None and model._sweep_ctas_beside_ema(B, torch.device("cuda", 0)) is None
    model.sweep_ctas = sweep_ctas
    _set(model.queue_x, g["queue0"])
    _set(model.backbone_hist.proj.weight, g["Whist0"])
    tap = Tap(model.backbone)
    for s in range(3):
        _set(model.backbone.proj.weight, g["W%d" % s])
        model.zero_grad()
        xq, xk = g["xq%d" % s].cuda(), g["xk%d" % s].cuda()
        logits, loss = model([[xq], [xk]], torch.arange(B).cuda(), torch.zeros(B, 2, 1).cuda(), 0.0)
        loss.backward()
        (fq,) = tap.pop()[0]
        assert rel_err(loss, g["loss%d" % s]) < LOSS_RTOL
        assert rel_err(fq.grad, g["dfeatq%d" % s]) < GRAD_RTOL
        assert rel_err(model.backbone.proj.weight.grad, g["dW%d" % s]) < GRAD_RTOL
        assert (logits.cpu() - g["logits%d" % s]).abs().max().item() < 2e-5
        # EMA: bit-exact; pointer / iteration counter: bit-exact integers
        assert torch.equal(model.backbone_hist.proj.weight.cpu(), g["Whist_after%d" % s])
        assert torch.equal(model.ptr.cpu(), g["ptr_after%d" % s]) and model.ptr.dtype == torch.int64
        assert torch.equal(model.iter.cpu(), g["iter_after%d" % s])
        # queue: untouched rows bit-exact, enqueued keys equal to 1 ulp (l2-norm sum order)
        qa, qr = model.queue_x.cpu(), g["queue_after%d" % s]
        lo = s * B
        assert torch.equal(qa[lo + B:], qr[lo + B:])
        assert (qa[:lo + B] - qr[:lo + B]).abs().max().item() < 2e-7
    assert model.check_device_status() == 0


def test_moco_multikey_multiview_queue(golden):
    g = golden("moco_multikey")
    B, D, K, T, m_ = int(g.scalar("B")), int(g.scalar("D")), int(g.scalar("K")), g.scalar("T"), g.scalar("m")
    cfg = make_cfg(CONTRASTIVE__TYPE="moco", CONTRASTIVE__T=T, CONTRASTIVE__DIM=D, CONTRASTIVE__QUEUE_LEN=K,
                   CONTRASTIVE__MOMENTUM=m_, CONTRASTIVE__MOCO_MULTI_VIEW_QUEUE=True)
    C, model = _model(cfg)
    model._batch_shuffle_on = False
    _set(model.queue_x, g["queue0"])
    _set(model.backbone.proj.weight, g["W0"])
    tap = Tap(model.backbone)
    for s in range(5):
        model.zero_grad()
        xs = [g["x%d_%d" % (s, i)].cuda() for i in range(3)]
        logits, loss = model([[x] for x in xs], torch.arange(B).cuda(), torch.zeros(B, 3, 1).cuda(), 0.0)
        loss.backward()
        (fq,) = tap.pop()[0]
        assert logits.shape == (2 * B, K + 1)
        assert rel_err(loss, g["loss%d" % s]) < LOSS_RTOL
        assert rel_err(fq.grad, g["dfeatq%d" % s]) < GRAD_RTOL
        assert (logits.cpu() - g["logits%d" % s]).abs().max().item() < 2e-5
        assert torch.equal(model.ptr.cpu(), g["ptr_after%d" % s])
        assert (model.queue_x.cpu() - g["queue_after%d" % s]).abs().max().item() < 2e-7
        assert torch.equal(model.backbone_hist.proj.weight.cpu(), g["Whist_after%d" % s])


def test_moco_index_none_eval_and_lean_logits(golden):
    cfg = make_cfg(CONTRASTIVE__TYPE="moco", CONTRASTIVE__T=0.1, CONTRASTIVE__DIM=32, CONTRASTIVE__QUEUE_LEN=64,
                   CONTRASTIVE__KNN_ON=True, CONTRASTIVE__LENGTH=300)
    C, model = _model(cfg)
    x = torch.randn(8, 32).cuda()
    f = model([[x], [x]], None, torch.zeros(8, 2, 1).cuda())
    assert torch.equal(f, model.backbone([x]))
    # knn bank update with duplicate indices: last wins, rows are the normalised q
    idx = torch.tensor([5, 9, 5, 1, 299, 0, 9, 7]).cuda()
    bank0 = model.knn_mem.memory.clone()
    model.materialize_logits = False
    logits, loss = model([[x], [x]], idx, torch.zeros(8, 2, 1).cuda(), 0.0)
    assert logits is None and loss.requires_grad
    q = O.l2_normalize(model.backbone([x]).detach().cpu())
    ref = bank0.cpu().clone()
    O.membank_update(ref, q, 1.0, idx.cpu(), torch.zeros_like(idx.cpu()))
    assert (model.knn_mem.memory.cpu() - ref).abs().max().item() < 2e-7
    model.eval()
    yd, yi = model([[x], [x]], idx, torch.zeros(8, 2, 1).cuda(), 0.0)
    assert yd.shape == (8, 200) and yi.dtype == torch.int64
    model.train()
    with pytest.raises(TypeError):  # time=None is indexed by the reference too
        model([[x], [x]], idx, None, 0.0)


# ------------------------------------------------------------------------------ BYOL
def test_byol_golden(golden):
    g = golden("byol")
    B, D, T, m_ = int(g.scalar("B")), int(g.scalar("D")), g.scalar("T"), g.scalar("m")
    cfg = make_cfg(CONTRASTIVE__TYPE="byol", CONTRASTIVE__T=T, CONTRASTIVE__DIM=D, CONTRASTIVE__QUEUE_LEN=256,
                   CONTRASTIVE__MOMENTUM=m_, CONTRASTIVE__PREDICTOR_DEPTHS=[1])
    C, model = _model(cfg)
    assert not model._batch_shuffle_on
    names = list(g["param_names"])
    for n, p in model.backbone_hist.named_parameters():
        _set(p, g["hist0/" + n])
    tap = Tap(model.backbone)
    for s in range(int(g.scalar("steps"))):
        for n, p in model.backbone.named_parameters():
            _set(p, g["online%d/%s" % (s, n)])
        model.zero_grad()
        x1, x2 = g["x1_%d" % s].cuda(), g["x2_%d" % s].cuda()
        logits, loss = model([[x1], [x2]], torch.arange(B).cuda(), None, 0.0)
        loss.backward()
        o = tap.pop()
        (f1, p1), (f2, p2) = o[0], o[1]
        # the loss is a mean of similarities in [-1/T, 1/T] that nearly cancel here
        assert abs(loss.item() - g["loss%d" % s].item()) < 5e-6 * abs(g["loss%d" % s].item()) + 2e-7 / T
        assert rel_err(p1.grad, g["dpred1_%d" % s]) < GRAD_RTOL
        assert rel_err(p2.grad, g["dpred2_%d" % s]) < GRAD_RTOL
        for n, p in model.backbone_hist.named_parameters():  # incl. the unused predictor (Q16)
            assert torch.equal(p.cpu(), g["hist_after%d/%s" % (s, n)]), n
        assert tuple(logits.shape) == tuple(int(v) for v in g["logits_shape%d" % s])
        assert torch.equal(logits[:, 0].cpu(), g["logits_col0_%d" % s]) and logits[:, 1:].abs().sum().item() == 0
    assert names


def test_sim_loss_public_method():
    cfg = make_cfg(CONTRASTIVE__TYPE="byol", CONTRASTIVE__T=0.3, CONTRASTIVE__DIM=48, CONTRASTIVE__QUEUE_LEN=64,
                   CONTRASTIVE__PREDICTOR_DEPTHS=[1])
    C, model = _model(cfg)
    q = torch.randn(10, 48).cuda().requires_grad_(True)
    k = torch.randn(10, 48).cuda()
    loss = model.sim_loss(q, k)
    loss.backward()
    qc = q.detach().cpu().requires_grad_(True)
    ref = O.byol_sim_loss(qc, k.cpu(), 0.3)
    ref.backward()
    assert rel_err(loss, ref.detach()) < 2e-6 and rel_err(q.grad, qc.grad) < 2e-6


# ---------------------------------------------------------------------------- SimCLR
def test_simclr_golden(golden):
    g = golden("simclr")
    B, D, T = int(g.scalar("B")), int(g.scalar("D")), g.scalar("T")
    cfg = make_cfg(CONTRASTIVE__TYPE="simclr", CONTRASTIVE__T=T, CONTRASTIVE__DIM=D, CONTRASTIVE__QUEUE_LEN=32,
                   TRAIN__BATCH_SIZE=B)
    C = register_backbones()
    cfg.MODEL.ARCH = "identity"
    from advise_video_ssl_b200 import _lib
    model = C.ContrastiveModel(cfg).cuda().train()
    # exact-fp32 CUDA-core kernels, then the default (tcgen05 tf32 when D allows; fp32 tolerance 1e-3)
    for impl, lt, gt in ((_lib.IMPL_SIMT, 5e-6, 5e-5), (_lib.IMPL_AUTO, 2e-4, 1e-3)):
        model.ntxent_impl = impl
        f1 = g["feat1"].cuda().requires_grad_(True)
        f2 = g["feat2"].cuda().requires_grad_(True)
        logits, loss = model([[f1], [f2]], torch.arange(B).cuda(), None, 0.0)
        loss.backward()
        assert rel_err(loss, g["loss"]) < lt
        assert rel_err(f1.grad, g["dfeat1"]) < gt and rel_err(f2.grad, g["dfeat2"]) < gt
        assert tuple(logits.shape) == tuple(int(v) for v in g["logits_shape"])
        assert torch.equal(logits[:, 0].cpu(), g["logits_col0"])


@pytest.mark.parametrize("impl_name", ["simt", "tc"])
@pytest.mark.parametrize("B,D,T", [(3, 8, 0.5), (100, 128, 0.1), (70, 256, 0.07), (256, 64, 0.2), (512, 256, 0.1),
                                   (33, 32, 0.1), (200, 96, 0.5)])
def test_ntxent_shapes_vs_closed_form(B, D, T, impl_name):
    from advise_video_ssl_b200 import ops, _lib
    torch.manual_seed(B)
    f1, f2 = torch.randn(B, D) * 2, torch.randn(B, D) * 0.5
    if impl_name == "tc" and D not in (64, 128, 256):  # whole 128-byte boxes of fp16; other D: CUDA-core kernels
        with pytest.raises(_lib.AvsslError, match="tcgen05 kernel needs D"):
            ops.ntxent(f1.cuda(), f2.cuda(), T, impl=_lib.IMPL_TC1X)
        return
    impl = _lib.IMPL_SIMT if impl_name == "simt" else _lib.IMPL_TC1X
    loss, d1, d2 = ops.ntxent(f1.cuda(), f2.cuda(), T, impl=impl)
    q1, q2 = O.l2_normalize(f1.double()), O.l2_normalize(f2.double())
    cl, G, _ = O.ntxent_closed_form(q1, q2, T)
    f = torch.cat([f1, f2]).double()
    q = torch.cat([q1, q2])
    df = (G - (G * q).sum(1, keepdim=True) * q) / f.norm(dim=1, keepdim=True)
    # CUDA-core kernels: exact fp32.  tcgen05 kernels: fp16 copies of the unit rows (the precision of
    # round-to-nearest tf32 on [-1, 1]), fp32 accumulation -- stated separately, inside north_star's 1e-3.
    lt, gt = (5e-6, 1e-4) if impl_name == "simt" else (2e-4, 1e-3)
    assert abs(loss.item() - cl.item()) < lt * abs(cl.item())
    assert rel_err(torch.cat([d1, d2]), df) < gt
    # element-wise on the entries that reach 5 % (exact-fp32 kernels, measured 6e-7) / 10 % (fp16-operand tensor-core
    # kernels, measured <= 4.7e-4) of the largest one
    if impl_name == "simt":
        assert rel_err_above_floor(torch.cat([d1, d2]), df, 0.05) < 2e-5
    else:
        assert rel_err_above_floor(torch.cat([d1, d2]), df, 0.1) < 2e-3


# ------------------------------------------------------------------------------ SwAV
def test_sinkhorn_golden(golden):
    from advise_video_ssl_b200 import ops
    g = golden("sinkhorn")
    for name in "abc":
        code = ops.sinkhorn(g["scores_" + name].cuda(), 0.05, 3)
        ref = g["code_" + name]
        assert rel_err(code, ref) < 2e-5
        assert torch.allclose(code.sum(1).cpu(), torch.ones(ref.shape[0]), atol=1e-5)


@pytest.mark.parametrize("Btot,P,keep,iters", [(256, 3000, 256, 3), (4096, 1000, 256, 3), (1, 5, 1, 1), (37, 129, 10, 0),
                                               (600, 3000, 64, 5)])
def test_sinkhorn_shapes_vs_oracle(Btot, P, keep, iters):
    """incl. BASELINE cfg5 (3000 prototypes x 256) and a queue-extended case that does
    not fit shared memory (recompute path)."""
    from advise_video_ssl_b200 import ops
    torch.manual_seed(P)
    scores = torch.rand(Btot, P) * 2 - 1
    code = ops.sinkhorn(scores.cuda(), 0.05, iters, keep_last=keep)
    ref = O.sinkhorn(torch.exp(scores.double() / 0.05), iters)[-keep:]
    assert code.shape == (keep, P)
    assert rel_err(code, ref) < 5e-5


def test_swav_golden(golden):
    g = golden("swav")
    B, D, T, n_crops = int(g.scalar("B")), int(g.scalar("D")), g.scalar("T"), int(g.scalar("n_crops"))
    cfg = make_cfg(CONTRASTIVE__TYPE="swav", CONTRASTIVE__T=T, CONTRASTIVE__DIM=D, CONTRASTIVE__QUEUE_LEN=64)
    C = register_backbones()
    cfg.MODEL.ARCH = "identity"
    model = C.ContrastiveModel(cfg).cuda().train()
    assert model.swav_prototypes.weight.shape == (1000, D)  # reference's hard-coded count
    _set(model.swav_prototypes.weight, g["W0"])
    feats = [g["feat%d" % i].cuda().requires_grad_(True) for i in range(n_crops)]
    logits, loss = model([[f] for f in feats], torch.arange(B).cuda(), None, 0.0)
    loss.backward()
    assert (model.swav_prototypes.weight.detach().cpu() - g["W_after"]).abs().max().item() < 2e-7
    assert rel_err(loss, g["loss"]) < 5e-6
    assert rel_err(model.swav_prototypes.weight.grad, g["dW"]) < 1e-4
    for i, f in enumerate(feats):
        assert rel_err(f.grad, g["dfeat%d" % i]) < 1e-4
    assert tuple(logits.shape) == tuple(int(v) for v in g["logits_shape"])


def test_swav_queue_golden(golden):
    g = golden("swav_queue")
    B, D, T, n_crops, L = (int(g.scalar("B")), int(g.scalar("D")), g.scalar("T"), int(g.scalar("n_crops")),
                           int(g.scalar("L")))
    cfg = make_cfg(CONTRASTIVE__TYPE="swav", CONTRASTIVE__T=T, CONTRASTIVE__DIM=D, CONTRASTIVE__QUEUE_LEN=64,
                   CONTRASTIVE__SWAV_QEUE_LEN=L)
    C = register_backbones()
    cfg.MODEL.ARCH = "identity"
    model = C.ContrastiveModel(cfg).cuda().train()
    for s in range(int(g.scalar("steps"))):
        _set(model.swav_prototypes.weight, g["Wpre%d" % s])
        model.zero_grad()
        feats = [g["feat%d_%d" % (s, i)].cuda().requires_grad_(True) for i in range(n_crops)]
        logits, loss = model([[f] for f in feats], torch.arange(B).cuda(), None, 15.0)
        loss.backward()
        assert bool(g["use_queue%d" % s]) == model.swav_use_the_queue
        assert (model.queue_swav.cpu() - g["queue_after%d" % s]).abs().max().item() < 2e-7
        assert rel_err(loss, g["loss%d" % s]) < 1e-5
        assert rel_err(model.swav_prototypes.weight.grad, g["dW%d" % s]) < 2e-4
        for i, f in enumerate(feats):
            assert rel_err(f.grad, g["dfeat%d_%d" % (s, i)]) < 2e-4


def test_swav_3000_prototypes_cfg5():
    """BASELINE configs[4] shape: 6 crops, 3000 prototypes, B=256 (prototype count is a
    cfg key here; the reference hard-codes 1000)."""
    B, D, T, n_crops, P = 256, 128, 0.1, 6, 3000
    cfg = make_cfg(CONTRASTIVE__TYPE="swav", CONTRASTIVE__T=T, CONTRASTIVE__DIM=D, CONTRASTIVE__QUEUE_LEN=64)
    cfg.CONTRASTIVE.SWAV_NUM_PROTOTYPES = P
    C = register_backbones()
    cfg.MODEL.ARCH = "identity"
    torch.manual_seed(5)
    model = C.ContrastiveModel(cfg).cuda().train()
    feats_c = [torch.randn(B, D) for _ in range(n_crops)]
    feats = [f.cuda().requires_grad_(True) for f in feats_c]
    W0 = model.swav_prototypes.weight.detach().cpu().clone()
    logits, loss = model([[f] for f in feats], torch.arange(B).cuda(), None, 0.0)
    loss.backward()
    W = O.swav_renorm_prototypes(W0).requires_grad_(True)
    fc = [f.clone().requires_grad_(True) for f in feats_c]
    outs = [O.swav_scores(f, W)[1] for f in fc]
    ref, _, _ = O.swav_loss(torch.cat(outs, 0), B, n_crops, T)
    ref.backward()
    assert rel_err(loss, ref.detach()) < 1e-5
    assert rel_err(model.swav_prototypes.weight.grad, W.grad) < 2e-4
    assert rel_err(feats[3].grad, fc[3].grad) < 2e-4


# ------------------------------------------------------------------- banks / mem mode
def test_membank_golden(golden):
    from advise_video_ssl_b200 import ops
    g = golden("membank")
    for tag, mom in (("half", 0.5), ("one", 1.0)):
        bank = g["m2d_%s_bank0" % tag].cuda()
        ops.membank_update(bank, g["m2d_%s_upd" % tag].cuda(), g["m2d_%s_ind" % tag].cuda(),
                           g["m2d_%s_time" % tag].cuda(), mom)
        ref = g["m2d_%s_bank1" % tag]
        changed = (ref != g["m2d_%s_bank0" % tag]).any(-1)
        assert torch.equal(bank.cpu()[~changed], ref[~changed])  # indices bit-exact: only those rows moved
        assert (bank.cpu() - ref).abs().max().item() < 2e-7
    bank = g["mi_bank0"].cuda()
    ops.membank_update(bank, g["mi_upd"].cuda(), g["mi_ind"].cuda(), g["mi_time"].cuda(), 0.7, interp=True)
    assert (bank.cpu() - g["mi_bank1"]).abs().max().item() < 2e-7
    bank = g["m1d_bank0"].cuda()
    ops.membank_update(bank, g["m1d_upd"].cuda(), g["m1d_ind"].cuda(), None, 0.3)
    assert (bank.cpu() - g["m1d_bank1"]).abs().max().item() < 2e-7
    # out-of-range index: flagged, row skipped
    status = torch.zeros(1, dtype=torch.int32, device="cuda")
    before = bank.clone()
    ops.membank_update(bank, torch.randn(1, 24).cuda(), torch.tensor([30]).cuda(), None, 0.3, status=status)
    assert int(status.item()) == 2 and torch.equal(bank, before)


@pytest.mark.parametrize("tag,mem_type,interp", [("1d", "1d", False), ("2di", "2d", True)])
def test_mem_mode_golden(golden, tag, mem_type, interp):
    g = golden("mem_mode")
    B, D, K, L, T, m_ = (int(g.scalar("B")), int(g.scalar("D")), int(g.scalar("K")), int(g.scalar("L")),
                         g.scalar("T"), g.scalar("m"))
    cfg = make_cfg(CONTRASTIVE__TYPE="mem", CONTRASTIVE__T=T, CONTRASTIVE__DIM=D, CONTRASTIVE__QUEUE_LEN=K,
                   CONTRASTIVE__LENGTH=L, CONTRASTIVE__MOMENTUM=m_, CONTRASTIVE__MEM_TYPE=mem_type,
                   CONTRASTIVE__INTERP_MEMORY=interp, CONTRASTIVE__KNN_ON=True)
    C = register_backbones()
    cfg.MODEL.ARCH = "identity"
    model = C.ContrastiveModel(cfg).cuda().train()
    _set(model.memory.memory, g[tag + "_bank0"])
    _set(model.knn_mem.memory, g[tag + "_knn0"])
    torch.manual_seed(72)  # same CPU draws for the negatives as the golden run
    prod, zero, flag = model([g[tag + "_featq"].cuda()], g[tag + "_index"].cuda(), torch.zeros(B).cuda(), 0.0)
    assert (zero, flag) == (0.0, True)  # the reference's 3-tuple
    assert (prod.cpu() - g[tag + "_prod"]).abs().max().item() < 5e-6
    assert (model.memory.memory.cpu() - g[tag + "_bank1"]).abs().max().item() < 2e-7
    assert (model.knn_mem.memory.cpu() - g[tag + "_knn1"]).abs().max().item() < 2e-7
    assert model.check_device_status() == 0


def test_contrastive_loss_module():
    from advise_video_ssl_b200 import losses
    torch.manual_seed(0)
    lg = (torch.randn(37, 1001) * 4)
    a = lg.cuda().requires_grad_(True)
    loss = losses.get_loss_func("contrastive_loss")(reduction="mean")(a)
    (loss * 3.0).backward()
    b = lg.clone().requires_grad_(True)
    ref = O.info_nce(b)
    (ref * 3.0).backward()
    assert rel_err(loss, ref.detach()) < 2e-6 and rel_err(a.grad, b.grad) < 1e-5


def test_normalize_module():
    from advise_video_ssl_b200.contrastive import Normalize
    x = torch.randn(9, 5, 33)
    for dim in (1, 2):
        a = x.cuda().requires_grad_(True)
        y = Normalize(dim=dim)(a)
        y.square().sum().backward()  # gradient of a constant: ~0
        b = x.clone().requires_grad_(True)
        yr = O.l2_normalize(b, dim=dim)
        assert (y.cpu() - yr.detach()).abs().max().item() < 3e-7
        w = torch.randn_like(x)
        a.grad = None
        (Normalize(dim=dim)(a) * w.cuda()).sum().backward()
        (yr * w).sum().backward()
        assert rel_err(a.grad, b.grad) < 2e-5


# ------------------------------------------------------------------- driver functions
@pytest.mark.parametrize("sequential", [False, True])
def test_contrastive_forward_moco(sequential):
    B, D, K, T = 16, 64, 256, 0.1
    cfg = make_cfg(CONTRASTIVE__TYPE="moco", CONTRASTIVE__T=T, CONTRASTIVE__DIM=D, CONTRASTIVE__QUEUE_LEN=K,
                   CONTRASTIVE__MOMENTUM=0.9, CONTRASTIVE__SEQUENTIAL=sequential)
    C, model = _model(cfg, seed=3)
    model._batch_shuffle_on = False
    W0 = model.backbone.proj.weight.detach().cpu().clone()
    queue0 = model.queue_x.cpu().clone()
    torch.manual_seed(4)
    xs = [torch.randn(B, D) for _ in range(2)]
    inputs = [[x.cuda()] for x in xs]
    time = torch.zeros(B, 2, 1).cuda()
    mdl, preds, partial_loss, perform_backward = C.contrastive_forward(
        model, cfg, inputs, torch.arange(B).cuda(), time, 0.0, None)
    # oracle replay
    hist = O.ema_update([W0], [torch.zeros_like(W0)], 0.9, 0)[0]
    keys = [O.l2_normalize(x @ hist.t()) for x in xs]
    if sequential:
        assert perform_backward is False
        tot = 0
        for k in range(2):
            _, _, l = O.moco_head(xs[k] @ W0.t(), keys[:k] + keys[k + 1:], queue0, T)
            tot = tot + l
        ref = tot / 4.0
        assert preds.shape == (2 * B, K + 1)
        assert model.backbone.proj.weight.grad is not None
        assert int(model.ptr.item()) == B  # keys[0] enqueued once at the end (:1166-1167)
    else:
        assert perform_backward is True
        _, _, ref = O.moco_head(xs[0] @ W0.t(), keys[1:], queue0, T)
        assert int(model.ptr.item()) == B
    assert rel_err(partial_loss, ref.detach()) < 5e-6


def test_parameter_surgery():
    cfg = make_cfg(CONTRASTIVE__TYPE="moco", CONTRASTIVE__DIM=16, CONTRASTIVE__QUEUE_LEN=256, TRAIN__BATCH_SIZE=64)
    C, model = _model(cfg)
    _, upd = C.contrastive_parameter_surgery(model, cfg, 0.1, 2)
    assert upd is False  # queue (256/64 = 4 iters) still filling
    _, upd = C.contrastive_parameter_surgery(model, cfg, 0.1, 4)
    assert upd is True


def test_state_dict_contract():
    cfg = make_cfg(CONTRASTIVE__TYPE="moco", CONTRASTIVE__DIM=16, CONTRASTIVE__QUEUE_LEN=64, CONTRASTIVE__KNN_ON=True,
                   CONTRASTIVE__LENGTH=10)
    C, model = _model(cfg)
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    for k, shape, dt in (("ptr", (1,), torch.int64), ("queue_x", (64, 16), torch.float32),
                         ("iter", (1,), torch.int64), ("knn_mem.memory", (10, 1, 16), torch.float32)):
        assert tuple(sd[k].shape) == shape and sd[k].dtype == dt
    assert "backbone.proj.weight" in sd and "backbone_hist.proj.weight" in sd
    assert not any(k.startswith("_status") for k in sd)
    assert all(not p.requires_grad for p in model.backbone_hist.parameters())
    # EMA pointer table survives load_state_dict (in-place copy) and is rebuilt after .to()
    x = torch.randn(4, 16).cuda()
    model([[x], [x]], torch.arange(4).cuda(), torch.zeros(4, 2, 1).cuda(), 0.0)
    plan = model._ema[2]
    model.load_state_dict(sd)
    model._update_history()
    assert model._ema[2] is plan and int(model.iter.item()) == 0
    model.cuda()
    assert model._ema is None
    np.testing.assert_equal(int(model.ptr.item()), 0)
    # the status word is a plain attribute that follows the module between devices
    assert model._status.is_cuda and "_status" not in dict(model.named_buffers())


def test_ema_plan_follows_rebound_parameter_storage():
    """The reference itself re-seats `p.data` every step (models/contrastive.py:172); a pointer table that
    survived such a rebind would read and write freed memory.  The table is re-validated on every call."""
    cfg = make_cfg(CONTRASTIVE__TYPE="moco", CONTRASTIVE__DIM=16, CONTRASTIVE__QUEUE_LEN=64, CONTRASTIVE__MOMENTUM=0.75)
    C, model = _model(cfg)
    model._update_history()           # iter == 0: copy, then blend
    plan = model._ema[2]
    w_on = model.backbone.proj.weight
    w_hi = model.backbone_hist.proj.weight
    keep_alive = w_hi.data            # noqa: F841  (the old storage stays allocated: a stale table would hit it)
    w_hi.data = w_hi.data.clone() * 0 + 3.0
    w_on.data = w_on.data.clone()
    with torch.no_grad():
        model.iter += 1
    model._update_history()
    assert model._ema[2] is not plan
    ref = w_on.detach().cpu() * (1.0 - 0.75) + torch.full_like(w_on.detach().cpu(), 3.0) * 0.75
    assert torch.equal(w_hi.detach().cpu(), ref)


@pytest.mark.parametrize("sweep_ctas", [None, 5])
@pytest.mark.parametrize("shuffle", [False, True])
def test_module_step_captures_into_a_cuda_graph(shuffle, sweep_ctas):
    """contrastive_forward + backward (the call tools/train.py makes) is capturable: no host synchronisation,
    allocation-stable, and replays reproduce the eager step (same kernels, same order)."""
    B, D, K, T = 32, 128, 1024, 0.1
    cfg = make_cfg(CONTRASTIVE__TYPE="moco", CONTRASTIVE__T=T, CONTRASTIVE__DIM=D, CONTRASTIVE__QUEUE_LEN=K,
                   CONTRASTIVE__MOMENTUM=0.99, MODEL__ARCH="identity")
    cfg.EMA_SHAPES = [(257, 33), (4096,), (128,)]
    C = register_backbones()
    torch.manual_seed(11)
    model = C.ContrastiveModel(cfg).cuda().train()
    model._batch_shuffle_on = shuffle
    ref = C.ContrastiveModel(cfg).cuda().train()
    ref._batch_shuffle_on = shuffle
    ref.load_state_dict(model.state_dict())
    model.sweep_ctas = ref.sweep_ctas = sweep_ctas  # 5: the two-launch head (sweep stream forked inside the capture)
    xq = torch.randn(B, D).cuda().requires_grad_(True)
    xk = torch.randn(B, D).cuda()
    index, time = torch.arange(B).cuda(), torch.zeros(B, 2, 1).cuda()

    def step(m, q):
        q.grad = None
        _, preds, loss, bwd = C.contrastive_forward(m, cfg, [[q], [xk]], index, time, 0.0)
        assert bwd is True
        loss.backward()
        return preds, loss

    for _ in range(3):  # warm-up: workspaces, exchange buffer, host mirror of `iter`
        step(model, xq)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        preds, loss = step(model, xq)
    n_replays = 4
    for _ in range(n_replays):
        graph.replay()
    torch.cuda.synchronize()
    # the same 3 + n_replays steps (capturing executes nothing), eagerly, on the twin
    xr = xq.detach().clone().requires_grad_(True)
    for _ in range(3 + n_replays):
        rp, rl = step(ref, xr)
    torch.cuda.synchronize()
    assert int(model.iter.item()) == int(ref.iter.item()) == 3 + n_replays
    assert torch.equal(model.ptr, ref.ptr)
    assert torch.equal(model.queue_x, ref.queue_x)   # keys are a row permutation away from the perm, not its values
    for a, b in zip(model.backbone_hist.parameters(), ref.backbone_hist.parameters()):
        assert torch.equal(a, b)
    assert torch.equal(loss, rl) and torch.equal(xq.grad, xr.grad) and torch.equal(preds, rp)
    assert model.check_device_status() == 0
