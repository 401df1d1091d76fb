"""CPU-only tests of the host-side logic: EMA chunk planning, the fp32 arithmetic
specification of the EMA kernel against the reference's golden vectors, swav pair
weights, cfg-driven module construction."""
import numpy as np
import pytest
import torch

from helpers import make_cfg, register_backbones


def test_ema_plan_table():
    from advise_video_ssl_b200 import ops
    numels = [1, 4096, 4097, 0, 10000]
    base_o, base_h = 0x10000000, 0x20000000
    optr, hptr, off = [], [], 0
    for n in numels:
        optr.append(base_o + 4 * off)
        hptr.append(base_h + 4 * off)
        off += n
    t = ops.ema_plan_table(optr, hptr, numels)
    assert len(t) == 1 + 1 + 2 + 0 + 3
    assert int(t["n"].sum()) == sum(numels)
    assert list(t["n"]) == [1, 4096, 4096, 1, 4096, 4096, 1808]
    # chunk k of a tensor starts 16 KiB after chunk k-1, in both arrays
    assert t["online"][1] - t["online"][0] == 4 and t["online"][3] - t["online"][2] == 4096 * 4
    assert t["hist"][5] - t["hist"][4] == 4096 * 4
    # alignment flag: tensor 0 is 16-byte aligned; tensor 1 starts 4 bytes later
    assert t["flags"][0] == 1 and t["flags"][1] == 0
    with pytest.raises(Exception):
        ops.ema_plan_table([base_o + 1], [base_h], [8])  # not 4-byte aligned


def test_ema_arithmetic_spec_matches_reference(golden):
    """The kernel computes fadd(fmul(o, f32(1-m)), fmul(h, f32(m))) with the scalars
    rounded once from the host doubles; emulate that in numpy float32 and compare
    bit-for-bit with what the reference's ATen ops produced."""
    g = golden("ema")
    names = list(g["names"])
    hist = {n: g["hist_init/" + n].numpy() for n in names}
    for s in range(len(g["epochs"])):
        m = g.scalar("mmt%d" % s)
        mf, omf = np.float32(m), np.float32(1.0 - m)
        for n in names:
            o = g["online%d/%s" % (s, n)].numpy()
            h = o if s == 0 else hist[n]
            new = (o * omf).astype(np.float32) + (h * mf).astype(np.float32)
            assert np.array_equal(new, g["hist%d/%s" % (s, n)].numpy()), (s, n)
            hist[n] = new


def test_swav_pair_weights():
    from advise_video_ssl_b200 import ops
    w = ops.swav_pair_weights(6, 2, 256)
    assert w.shape == (2, 6) and w[0, 0] == 0 and w[1, 1] == 0
    assert np.isclose(w.sum(), 1.0 / 256)
    assert np.allclose(w[0, 1:], 1.0 / (256 * 5 * 2))


def test_module_builds_on_cpu_and_rejects_cpu_compute():
    C = register_backbones()
    for typ, extra in (("moco", {}), ("byol", {"CONTRASTIVE__PREDICTOR_DEPTHS": [1]}), ("swav", {}), ("simclr", {}),
                       ("mem", {"CONTRASTIVE__LENGTH": 20})):
        cfg = make_cfg(CONTRASTIVE__TYPE=typ, CONTRASTIVE__DIM=16, CONTRASTIVE__QUEUE_LEN=32, **extra)
        m = C.ContrastiveModel(cfg)
        assert m.type == typ
    cfg = make_cfg(CONTRASTIVE__TYPE="moco", CONTRASTIVE__DIM=16, CONTRASTIVE__QUEUE_LEN=32)
    m = C.ContrastiveModel(cfg).train()
    assert m.ptr.dtype == torch.int64 and m.iter.dtype == torch.int64 and m.queue_x.shape == (32, 16)
    assert m._batch_shuffle_on
    cfg2 = make_cfg(CONTRASTIVE__TYPE="moco", CONTRASTIVE__DIM=16, CONTRASTIVE__QUEUE_LEN=32,
                    BN__NORM_TYPE="sync_batchnorm", BN__NUM_SYNC_DEVICES=1)
    assert not C.ContrastiveModel(cfg2)._batch_shuffle_on  # sync BN over all GPUs (:91-99)
    if not torch.cuda.is_available():
        x = torch.randn(4, 16)
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            m([[x], [x]], torch.arange(4), torch.zeros(4, 2, 1), 0.0)
    with pytest.raises(NotImplementedError):
        bad = C.ContrastiveModel(make_cfg(CONTRASTIVE__TYPE="self", CONTRASTIVE__DIM=16, CONTRASTIVE__QUEUE_LEN=32))
        bad([[torch.randn(2, 16)]], torch.arange(2))


def test_momentum_anneal_matches_oracle():
    from oracle import contrastive_oracle as O
    C = register_backbones()
    cfg = make_cfg(CONTRASTIVE__TYPE="moco", CONTRASTIVE__DIM=16, CONTRASTIVE__QUEUE_LEN=32, CONTRASTIVE__MOMENTUM=0.99,
                   CONTRASTIVE__MOMENTUM_ANNEALING=True, SOLVER__MAX_EPOCH=200)
    m = C.ContrastiveModel(cfg)
    for ep in (0.0, 1.5, 77.7, 200.0):
        m.momentum_anneal_cosine(ep)
        assert m.mmt == O.momentum_cosine(0.99, ep, 200)


def test_queue_divisibility_assert():
    """K % n == 0 stays a host AssertionError (models/contrastive.py:284)."""
    C = register_backbones()
    cfg = make_cfg(CONTRASTIVE__TYPE="moco", CONTRASTIVE__DIM=16, CONTRASTIVE__QUEUE_LEN=32)
    m = C.ContrastiveModel(cfg)
    with pytest.raises(AssertionError):
        m._dequeue_and_enqueue([torch.randn(5, 16)])


# ------------------------------------------------- round-1 additions: no silent CPU paths
@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_new_entry_points_have_no_cpu_fallback():
    from advise_video_ssl_b200 import ops, optimizer, temporal
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.PeerExchange(8, 32)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.MultiTensorNorm([torch.randn(16)])
    p = torch.nn.Parameter(torch.zeros(16))
    p.grad = torch.randn(16)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        optimizer.get_grad_norm_([p])                       # the 2-norm is the CUDA path, never a torch fallback
    assert optimizer.get_grad_norm_([torch.nn.Parameter(torch.zeros(3))]).item() == 0.0   # no grads: 0.0 (:380-381)

    class Holder(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.temporal_encoder = torch.nn.Linear(4, 4)
            self.temporal_encoder_hist = torch.nn.Linear(4, 4)
            self.head_projector = torch.nn.Linear(4, 2)
            self.head_projector_hist = torch.nn.Linear(4, 2)
            self.mmt, self.T = 0.99, 0.1

    h = Holder()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        temporal.update_history(h)


def test_momentum_pairs_follow_reference_order():
    """temporal._pairs walks temporal_encoder_hist then head_projector_hist by NAME
    (models/temporal_modeling.py:220-237), whatever the registration order of the online modules."""
    from advise_video_ssl_b200 import temporal

    class Holder(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.head_projector = torch.nn.Linear(4, 2)
            self.temporal_encoder = torch.nn.Sequential(torch.nn.Linear(4, 4), torch.nn.LayerNorm(4))
            self.temporal_encoder_hist = torch.nn.Sequential(torch.nn.Linear(4, 4), torch.nn.LayerNorm(4))
            self.head_projector_hist = torch.nn.Linear(4, 2)

    h = Holder()
    online, hist = temporal._pairs(h)
    names = [n for n, _ in h.temporal_encoder_hist.named_parameters()] + [n for n, _ in h.head_projector_hist.named_parameters()]
    assert len(online) == len(hist) == len(names) == 6
    enc, proj = dict(h.temporal_encoder.named_parameters()), dict(h.head_projector.named_parameters())
    for i, n in enumerate(names):
        src = enc[n] if i < 4 else proj[n]
        assert online[i].data_ptr() == src.data_ptr() and online[i].shape == hist[i].shape


def test_state_dict_contract_matches_reference():
    """Checkpoint contract (utils/misc.py:118-137,300-339): names, shapes, dtypes of every state_dict
    entry and the frozen-parameter set, in every mode, against the unmodified reference
    (tests/golden/state_dict_contract.json, generated by make_golden_statedict.py)."""
    import json
    import os
    C = register_backbones()
    path = os.path.join(os.path.dirname(__file__), "golden", "state_dict_contract.json")
    ref = json.load(open(path))
    assert set(ref) == {"moco", "moco_knn", "byol", "swav", "swav_queue", "simclr", "mem_1d", "mem_2d"}
    for name, case in ref.items():
        cfg = make_cfg(**case["cfg"])
        torch.manual_seed(0)
        m = C.ContrastiveModel(cfg)
        mine = sorted([k, list(v.shape), str(v.dtype)] for k, v in m.state_dict().items())
        assert mine == case["state_dict"], (name, [x for x in mine if x not in case["state_dict"]],
                                            [x for x in case["state_dict"] if x not in mine])
        frozen = sorted(n for n, p in m.named_parameters() if not p.requires_grad)
        assert frozen == case["frozen"], name


# ------------------------------------------------------------------ shuffle plan (A6, host side)
@pytest.mark.parametrize("world,bsz", [(1, 7), (2, 5), (4, 3), (8, 64)])
def test_shuffle_plan_matches_gather_then_select(world, bsz):
    """The all-to-all lists of every rank, replayed in-process, reproduce the reference's
    cat_all_gather(x)[perm.view(W, -1)[rank]] (models/contrastive.py:186-207) and its idx_restore."""
    import numpy as np
    from advise_video_ssl_b200.shuffle import ShufflePlan
    from oracle import contrastive_oracle as O
    g = torch.Generator().manual_seed(world * 100 + bsz)
    parts = [torch.randn(bsz, 3, generator=g) for _ in range(world)]
    perm = torch.randperm(world * bsz, generator=g)
    ref_x, ref_restore = O.shuffle_emulated(parts, perm)
    cpu = torch.device("cpu")
    plans = [ShufflePlan(perm.numpy(), world, r, bsz, cpu) for r in range(world)]
    for r, plan in enumerate(plans):
        assert torch.equal(plan.restore, ref_restore) and plan.restore.dtype == torch.int64
        if world == 1:
            assert torch.equal(plan.shuffled(parts[0]), ref_x[0])
            continue
        assert sum(plan.send_counts) == bsz == sum(plan.recv_counts)
        # what rank r receives: from every source s, s's block destined to r
        recv = []
        for s_rank, sp in enumerate(plans):
            off = int(np.sum(sp.send_counts[:r]))
            assert sp.send_counts[r] == plan.recv_counts[s_rank]
            recv.append(parts[s_rank].index_select(0, sp.send_rows)[off:off + sp.send_counts[r]])
        got = torch.cat(recv).index_select(0, plan.place)
        assert torch.equal(got, ref_x[r])


# ------------------------------------------------------------------ projection tail, host side
def test_fuse_projection_tail_swaps_only_the_last_linear_and_shares_parameters():
    import torch.nn as nn
    from advise_video_ssl_b200 import head_helper as H

    class MLP(nn.Module):  # layout of the reference's MLPHead (models/head_helper.py:36-59)
        def __init__(self, dout):
            super().__init__()
            self.projection = nn.Sequential(nn.Linear(32, 64, bias=False), nn.BatchNorm1d(64), nn.ReLU(inplace=True),
                                            nn.Linear(64, dout))

    class Head(nn.Module):  # ... and of ResNetBasicHead (:135-182)
        def __init__(self, dout, n_pred=0):
            super().__init__()
            self.projection = MLP(dout)
            self.predictors = nn.ModuleList([MLP(dout) for _ in range(n_pred)])

    class Net(nn.Module):
        def __init__(self, **kw):
            super().__init__()
            self.stem = nn.Linear(8, 32)
            self.head = Head(**kw)

    class Pair(nn.Module):
        def __init__(self):
            super().__init__()
            self.backbone, self.backbone_hist = Net(dout=128), Net(dout=128)

    pair = Pair()
    keys = list(pair.state_dict().keys())
    params = [id(p) for p in pair.parameters()]
    last = pair.backbone.head.projection.projection[3]
    assert H.fuse_projection_tail(pair) == 2
    new = pair.backbone.head.projection.projection[3]
    assert isinstance(new, H.LinearNormalize) and new.weight is last.weight and new.bias is last.bias
    assert isinstance(pair.backbone.head.projection.projection[0], nn.Linear)       # only the tail
    assert list(pair.state_dict().keys()) == keys and [id(p) for p in pair.parameters()] == params
    assert H.fuse_projection_tail(pair) == 0                                         # idempotent
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        new(torch.randn(4, 64))                                                      # the kernel is the only path
    # left alone: heads with predictors (BYOL), outputs wider than the kernel takes, in_features % 4 != 0
    assert H.fuse_projection_tail(Net(dout=128, n_pred=1)) == 0
    assert H.fuse_projection_tail(Net(dout=300)) == 0
    assert H.fuse_projection_tail(nn.Sequential(nn.Linear(8, 6), nn.ReLU(), nn.Linear(6, 16))) == 0
    # a bare Sequential and a single-Linear projection (SSL.NUM_MLP_LAYERS == 1, :135-136) are recognised
    seq = nn.Sequential(nn.Linear(8, 16), nn.ReLU(), nn.Linear(16, 32, bias=False))
    assert H.fuse_projection_tail(seq) == 1 and seq[2].bias is None and "2.bias" not in seq.state_dict()

    class OneLayerHead(nn.Module):
        def __init__(self):
            super().__init__()
            self.projection = nn.Linear(16, 128)
    one = OneLayerHead()
    assert H.fuse_projection_tail(one) == 1 and isinstance(one.projection, H.LinearNormalize)
    with pytest.raises(ValueError):
        H.LinearNormalize(16, 512)


def test_topk_plan_limits():
    """avssl_topk_rows_workspace_bytes doubles as the capability query of the kNN top-k: k <= min(M, 1024), any M."""
    from advise_video_ssl_b200._lib import lib
    assert lib.avssl_topk_rows_workspace_bytes(64, 239975, 200) > 0
    assert lib.avssl_topk_rows_workspace_bytes(1, 10_000_000, 1024) > 0
    assert lib.avssl_topk_rows_workspace_bytes(64, 239975, 1025) == 0     # beyond the candidate lists
    assert lib.avssl_topk_rows_workspace_bytes(64, 100, 101) == 0         # k > M: torch.topk raises as well
    assert lib.avssl_topk_rows_workspace_bytes(0, 100, 10) == 0
    # workspace = N x G x KP composites, G <= 32 lists
    assert lib.avssl_topk_rows_workspace_bytes(1, 10_000_000, 200) <= 1 * 32 * 256 * 8
