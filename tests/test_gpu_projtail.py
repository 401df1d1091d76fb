"""-m gpu parity tests of the projection tail: the projection MLP's last Linear with the
head's Normalize as its epilogue (`csrc/projtail.cu`, `advise_video_ssl_b200/head_helper.py`), through the C-ABI.

  * golden `projtail.npz`: the reference's own MLPHead + Normalize, forward and backward (make_golden_projtail.py);
  * the oracle (`O.projection_tail` + autograd) on seeded inputs at the BASELINE shapes and ragged ones;
  * structural properties: fused output == Normalize kernel applied to the fused plain Linear, bit for bit;
  * the ContrastiveModel step with a fused tail against the same model unfused.

Tolerances (fp32 accumulation in another order than the reference's sgemm): forward 2e-6 absolute on unit rows,
gradients 2e-5 relative in max-norm (`helpers.rel_err`); north_star allows 1e-3."""
import pytest
import torch
import torch.nn as nn

from helpers import make_cfg, register_backbones, rel_err
from oracle import contrastive_oracle as O

pytestmark = pytest.mark.gpu
FWD_ATOL = 2e-6
GRAD_RTOL = 2e-5


def _mlp(dim_in, dim_out, mlp_dim, layers, bn_on, bias):
    """Same layer sequence as the reference's MLPHead (models/head_helper.py:36-59) out of stock torch modules."""
    mods = [nn.Linear(dim_in, mlp_dim, bias=False if bn_on else bias)]
    for i in range(1, layers):
        if bn_on:
            mods.append(nn.BatchNorm1d(mlp_dim))
        mods.append(nn.ReLU(inplace=True))
        last = i == layers - 1
        mods.append(nn.Linear(mlp_dim, dim_out if last else mlp_dim, bias=bias if last else (False if bn_on else bias)))
    return nn.Sequential(*mods)


@pytest.mark.parametrize("case", ["a", "b"])
def test_projection_tail_golden(golden, case):
    from advise_video_ssl_b200 import head_helper as H
    g = golden("projtail")
    B, dim_in, mlp_dim, dim_out, layers, bn_on, bias = [int(v) for v in g[case + "_cfg"]]

    class Head(nn.Module):  # `.projection` as in MLPHead, so that state_dict keys line up
        def __init__(self):
            super().__init__()
            self.projection = _mlp(dim_in, dim_out, mlp_dim, layers, bool(bn_on), bool(bias))

        def forward(self, x):
            return self.projection(x)

    head = Head().train()
    sd = {k[len(case) + 4:]: g[k] for k in g.keys() if k.startswith(case + "_sd_")}
    head.load_state_dict(sd)  # strict: same parameter / buffer names as the reference's MLPHead
    head = head.cuda()
    keys_before = list(head.state_dict().keys())
    assert H.fuse_projection_tail(head) == 1 and H.fuse_projection_tail(head) == 0
    assert isinstance(head.projection[len(head.projection) - 1], H.LinearNormalize)
    assert list(head.state_dict().keys()) == keys_before
    h = g[case + "_h"].cuda().requires_grad_(True)
    q = head(h)
    (q * g[case + "_G"].cuda()).sum().backward()
    assert (q.detach().cpu() - g[case + "_q"]).abs().max().item() < FWD_ATOL
    assert rel_err(h.grad, g[case + "_dh"]) < GRAD_RTOL
    for name, p in head.named_parameters():
        assert rel_err(p.grad, g["%s_grad_%s" % (case, name)]) < GRAD_RTOL, name


@pytest.mark.parametrize("B,Kin,Dout,bias", [
    (64, 2048, 128, True),     # BASELINE configs[1]: SSL.MLP_DIM 2048 -> CONTRASTIVE.DIM 128
    (512, 2048, 256, True),    # configs[2] per rank
    (1, 4, 2, True), (9, 2052, 129, False), (130, 260, 100, True), (8, 512, 256, False), (65, 1028, 32, True)])
def test_projection_tail_against_oracle(B, Kin, Dout, bias):
    from advise_video_ssl_b200 import ops
    gen = torch.Generator().manual_seed(B * 7 + Kin + Dout)
    x = torch.randn(B, Kin, generator=gen).relu_()          # what the layer sees: BN + ReLU output
    W = torch.randn(Dout, Kin, generator=gen) / Kin ** 0.5
    b = torch.randn(Dout, generator=gen) * 0.1 if bias else None
    G = torch.randn(B, Dout, generator=gen)
    xr, Wr = x.clone().requires_grad_(True), W.clone().requires_grad_(True)
    br = b.clone().requires_grad_(True) if bias else None
    q_ref = O.projection_tail(xr, Wr, br)
    (q_ref * G).sum().backward()

    xd, Wd, bd = x.cuda(), W.cuda(), (b.cuda() if bias else None)
    q, nrm = ops.linear_l2norm_fwd(xd, Wd, bd)
    dx, dW, db = ops.linear_l2norm_bwd(xd, Wd, q, nrm, G.cuda(), need_db=bias)
    assert (q.cpu() - q_ref.detach()).abs().max().item() < FWD_ATOL
    y_ref = torch.nn.functional.linear(x, W, b)
    assert rel_err(nrm, y_ref.norm(dim=1)) < 2e-6
    assert rel_err(dx, xr.grad) < GRAD_RTOL and rel_err(dW, Wr.grad) < GRAD_RTOL
    if bias:
        assert rel_err(db, br.grad) < GRAD_RTOL
    else:
        assert db is None

    # structure: the epilogue IS the Normalize kernel (same bits as l2norm_fwd of the plain fused Linear), and the
    # backward's prologue IS l2norm_bwd (dx of the fused launch == plain-Linear backward fed with l2norm_bwd's dy)
    y, _ = ops.linear_l2norm_fwd(xd, Wd, bd, normalize=False)
    assert (y.cpu() - y_ref).abs().max().item() < 2e-5 * max(1.0, y_ref.abs().max().item())
    q2, nrm2 = ops.l2norm_fwd(y)
    assert torch.equal(q2, q) and torch.equal(nrm2, nrm)
    dy = ops.l2norm_bwd(q, nrm, G.cuda())
    dx2, dW2, db2 = ops.linear_l2norm_bwd(xd, Wd, q, nrm, dy, normalize=False, need_db=bias)
    assert rel_err(dx2, dx) < 1e-6 and rel_err(dW2, dW) < 1e-6 and (not bias or rel_err(db2, db) < 1e-6)
    # optional outputs
    only_dx = ops.linear_l2norm_bwd(xd, Wd, q, nrm, G.cuda(), need_dw=False, need_db=False)
    assert only_dx[1] is None and only_dx[2] is None and torch.equal(only_dx[0], dx)
    # deterministic: a second launch gives the same bits
    q3, _ = ops.linear_l2norm_fwd(xd, Wd, bd)
    assert torch.equal(q3, q)


def test_projection_tail_rejects_unsupported_shapes():
    from advise_video_ssl_b200 import _lib, head_helper as H, ops
    x, W = torch.randn(4, 6).cuda(), torch.randn(8, 6).cuda()
    with pytest.raises(_lib.AvsslError):
        ops.linear_l2norm_fwd(x, W)                      # Kin % 4 != 0
    with pytest.raises(ValueError):
        H.LinearNormalize(8, 300)
    seq = nn.Sequential(nn.Linear(8, 8), nn.ReLU(), nn.Linear(8, 300)).cuda()
    assert H.fuse_projection_tail(seq) == 0 and isinstance(seq[2], nn.Linear)   # left as it is
    with pytest.raises(RuntimeError):
        ops.linear_l2norm_fwd(torch.randn(4, 8), torch.randn(8, 8))  # CPU tensors: no fallback


class _HeadedBackbone(nn.Module):
    """Backbone stand-in with the reference's module layout: `.head.projection` = MLPHead-like `.projection` Sequential."""

    def __init__(self, cfg):
        super().__init__()
        d = cfg.CONTRASTIVE.DIM

        class MLP(nn.Module):
            def __init__(self):
                super().__init__()
                self.projection = _mlp(d, d, 64, 3, True, True)

            def forward(self, x):
                return self.projection(x)

        class Head(nn.Module):
            def __init__(self):
                super().__init__()
                self.projection = MLP()
                self.predictors = nn.ModuleList()

            def forward(self, x):
                return self.projection(x)

        self.head = Head()

    def forward(self, x):
        return self.head(x[0] if isinstance(x, (list, tuple)) else x)


@pytest.mark.parametrize("mode", ["moco", "simclr"])
def test_contrastive_model_with_fused_tail(mode):
    """The module step with the tails of both encoders fused == the same model unfused (loss 2e-6, parameter
    gradients 5e-4 in max-norm: the head's own tolerance), identical state_dict keys, the momentum update still
    bit-exact, the queue equal to 1 ulp."""
    from advise_video_ssl_b200 import head_helper as H
    C = register_backbones()
    C._MODEL_TYPES["headed"] = _HeadedBackbone
    B, D, K = 64, 128, 512
    cfg = make_cfg(CONTRASTIVE__TYPE=mode, CONTRASTIVE__T=0.1, CONTRASTIVE__DIM=D, CONTRASTIVE__QUEUE_LEN=K,
                   CONTRASTIVE__MOMENTUM=0.9, MODEL__ARCH="headed", TRAIN__BATCH_SIZE=B)
    torch.manual_seed(3)
    plain = C.ContrastiveModel(cfg).cuda().train()
    fused = C.ContrastiveModel(cfg).cuda().train()
    fused.load_state_dict(plain.state_dict())
    n = H.fuse_projection_tail(fused)
    assert n == (2 if mode == "moco" else 1)
    assert list(fused.state_dict().keys()) == list(plain.state_dict().keys())
    from advise_video_ssl_b200 import _lib
    for m in (plain, fused):
        m._batch_shuffle_on = False
        m.ntxent_impl = _lib.IMPL_SIMT  # exact-fp32 NT-Xent: the comparison then only sees the tail
    xq, xk = torch.randn(B, D).cuda(), torch.randn(B, D).cuda()
    index, time = torch.arange(B).cuda(), torch.zeros(B, 2, 1).cuda()
    for step in range(2):
        outs = []
        for m in (plain, fused):
            m.zero_grad()
            _, loss = m([[xq], [xk]], index, time, 0.0)
            loss.backward()
            outs.append(loss.detach())
        assert rel_err(outs[1], outs[0]) < 2e-6
        for (name, p), (_, pf) in zip(plain.backbone.named_parameters(), fused.backbone.named_parameters()):
            assert rel_err(pf.grad, p.grad) < 5e-4, name
        if mode == "moco":
            for (name, p), (_, pf) in zip(plain.backbone_hist.named_parameters(), fused.backbone_hist.named_parameters()):
                assert torch.equal(p, pf), name
            assert torch.equal(plain.ptr, fused.ptr)
            assert (plain.queue_x - fused.queue_x).abs().max().item() < 2e-7


def test_fused_tail_captures_into_a_cuda_graph():
    from advise_video_ssl_b200 import head_helper as H
    torch.manual_seed(0)
    tail = H.LinearNormalize(256, 128).cuda()
    x = torch.randn(64, 256).cuda().requires_grad_(True)
    G = torch.randn(64, 128).cuda()

    def step():
        tail.zero_grad(set_to_none=False)
        if x.grad is not None:
            x.grad.zero_()
        q = tail(x)
        (q * G).sum().backward()
        return q

    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(2):
            step()
    torch.cuda.current_stream().wait_stream(s)
    q_e = step().detach().clone()
    gx_e, gw_e = x.grad.clone(), tail.weight.grad.clone()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        q_g = step()
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(q_g, q_e) and torch.equal(x.grad, gx_e) and torch.equal(tail.weight.grad, gw_e)


@pytest.mark.parametrize("B", [128, 200])
def test_linear_normalize_module_both_backward_routes(B):
    """Up to 128 rows forward and backward are the fused launches; above, library GEMMs + the Normalize kernels
    (autograd.LinearNormalize.FUSED_MAX_ROWS).  Both against the oracle."""
    from advise_video_ssl_b200 import head_helper as H
    torch.manual_seed(B)
    tail = H.LinearNormalize(260, 96).cuda()
    x = torch.randn(B, 260).relu_()
    G = torch.randn(B, 96)
    xd = x.cuda().requires_grad_(True)
    (tail(xd) * G.cuda()).sum().backward()
    xr = x.clone().requires_grad_(True)
    Wr = tail.weight.detach().cpu().clone().requires_grad_(True)
    br = tail.bias.detach().cpu().clone().requires_grad_(True)
    (O.projection_tail(xr, Wr, br) * G).sum().backward()
    assert rel_err(xd.grad, xr.grad) < GRAD_RTOL and rel_err(tail.weight.grad, Wr.grad) < GRAD_RTOL
    assert rel_err(tail.bias.grad, br.grad) < GRAD_RTOL


def test_linear_normalize_leaves_other_inputs_to_the_plain_linear():
    """5-D inputs (fully-convolutional inference averages AFTER the projection) and non-fp32 inputs are projected
    without the epilogue: ContrastiveModel normalises the backbone output itself."""
    from advise_video_ssl_b200 import head_helper as H
    torch.manual_seed(0)
    tail = H.LinearNormalize(64, 32).cuda()
    x5 = torch.randn(2, 1, 3, 3, 64).cuda()
    ref = torch.nn.functional.linear(x5, tail.weight, tail.bias)
    assert torch.equal(tail(x5), ref)
    xh = torch.randn(4, 64).cuda().half()
    assert tail(xh).dtype == torch.float16
