"""Autograd bridges between torch's tape and the C-ABI kernels.

Every kernel on the contrastive path produces its gradient in the forward launch
(the loss is a scalar, so d loss / d input is known up to the upstream factor);
`backward` only scales the stored gradient.  Inputs are detached before they reach
the library: torch sees one opaque node per fused op.
"""
import torch

from . import ops


def _scaled(grad, g):
    """grad * upstream factor.  The factor is a 0-dim tensor (1.0 from `loss.backward()`,
    the loss scale under AMP); one elementwise launch."""
    return grad * g


class L2NormRows(torch.autograd.Function):
    """y = x / max(||x||, eps) for every row of a [n, D] tensor (K2,
    models/contrastive.py:923-934 with eps = 0, F.normalize with eps = 1e-12)."""

    @staticmethod
    def forward(ctx, x, eps):
        y, nrm = ops.l2norm_fwd(x.detach().contiguous(), eps)
        ctx.save_for_backward(y, nrm)
        ctx.eps = eps
        return y

    @staticmethod
    def backward(ctx, dy):
        y, nrm = ctx.saved_tensors
        return ops.l2norm_bwd(y, nrm, dy.contiguous(), ctx.eps), None


def l2norm_lastdim(x, eps=0.0):
    """Normalise the last dimension of a tensor with >= 2 dims through the row kernel."""
    shape = x.shape
    return L2NormRows.apply(x.reshape(-1, shape[-1]), eps).reshape(shape)


class LinearNormalize(torch.autograd.Function):
    """q = Normalize(x W^T + b): the projection MLP's last Linear (models/head_helper.py:52-58) with the head's
    Normalize (models/contrastive.py:923-934) as its epilogue; one launch forward, one launch backward."""

    # The fused launches are a latency play: exact fp32 on the CUDA cores, split-K over a cluster forward, every CTA
    # re-deriving dy backward.  Measured on B200 (profiles/r2_next_bench.jsonl): they win at the head's batch (B = 64,
    # Kin 2048 -> 128: 10.7 / 12.6 us against 27.0 / 24.9 us for cuBLAS + the Normalize kernels) and up to 128 rows;
    # at B = 512 the GEMMs are flop-bound and belong to the library (cuBLAS, then the Normalize kernels).
    FUSED_MAX_ROWS = 128

    @staticmethod
    def forward(ctx, x, weight, bias, eps, normalize):
        xd, wd = x.detach().contiguous(), weight.detach().contiguous()
        bd = None if bias is None else bias.detach().contiguous()
        if xd.shape[0] <= LinearNormalize.FUSED_MAX_ROWS:
            q, nrm = ops.linear_l2norm_fwd(xd, wd, bd, eps, normalize)
        else:
            y = torch.nn.functional.linear(xd, wd, bd)
            q, nrm = ops.l2norm_fwd(y, eps) if normalize else (y, y.new_empty(y.shape[0]))
        ctx.save_for_backward(xd, wd, q, nrm)
        ctx.eps, ctx.normalize, ctx.has_bias = eps, normalize, bias is not None
        return q

    @staticmethod
    def backward(ctx, grad_q):
        x, w, q, nrm = ctx.saved_tensors
        need = ctx.needs_input_grad
        want_db = ctx.has_bias and need[2]
        if x.shape[0] <= LinearNormalize.FUSED_MAX_ROWS:
            dx, dw, db = ops.linear_l2norm_bwd(x, w, q, nrm, grad_q, ctx.eps, ctx.normalize, need_dx=need[0],
                                               need_dw=need[1], need_db=want_db)
            return dx, dw, db, None, None
        dy = ops.l2norm_bwd(q, nrm, grad_q, ctx.eps) if ctx.normalize else grad_q.contiguous()
        return (dy @ w if need[0] else None, dy.t() @ x if need[1] else None, dy.sum(0) if want_db else None, None, None)


class MocoInfoNce(torch.autograd.Function):
    """K2 + K3 (+ K4, + C3 wait): q = f/||f||, logits against [key; queue], InfoNCE, d loss / d f
    (models/contrastive.py:462-503, models/losses.py:20-25).

    `plan` is a dict of keyword arguments for `ops.moco_infonce` (keys or peer exchange,
    enqueue targets, implementation, output tensors); outputs: loss (0-dim, differentiable in
    `feat_q`), logits ([n_keys*B, K+1] or an empty tensor) and q ([B, D]), both detached.
    """

    @staticmethod
    def forward(ctx, feat_q, queue, T, plan):
        res = ops.moco_infonce(feat_q.detach().contiguous(), plan.get("keys"), queue, T, **plan["kw"])
        ctx.save_for_backward(res["dfeat"])
        ctx.set_materialize_grads(False)  # no zero-filled [n_keys*B, K+1] gradient for the detached logits
        logits = res["logits"] if res["logits"] is not None else feat_q.new_empty(0)
        ctx.mark_non_differentiable(logits, res["q"])
        return res["loss"].reshape(()), logits, res["q"]

    @staticmethod
    def backward(ctx, g_loss, _g_logits, _g_q):
        (dfeat,) = ctx.saved_tensors
        return (None if g_loss is None else _scaled(dfeat, g_loss)), None, None, None


class ByolSimilarity(torch.autograd.Function):
    """-mean_n(p_n . k_n) / T, optionally with p = pred / ||pred|| fused in (K7,
    models/contrastive.py:243-249 and :533)."""

    @staticmethod
    def forward(ctx, pred, key, T, normalize):
        loss, dpred = ops.byol_simloss(pred.detach().contiguous(), key.detach().contiguous(), T,
                                       normalize=normalize, want_grad=True)
        ctx.save_for_backward(dpred)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        (dpred,) = ctx.saved_tensors
        return _scaled(dpred, g), None, None, None


class NtXentRows(torch.autograd.Function):
    """SimCLR NT-Xent over this rank's 2B rows against all 2N gathered columns (K6 with C4 / C5,
    models/contrastive.py:770-792, utils/distributed.py:131-155)."""

    @staticmethod
    def forward(ctx, feat1, feat2, T, impl, xchg=None, status=None):
        loss, d1, d2 = ops.ntxent(feat1.detach().contiguous(), feat2.detach().contiguous(), T, impl=impl, xchg=xchg,
                                  status=status)
        ctx.save_for_backward(d1, d2)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        d1, d2 = ctx.saved_tensors
        return _scaled(d1, g), _scaled(d2, g), None, None, None, None


class SwavSwappedCe(torch.autograd.Function):
    """SwAV swapped-prediction cross-entropy over every (assign crop, other crop) pair (K11,
    models/contrastive.py:672-679)."""

    @staticmethod
    def forward(ctx, scores, codes, n_crops, bs, T):
        loss, dscores = ops.swav_ce(scores.detach().contiguous(), codes, n_crops, bs, T)
        ctx.save_for_backward(dscores)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        (dscores,) = ctx.saved_tensors
        return _scaled(dscores, g), None, None, None, None


class BankDot(torch.autograd.Function):
    """prod[n, k] = q_n . bank[ind[n, k], time[n, k]] / T without the [B, K+1, D] gather (K14,
    models/contrastive.py:399-434).  The reference throws the mem-mode loss away (:436, :442), so
    the backward is a cold path: torch gather + einsum."""

    @staticmethod
    def forward(ctx, q, bank, ind, time, T, interp, status):
        prod = ops.membank_gather_dot(bank, q.detach().contiguous(), ind, time, T, interp=interp, status=status)
        ctx.save_for_backward(bank, ind, time)
        ctx.T, ctx.interp = T, interp
        return prod

    @staticmethod
    def backward(ctx, g):
        bank, ind, time = ctx.saved_tensors
        rows3 = bank if bank.dim() == 3 else bank.unsqueeze(1)
        last = rows3.shape[1] - 1
        flat_ind = ind.reshape(-1)
        if ctx.interp:
            lo = time.floor().long().clamp(0, last)
            hi = (lo + 1).clamp(0, last)
            w_hi = 1 - (time - lo).reshape(-1, 1).float()  # the reference's "hack for inverse" (:980)
            rows = rows3[flat_ind, lo.reshape(-1)] * (1 - w_hi) + rows3[flat_ind, hi.reshape(-1)] * w_hi
        else:
            rows = rows3[flat_ind, time.long().reshape(-1)]
        rows = rows.view(ind.shape[0], -1, rows3.shape[-1])
        return torch.einsum("nk,nkc->nc", g, rows) / ctx.T, None, None, None, None, None, None
