"""`ContrastiveLoss` / `get_loss_func("contrastive_loss")` — the part of the
reference's `models/losses.py` (:15-25, :137, :144-152) that sits on the contrastive
path, backed by the CUDA cross-entropy-against-class-0 kernels."""
import torch
import torch.nn as nn

from . import ops


class _CeTarget0Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits):
        lg = logits.detach().contiguous()
        loss, lse = ops.ce_target0_fwd(lg)
        ctx.save_for_backward(lg, lse)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        lg, lse = ctx.saved_tensors
        return ops.ce_target0_bwd(lg, lse, g)


class ContrastiveLoss(nn.Module):
    """models/losses.py:15-25: CrossEntropyLoss(inputs, zeros), i.e. InfoNCE with the
    positive in column 0."""

    def __init__(self, reduction="mean"):
        super(ContrastiveLoss, self).__init__()
        if reduction != "mean":
            raise NotImplementedError("ContrastiveLoss: only reduction='mean' has a CUDA kernel")
        self.reduction = reduction

    def forward(self, inputs, dummy_labels=None):
        return _CeTarget0Fn.apply(inputs)


_LOSSES = {"contrastive_loss": ContrastiveLoss}


def get_loss_func(loss_name):
    """models/losses.py:144-152 (only the contrastive entry lives in this package)."""
    if loss_name not in _LOSSES.keys():
        raise NotImplementedError("Loss {} is not supported".format(loss_name))
    return _LOSSES[loss_name]
