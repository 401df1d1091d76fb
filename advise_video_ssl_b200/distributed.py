"""Collective helpers of the contrastive path, same names and semantics as the
reference's `utils/distributed.py:79-155` (plus the four functions it re-exports
from the un-vendored `pytorchvideo.layers.distributed`, utils/distributed.py:5-13).

B200 mapping: one NCCL call per collective into a preallocated
destination (`all_gather_into_tensor`, `reduce_scatter_tensor`, `all_to_all_single`)
instead of list-of-tensors gathers followed by `cat`; the process group is the one
the host application initialised (we never create or destroy it).
"""
import torch
import torch.distributed as dist

_LOCAL_PROCESS_GROUP = None  # aliased to WORLD by the reference (utils/distributed.py:66)


def _on():
    return dist.is_available() and dist.is_initialized()


def get_rank():
    """utils/distributed.py:79-87."""
    return dist.get_rank() if _on() else 0


def get_world_size():
    return dist.get_world_size() if _on() else 1


def get_local_size():
    """pytorchvideo get_local_size; the reference aliases the local group to WORLD
    (utils/distributed.py:66), so local == global on one box."""
    if not _on():
        return 1
    return dist.get_world_size(group=_LOCAL_PROCESS_GROUP)


def get_local_rank():
    if not _on():
        return 0
    return dist.get_rank(group=_LOCAL_PROCESS_GROUP)


def _backend_has_fused(t):
    return t.is_cuda


def _gather_cat(t, group=None):
    """cat(all_gather(t)) along dim 0 in one collective when the backend allows."""
    t = t.contiguous()
    ws = dist.get_world_size(group=group)
    out = torch.empty((ws * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    if _backend_has_fused(t):
        dist.all_gather_into_tensor(out, t, group=group)
    else:  # gloo (CPU unit tests): list gather into views of the same buffer
        dist.all_gather(list(out.chunk(ws, dim=0)), t, group=group)
    return out


def cat_all_gather(tensors, local=False):
    """pytorchvideo.layers.distributed.cat_all_gather: gather over WORLD (or the local
    group) and concatenate along dim 0."""
    if not _on():
        return tensors
    return _gather_cat(tensors, group=_LOCAL_PROCESS_GROUP if local else None)


def all_reduce(tensors, average=True):
    """utils/distributed.py:90-106 (in place, SUM then optional 1/world scaling)."""
    for t in tensors:
        dist.all_reduce(t, async_op=False)
    if average:
        ws = dist.get_world_size()
        for t in tensors:
            t.mul_(1.0 / ws)
    return tensors


def all_gather(tensors):
    """utils/distributed.py:109-128: gather each tensor from every rank, cat on dim 0."""
    return [_gather_cat(t) for t in tensors]


class AllGatherWithGradient(torch.autograd.Function):
    """utils/distributed.py:131-155.  Forward: all_gather + cat.  Backward: the
    reference all-reduces (SUM) the full gathered gradient and keeps this rank's
    slice, i.e. a reduce-scatter — issued as one here."""

    @staticmethod
    def forward(ctx, input):
        return _gather_cat(input)

    @staticmethod
    def backward(ctx, grad_output):
        ws = dist.get_world_size()
        grad_output = grad_output.contiguous()
        mb = grad_output.size(0) // ws
        if _backend_has_fused(grad_output):
            out = torch.empty((mb,) + tuple(grad_output.shape[1:]), dtype=grad_output.dtype,
                              device=grad_output.device)
            dist.reduce_scatter_tensor(out, grad_output, op=dist.ReduceOp.SUM)
            return out
        dist.all_reduce(grad_output)
        r = dist.get_rank()
        return grad_output[r * mb:(r + 1) * mb]
