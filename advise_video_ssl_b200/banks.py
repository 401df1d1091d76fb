"""`Normalize`, `Memory` and `Memory1D` of the reference's `models/contrastive.py`
(:923-934, :937-1039, :1042-1080) on the row kernels (K2) and the fused bank kernels (K5).

Same constructor arguments, buffer name (`memory`), shapes and initial distribution as the
reference, so checkpoints and the RNG stream at construction time are interchangeable.
"""
import math

import torch
import torch.nn as nn

from . import distributed as du
from . import ops
from .autograd import l2norm_lastdim


def _bank_init(*shape):
    """U(-s, s) with s = 1/sqrt(dim/3): the reference's bank / queue initialisation."""
    s = 1.0 / math.sqrt(shape[-1] / 3)
    return torch.rand(*shape).mul_(2 * s).add_(-s)


class Normalize(nn.Module):
    """x / (sum_dim x^power)^(1/power); only the power the reference ever uses (2) has a kernel.
    No epsilon, exactly like models/contrastive.py:929-934."""

    def __init__(self, power=2, dim=1):
        super(Normalize, self).__init__()
        self.dim = dim
        self.power = power

    def forward(self, x):
        if self.power != 2:
            raise NotImplementedError("Normalize: only power=2 has a CUDA kernel")
        axis = self.dim % x.dim()
        if axis == x.dim() - 1:
            return l2norm_lastdim(x, 0.0)
        moved = x.movedim(axis, -1).contiguous()
        return l2norm_lastdim(moved, 0.0).movedim(-1, axis)


def _gathered(num_gpus, mem, ind, time):
    """C8: every rank applies every rank's update, in rank order (utils/distributed.py:109-128)."""
    if num_gpus > 1:
        return du.all_gather([mem, ind, time])
    return mem, ind, time


class Memory(nn.Module):
    """[length, duration, dim] bank with an optional time axis."""

    def __init__(self, length, duration, dim, cfg):
        super(Memory, self).__init__()
        self.length, self.duration, self.dim = length, duration, dim
        self.register_buffer("memory", _bank_init(length, duration, dim))
        self.device = self.memory.device
        self.l2_norm = Normalize(dim=1)
        self.l2_norm2d = Normalize(dim=2)
        self.num_gpus = cfg.NUM_GPUS

    def resize(self, length, duration, dim):
        """Re-draw the bank with a new geometry on the device it lives on (:954-965)."""
        where = self.memory.device
        self.length, self.duration, self.dim = length, duration, dim
        del self.memory
        self.memory = _bank_init(length, duration, dim).to(where)

    @torch.no_grad()
    def get(self, ind, time, interp=False):
        """bank[ind, time] as [batch, -1, dim]; with `interp` the blend of the two neighbouring
        time slots, weights as in the reference (its `1 - frac` lands on the later slot, :976-981)."""
        bank = self.memory
        rows = ind.reshape(-1)
        if not interp:
            picked = bank[rows, time.long().reshape(-1), :]
        else:
            last = bank.shape[1] - 1
            lo = time.floor().long().clamp_(0, last)
            hi = (lo + 1).clamp_(0, last)
            w_hi = 1 - (time - lo).reshape(-1, 1).float()
            picked = bank[rows, lo.reshape(-1), :] * (1 - w_hi) + bank[rows, hi.reshape(-1), :] * w_hi
        return picked.view(ind.size(0), -1, self.dim)

    @torch.no_grad()
    def update(self, mem, momentum, ind, time, interp=False, status=None):
        """bank[ind, time] <- l2norm(mem * m + old * (1 - m)) (K5), after the cross-rank gather."""
        mem, ind, time = _gathered(self.num_gpus, mem, ind, time)
        ops.membank_update(self.memory, mem.detach().reshape(mem.size(0), -1).contiguous(), ind, time, momentum,
                           interp=interp, status=status)

    def forward(self, inputs):
        pass


class Memory1D(nn.Module):
    """[length, dim] bank (duration must be 1)."""

    def __init__(self, length, duration, dim, cfg):
        super(Memory1D, self).__init__()
        assert duration == 1
        self.length, self.duration, self.dim = length, duration, dim
        self.register_buffer("memory", _bank_init(length, dim))
        self.l2_norm = Normalize(dim=1)
        self.num_gpus = cfg.NUM_GPUS

    @torch.no_grad()
    def get(self, ind, time, interp=False):
        picked = self.memory.index_select(0, ind.reshape(-1))
        if ind.dim() == 1:
            return picked.view(ind.size(0), self.dim)
        return picked.view(ind.size(0), -1, self.dim)

    @torch.no_grad()
    def update(self, mem, momentum, ind, time, interp=False, status=None):
        mem, ind, time = _gathered(self.num_gpus, mem, ind, time)
        ops.membank_update(self.memory, mem.detach().reshape(mem.size(0), -1).contiguous(), ind.long(), None, momentum,
                           interp=False, status=status)
