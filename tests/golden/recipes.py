"""Seeded recipes for inputs too large to commit (shared by make_golden.py and tests).

torch's CPU generator is deterministic for a given torch build; the GPU box runs
the same image, and `moco_cfg1.npz` stores a checksum of the regenerated queue so
a drift would be caught rather than silently compared.
"""
import math

import torch


def uniform_queue(K, D, seed):
    """Same distribution as the reference queue init (models/contrastive.py:85-89):
    U(-s, s), s = 1/sqrt(D/3)."""
    g = torch.Generator().manual_seed(seed)
    stdv = 1.0 / math.sqrt(D / 3)
    return torch.rand(K, D, generator=g).mul_(2 * stdv).add_(-stdv)


def moco_cfg1():
    """BASELINE.json configs[0]: MoCo head, B=64, D=128, K=65536, T=0.1, world 1."""
    B, D, K, T, m = 64, 128, 65536, 0.1, 0.999
    g = torch.Generator().manual_seed(0)
    xq = torch.randn(B, D, generator=g)
    xk = torch.randn(B, D, generator=g)
    W = torch.randn(D, D, generator=g) / math.sqrt(D)
    return dict(B=B, D=D, K=K, T=T, m=m, xq=xq, xk=xk, W=W, queue=uniform_queue(K, D, 1234))
