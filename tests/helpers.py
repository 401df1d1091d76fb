"""Shared test helpers: a cfg with the reference's defaults for the keys the
contrastive path reads (configs/defaults.py:87-158 and friends) and a stub backbone
identical to the one the golden vectors were generated with (tests/golden/ref_shim.py)."""
import math
from types import SimpleNamespace

import torch
import torch.nn as nn


class Node(SimpleNamespace):
    def __getitem__(self, k):
        return getattr(self, k)


def make_cfg(**over):
    cfg = Node(
        NUM_GPUS=1, NUM_SHARDS=1, SHARD_ID=0,
        MODEL=Node(MODEL_NAME="ContrastiveModel", ARCH="stub"),
        BN=Node(NORM_TYPE="batchnorm", NUM_SYNC_DEVICES=1),
        DATA=Node(TRAIN_CROP_NUM_TEMPORAL=2, TRAIN_CROP_NUM_SPATIAL=1),
        SOLVER=Node(MAX_EPOCH=300),
        TRAIN=Node(BATCH_SIZE=64),
        CONTRASTIVE=Node(
            T=0.07, DIM=128, HIDDEN_DIM=4096, LENGTH=239975, QUEUE_LEN=65536, MOMENTUM=0.5,
            MOMENTUM_ANNEALING=False, TYPE="mem", INTERP_MEMORY=False, MEM_TYPE="1d",
            LOCAL_SHUFFLE_BN=True, MOCO_MULTI_VIEW_QUEUE=False, PREDICTOR_DEPTHS=[],
            SEQUENTIAL=False, SIMCLR_DIST_ON=True, SWAV_QEUE_LEN=0, KNN_ON=False),
    )
    for k, v in over.items():
        node = cfg
        parts = k.split("__")
        for p in parts[:-1]:
            node = getattr(node, p)
        setattr(node, parts[-1], v)
    return cfg


class StubBackbone(nn.Module):
    """forward([x]) -> x @ W^T, or [feat, pred] when PREDICTOR_DEPTHS is non-empty."""

    def __init__(self, cfg):
        super().__init__()
        d = cfg.CONTRASTIVE.DIM
        self.proj = nn.Linear(getattr(cfg, "STUB_IN_DIM", d), d, bias=False)
        self.predictor = nn.Linear(d, d, bias=True) if len(cfg.CONTRASTIVE.PREDICTOR_DEPTHS) > 0 else None

    def forward(self, x):
        if isinstance(x, (list, tuple)):
            x = x[0]
        f = self.proj(x)
        if self.predictor is not None:
            return [f, self.predictor(f)]
        return f


class IdentityBackbone(nn.Module):
    """Feeds embeddings straight to the head; one dummy parameter list for the EMA."""

    def __init__(self, cfg):
        super().__init__()
        shapes = getattr(cfg, "EMA_SHAPES", [(8,)])
        self.params = nn.ParameterList([nn.Parameter(torch.randn(s) * 0.02) for s in shapes])

    def forward(self, x):
        return x[0] if isinstance(x, (list, tuple)) else x


class Tap:
    """Records the backbone outputs of a model (with grads), like make_golden.Tap."""

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


def register_backbones():
    from advise_video_ssl_b200 import contrastive as C
    C._MODEL_TYPES["stub"] = StubBackbone
    C._MODEL_TYPES["identity"] = IdentityBackbone
    return C


def rel_err(a, ref):
    ref = ref.double().cpu()
    return (a.double().cpu() - ref).abs().max().item() / max(ref.abs().max().item(), 1e-30)


def rel_err_above_floor(a, ref, floor=0.05):
    """ELEMENT-WISE relative error over the entries whose reference magnitude is at least `floor` times the largest
    one (entries near zero have no meaningful relative error).  north_star's "1e-3 relative" in this form; measured
    values per kernel: profiles/r2_elementwise_grad_error.jsonl."""
    a, ref = a.double().cpu(), ref.double().cpu()
    keep = ref.abs() >= floor * ref.abs().max()
    return ((a - ref).abs()[keep] / ref.abs()[keep]).max().item()


def uniform_like_reference(shape, dim):
    stdv = 1.0 / math.sqrt(dim / 3)
    return torch.rand(*shape).mul_(2 * stdv).add_(-stdv)
