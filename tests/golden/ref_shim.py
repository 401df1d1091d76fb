"""Test-infrastructure ONLY: import the *unmodified* reference (`/root/reference`)
in a container that lacks its third-party dependencies and has no GPU.

Used by `tests/golden/make_golden.py` to generate the committed golden vectors
that pin `oracle/`.  Nothing in the product package, the
`-m gpu` tests, `smoke()` or `bench.py` imports this module: `/root/reference`
does not exist on the GPU box.

What it does (no reference file is modified or copied):
  * pre-seeds `sys.modules` with minimal stand-ins for the eleven leaf modules of
    the five absent packages (fvcore, pytorchvideo, megfile, tensorboardX,
    open_clip) that `models/__init__.py` pulls in eagerly;
  * neutralises `Tensor.cuda` / `Module.cuda` so the hard-coded `.cuda()` calls
    (models/losses.py:21-22, models/contrastive.py:64,199,397,...) run on CPU;
  * offers `make_cfg()` with the keys the reference forgets to define
    (`NUM_SHARDS`, `TRAIN.BATCH_SIZE`);
  * offers `register_stub_backbone()` that feeds `[B, D]` embeddings straight to
    the head through the module-level `_MODEL_TYPES` dict (contrastive.py:20-28).
"""
import copy
import os
import sys
import types

import torch
import torch.nn as nn

REFERENCE_ROOT = os.environ.get("AVSSL_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "models", "contrastive.py"))


# --------------------------------------------------------------------------- stubs
class _CfgNode(dict):
    """Attribute-access dict standing in for fvcore.common.config.CfgNode."""

    def __init__(self, init=None, **kw):
        super().__init__()
        for k, v in (init or {}).items():
            self[k] = _CfgNode(v) if isinstance(v, dict) and not isinstance(v, _CfgNode) else v

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v

    def clone(self):
        return copy.deepcopy(self)

    def merge_from_file(self, *a, **k):
        raise NotImplementedError("stub")

    def merge_from_list(self, lst):
        for k, v in zip(lst[0::2], lst[1::2]):
            node = self
            parts = k.split(".")
            for p in parts[:-1]:
                node = node[p]
            node[parts[-1]] = v

    def freeze(self):
        pass

    def defrost(self):
        pass


class _Registry:
    def __init__(self, name):
        self._name = name
        self._map = {}

    def register(self, obj=None):
        if obj is None:
            def deco(o):
                self._map[o.__name__] = o
                return o
            return deco
        self._map[obj.__name__] = obj
        return obj

    def get(self, name):
        return self._map[name]

    def __contains__(self, name):
        return name in self._map


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _cat_all_gather(tensors, local=False):
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return tensors
    ws = dist.get_world_size()
    out = [torch.ones_like(tensors) for _ in range(ws)]
    dist.all_gather(out, tensors, async_op=False)
    return torch.cat(out, dim=0)


def _ws():
    import torch.distributed as dist
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def _rk():
    import torch.distributed as dist
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


_installed = False


def install():
    """Install the stand-ins and put the reference on sys.path. Idempotent."""
    global _installed
    if _installed:
        return
    if not reference_available():
        raise RuntimeError("reference tree not found at %s" % REFERENCE_ROOT)

    # fvcore
    _mod("fvcore")
    _mod("fvcore.common")
    _mod("fvcore.common.registry", Registry=_Registry)
    _mod("fvcore.common.config", CfgNode=_CfgNode)
    _mod("fvcore.nn")

    def c2_msra_fill(module):
        nn.init.kaiming_normal_(module.weight, mode="fan_out", nonlinearity="relu")
        if getattr(module, "bias", None) is not None:
            nn.init.constant_(module.bias, 0)

    def c2_xavier_fill(module):
        nn.init.kaiming_uniform_(module.weight, a=1)
        if getattr(module, "bias", None) is not None:
            nn.init.constant_(module.bias, 0)

    _mod("fvcore.nn.weight_init", c2_msra_fill=c2_msra_fill, c2_xavier_fill=c2_xavier_fill)

    # pytorchvideo
    _mod("pytorchvideo")
    _mod("pytorchvideo.layers")
    _mod(
        "pytorchvideo.layers.distributed",
        cat_all_gather=_cat_all_gather,
        get_world_size=_ws,
        get_local_size=_ws,
        get_local_rank=_rk,
        get_local_process_group=lambda: None,
        init_distributed_training=lambda *a, **k: None,
        _LOCAL_PROCESS_GROUP=None,
    )

    class NaiveSyncBatchNorm1d(nn.BatchNorm1d):
        def __init__(self, num_sync_devices=None, global_sync=False, **kw):
            super().__init__(**kw)

    class NaiveSyncBatchNorm3d(nn.BatchNorm3d):
        def __init__(self, num_sync_devices=None, global_sync=False, **kw):
            super().__init__(**kw)

    _mod(
        "pytorchvideo.layers.batch_norm",
        NaiveSyncBatchNorm1d=NaiveSyncBatchNorm1d,
        NaiveSyncBatchNorm3d=NaiveSyncBatchNorm3d,
    )

    class Swish(nn.Module):
        def forward(self, x):
            return x * torch.sigmoid(x)

    _mod("pytorchvideo.layers.swish", Swish=Swish)
    _mod("pytorchvideo.losses")

    class SoftTargetCrossEntropyLoss(nn.Module):
        def __init__(self, *a, **k):
            super().__init__()

    _mod(
        "pytorchvideo.losses.soft_target_cross_entropy",
        SoftTargetCrossEntropyLoss=SoftTargetCrossEntropyLoss,
    )

    # megfile / tensorboardX / open_clip
    _mod("megfile", smart_open=open, smart_exists=os.path.exists)
    _mod("megfile.s3", s3_listdir=lambda *a, **k: [])

    class SummaryWriter:
        def __init__(self, *a, **k):
            pass

    _mod("tensorboardX", SummaryWriter=SummaryWriter)
    _mod("open_clip")

    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
        nn.Module.cuda = lambda self, *a, **k: self

    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    _installed = True


def load_reference():
    """Returns the reference's `models.contrastive` module (unmodified source)."""
    install()
    import models.contrastive as rc  # noqa: E402  (from /root/reference)
    return rc


# ------------------------------------------------------------------- cfg + backbones
def make_cfg(**over):
    """Reference defaults (configs/defaults.py) + the keys it forgets (Q1, Q2)."""
    install()
    from configs.defaults import get_cfg
    cfg = get_cfg()
    cfg.NUM_SHARDS = 1
    cfg.SHARD_ID = 0
    cfg.TRAIN.BATCH_SIZE = 64
    cfg.NUM_GPUS = 1
    cfg.MODEL.MODEL_NAME = "ContrastiveModel"
    cfg.MODEL.ARCH = "stub"
    cfg.BN.NORM_TYPE = "batchnorm"
    cfg.BN.NUM_SYNC_DEVICES = 1
    cfg.SSL.BN_SYNC_MLP = False
    cfg.CONTRASTIVE.KNN_ON = False
    lst = []
    for k, v in over.items():
        lst += [k.replace("__", "."), v]
    cfg.merge_from_list(lst)
    return cfg


class StubBackbone(nn.Module):
    """forward([x]) -> x @ W^T, or [feat, pred] when `predictor` (byol)."""

    def __init__(self, cfg):
        super().__init__()
        d_in = getattr(cfg, "STUB_IN_DIM", cfg.CONTRASTIVE.DIM)
        self.proj = nn.Linear(d_in, cfg.CONTRASTIVE.DIM, bias=False)
        self.predictor = None
        if len(cfg.CONTRASTIVE.PREDICTOR_DEPTHS) > 0:
            self.predictor = nn.Linear(cfg.CONTRASTIVE.DIM, cfg.CONTRASTIVE.DIM, bias=True)

    def forward(self, x):
        if isinstance(x, (list, tuple)):
            x = x[0]
        f = self.proj(x)
        if self.predictor is not None:
            return [f, self.predictor(f)]
        return f


def register_stub_backbone(rc):
    rc._MODEL_TYPES["stub"] = StubBackbone
