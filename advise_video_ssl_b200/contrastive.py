"""B200-native `ContrastiveModel`: the drop-in for the reference's `models/contrastive.py`
(same class / function names, arguments, return values, buffer names, error behaviour).
All arithmetic on the path runs in the hand-written sm_100a kernels behind
`include/avssl_b200.h`; this file is host orchestration written against the behavioural
contract of the reference (file:line below are in /root/reference), not a translation of it.

Step anatomy in MoCo mode (A2-A6; the step `bench.py` times through `contrastive_forward`):

    launch 1  EMA of the key encoder + `iter += 1`                         (K1,  :158-172, :313-314)
    [hist encoder forward - the host application's backbone]
    launch 2  Normalize(key features) [+ store into every rank's exchange  (K2 + C3 push, :350, :216-230)
              buffer over NVLink, W > 1]
    launch 3  Normalize(f) + logits + InfoNCE fwd/bwd [+ wait for the      (K2 + K3 + C3 wait + K4 + C9,
              exchange, un-shuffle by index] + queue ring write            :462-503, :263-292)

What deliberately differs from the reference:
  * no host synchronisation on the step path: `iter` / `ptr` stay on the device, the shuffle
    permutation travels between hosts, and the reference's asserts on `ptr` / bank indices become
    bits of a device status word (`check_device_status()`);
  * history parameters are updated in place (the reference rebinds `.data`);
  * `logits` is returned detached (the loss carries the gradient) and can be skipped with
    `materialize_logits = False`; the dummy logits of byol / swav / simclr are a cached constant;
  * multi-GPU queue consistency (C9) is explicit: `queue_mode` "reference" (default) makes every
    rank enqueue rank 0's keys - what the reference's DDP buffer broadcast (models/build.py:76-83)
    leaves in every queue - without broadcasting 33.5 MB per step; "canonical" enqueues all W*B keys;
    "local" is the reference's raw per-rank behaviour (correct only under that DDP broadcast);
  * the SwAV prototype count is `cfg.CONTRASTIVE.SWAV_NUM_PROTOTYPES` when present (reference: 1000).
Backbones are not part of this package: `_MODEL_TYPES` is filled from the host application's
`models.video_model_builder` when importable (inside the reference tree), or by the caller.
"""
import contextlib
import logging
import math
import os

import numpy as np
import torch
import torch.nn as nn

from . import _lib, ops
from . import distributed as du
from . import losses
from .autograd import (BankDot, ByolSimilarity, MocoInfoNce, NtXentRows, SwavSwappedCe, l2norm_lastdim)
from .banks import Memory, Memory1D, Normalize, _bank_init
from .shuffle import ShufflePlan, broadcast_from_rank0

logger = logging.getLogger(__name__)

# models/contrastive.py:20-28 - architecture name -> backbone class
_MODEL_TYPES = {}
try:
    from models.video_model_builder import X3D, MViT, ResNet, SlowFast  # type: ignore

    for _name, _cls in (("slowfast", SlowFast), ("slow", ResNet), ("c2d", ResNet), ("i3d", ResNet),
                        ("slow_c2d", ResNet), ("x3d", X3D), ("mvit", MViT)):
        _MODEL_TYPES[_name] = _cls
except Exception:  # pragma: no cover - outside the reference tree the caller registers backbones
    pass

try:
    from models.build import MODEL_REGISTRY  # type: ignore
except Exception:  # pragma: no cover
    class _Registry(dict):
        """Just enough of fvcore's Registry for `MODEL_REGISTRY.get("ContrastiveModel")(cfg)`."""

        def register(self, obj=None):
            if obj is None:
                return self.register
            self[obj.__name__] = obj
            return obj

        def get(self, name):
            return self[name]

    MODEL_REGISTRY = _Registry()

QUEUE_MODES = ("reference", "canonical", "local")


def _opt(node, name, default):
    """cfg keys this package adds are optional: the host's yaml need not know them."""
    try:
        return getattr(node, name)
    except (AttributeError, KeyError):
        return default


def _first_if_list(x):
    return (x[0], list(x[1:])) if isinstance(x, (list, tuple)) else (x, [])


class _KeyBatch:
    """The key features of one encoder pass, possibly not yet in this rank's original row order.

    `local` (when set) holds this rank's normalised keys in the original row order.  Otherwise the rows
    are still where the encoder / the exchange left them and consumers index them instead of
    materialising an un-shuffled copy:
      xchg   every rank's normalised rows in the NVLink exchange buffers (encoder order, rank-major) -- when
             `table` is set as well, this rank's raw rows that the head launch still has to normalise and push;
      table  otherwise a [world * rows, D] tensor on this device in the same order (`raw`: un-normalised);
      restore ([W, B] device int64, None without shuffle) maps original rows to that order.
    """

    __slots__ = ("local", "xchg", "table", "raw", "restore", "rank", "rows", "world")

    def __init__(self, local=None, xchg=None, table=None, raw=False, restore=None, rank=0, rows=0, world=1):
        self.local, self.xchg, self.table, self.raw = local, xchg, table, raw
        self.restore, self.rank, self.rows, self.world = restore, rank, rows, world


class ContrastiveModel(nn.Module):
    """Contrastive head in its mem / moco / byol / swav / simclr modes (models/contrastive.py:31-916)."""

    def __init__(self, cfg):
        super(ContrastiveModel, self).__init__()
        ct = cfg.CONTRASTIVE
        self.cfg = cfg
        self.backbone = _MODEL_TYPES[cfg.MODEL.ARCH](cfg)
        self.type, self.T, self.dim = ct.TYPE, ct.T, ct.DIM
        self.length, self.k = ct.LENGTH, ct.QUEUE_LEN
        self.mmt, self.momentum_annealing = ct.MOMENTUM, ct.MOMENTUM_ANNEALING
        self.duration = 1
        self.num_gpus = cfg.NUM_GPUS
        self.l2_norm = Normalize()
        self.knn_num_imgs = 0
        self.knn_on = ct.KNN_ON
        self.train_labels = np.zeros((0,), dtype=np.int32)
        self.num_pos = 2
        self.num_crops = cfg.DATA.TRAIN_CROP_NUM_TEMPORAL * cfg.DATA.TRAIN_CROP_NUM_SPATIAL
        self.nce_loss_fun = losses.get_loss_func("contrastive_loss")(reduction="mean")
        self.softmax = nn.Softmax(dim=1)

        # ---- knobs of the B200 path (absent from the reference)
        self.materialize_logits = True
        self.infonce_impl = _lib.IMPL_AUTO
        # MoCo head in two launches (ops.moco_infonce_sweep): the sweep of q against the queue starts right after the
        # query encoder, on its own stream, and runs beside the momentum update / shuffle / key encoder that the
        # reference issues next (:462 then :478).  sweep_ctas: None = sized against the momentum update, 0 = every SM.
        self.overlap_sweep = True
        self.sweep_ctas = int(os.environ["AVSSL_SWEEP_CTAS"]) if os.environ.get("AVSSL_SWEEP_CTAS") else None
        self.ntxent_impl = _lib.IMPL_AUTO  # tcgen05 tf32 when D allows it; IMPL_SIMT = exact fp32
        self.queue_mode = str(_opt(ct, "QUEUE_MODE", "reference"))
        assert self.queue_mode in QUEUE_MODES, "CONTRASTIVE.QUEUE_MODE must be one of %s" % (QUEUE_MODES,)
        # C3 over NVLink peer memory: None = decide at the first multi-GPU step (on when all ranks of
        # the shuffle group share one box and CUDA IPC works), True / False = forced
        self._peer_exchange_on = _opt(ct, "PEER_EXCHANGE", None)
        self._peer_xchgs = {}
        self._ema = None          # (online params, history params, ops.EmaPlan)
        self._iter_seen = None    # host mirror of `iter`: [(data_ptr, version), value]
        self._const = {}          # small cached device constants (dummy logits, row index ranges)
        self._status = torch.zeros(1, dtype=torch.int32)  # device status word; plain attribute, moved by _apply

        build = getattr(self, "_build_" + self.type, None)
        if build is not None:
            build(cfg)
        self.simclr_dist_on = ct.SIMCLR_DIST_ON
        if self.knn_on:
            self.knn_mem = Memory(self.length, 1, self.dim, cfg)

    # ---- per-mode state (buffer names / shapes / dtypes / draw order = the checkpoint contract, :72-129)
    def _build_mem(self, cfg):
        self.mem_type = cfg.CONTRASTIVE.MEM_TYPE
        bank = Memory1D if self.mem_type == "1d" else Memory
        self.memory = bank(self.length, self.duration, self.dim, cfg)
        self.examplar_type = "video"
        self.interp = cfg.CONTRASTIVE.INTERP_MEMORY

    def _build_momentum_pair(self, cfg):
        self.backbone_hist = _MODEL_TYPES[cfg.MODEL.ARCH](cfg)
        for p in self.backbone_hist.parameters():
            p.requires_grad = False
        self.register_buffer("ptr", torch.tensor([0]))
        self.ptr.requires_grad = False
        self.register_buffer("queue_x", _bank_init(self.k, self.dim))
        self.register_buffer("iter", torch.zeros([1], dtype=torch.long))
        # shuffle BN is pointless when sync-BN already spans every GPU, and BYOL never shuffles (:91-99)
        bn = cfg.BN
        spans_all = "sync" in bn.NORM_TYPE and bn.NUM_SYNC_DEVICES == cfg.NUM_GPUS
        self._batch_shuffle_on = not (spans_all or self.type == "byol")

    _build_moco = _build_momentum_pair
    _build_byol = _build_momentum_pair

    def _build_swav(self, cfg):
        self.swav_use_public_code = True
        n_proto = int(_opt(cfg.CONTRASTIVE, "SWAV_NUM_PROTOTYPES", 1000))
        self.swav_prototypes = nn.Linear(self.dim, n_proto, bias=False)
        self.swav_eps_sinkhorn = 0.05
        self.swav_use_the_queue = False
        rows = cfg.CONTRASTIVE.SWAV_QEUE_LEN
        if rows > 0:
            self.register_buffer("queue_swav", torch.zeros(2, rows // du.get_world_size(), self.dim))

    def _build_simclr(self, cfg):
        self._simclr_precompute_pos_neg_mask_multi()

    def _simclr_precompute_pos_neg_mask_multi(self):
        """The reference builds float64 masks here (:806-846) that only its dead `distributed_loss`
        branch (:748-768) reads; the attributes exist, empty."""
        self.pos_mask, self.neg_mask = [], None

    # ------------------------------------------------------------------ housekeeping
    def _apply(self, fn, *args, **kwargs):
        # .cuda() / .to() / .float() move storages: every cached device pointer or constant is stale
        self._ema = None
        self._iter_seen = None
        self._const = {}
        out = super(ContrastiveModel, self)._apply(fn, *args, **kwargs)
        self._status = fn(self._status)
        return out

    def _load_from_state_dict(self, *args, **kwargs):
        self._iter_seen = None  # `iter` may change under us
        return super(ContrastiveModel, self)._load_from_state_dict(*args, **kwargs)

    def check_device_status(self):
        """Raise what the reference's host asserts on `ptr` / bank indices would have raised, from the
        device status word (one synchronisation: call it at logging / epoch boundaries)."""
        flags = int(self._status.item())
        assert not (flags & _lib.DEVFLAG_QUEUE_OVERRUN), "queue overrun: ptr + n > K (models/contrastive.py:285)"
        if flags & _lib.DEVFLAG_BAD_INDEX:
            raise IndexError("memory-bank / exchange index out of range")
        if flags & _lib.DEVFLAG_PEER_TIMEOUT:
            raise RuntimeError("a peer rank's keys did not arrive within the exchange timeout (dead or stalled rank)")
        return flags

    def _cached(self, key, make):
        t = self._const.get(key)
        if t is None:
            t = self._const[key] = make()
        return t

    def _cached_dummy_logits(self, n, device):
        """K8: zeros [n, K+1] with column 0 = 9999 (:585-592), built once on the device."""
        def make():
            t = torch.zeros(n, self.k + 1, dtype=torch.float, device=device)
            t[:, 0] = 9999.0
            return t
        return self._cached(("dummy", n, device), make)

    def _row_range(self, lo, hi, device):
        return self._cached(("rows", lo, hi, device), lambda: torch.arange(lo, hi, dtype=torch.int64, device=device))

    # ---------------------------------------------------------------------- kNN bank
    @torch.no_grad()
    def knn_mem_update(self, q_knn, index):
        if self.knn_on:
            self.knn_mem.update(q_knn, momentum=1.0, ind=index, time=torch.zeros_like(index), interp=False,
                                status=self._status)

    @torch.no_grad()
    def init_knn_labels(self, train_loader):
        """Labels of the whole training set next to the kNN bank (:142-156)."""
        logger.info("initializing knn labels")
        labels = np.asarray(train_loader.dataset._labels, dtype=np.int32)
        self.num_imgs = int(labels.shape[0])
        self.train_labels = torch.from_numpy(labels).long().to(self.knn_mem.memory.device)
        if self.length != self.num_imgs:
            logger.error("Kinetics dataloader size: {} differs from memorybank length {}".format(
                self.num_imgs, self.length))
            self.knn_mem.resize(self.num_imgs, 1, self.dim)

    @torch.no_grad()
    def eval_knn(self, q_knn, knn_k=200):
        """Similarity of the queries to every bank row and its top-k (:232-241); eval only."""
        bank = self.knn_mem.memory
        q = q_knn.reshape(q_knn.size(0), -1)
        bank = bank.reshape(bank.size(0), -1)
        if q.is_cuda and q.dtype == torch.float32 and ops.topk_rows_supported(q.size(0), bank.size(0), knn_k):
            # similarities from the head's tcgen05 mainloop, exact top-k behind it (csrc/knn.cu)
            return ops.knn_similarity_topk(q.detach().contiguous(), bank.contiguous(), knn_k)
        return (q @ bank.t()).topk(knn_k, dim=1, largest=True, sorted=True)  # k > 1024 or a bank beyond the plan

    # ------------------------------------------------------------------- K1: momentum update
    def _ema_state(self):
        """Parameter pairs in `backbone_hist.named_parameters()` order (:164-172) and the device
        pointer table over them.  The table caches raw addresses, so it is re-validated on every
        call: anything that re-seats a parameter's storage without `_apply` (`p.data = ...` as the
        reference itself does, `load_state_dict(assign=True)`, flattening) rebuilds it."""
        st = self._ema
        if st is not None:
            on, hi, plan = st
            if plan.matches([p.data for p in on], [p.data for p in hi]):
                return st
        online = dict(self.backbone.named_parameters())
        pairs = [(online[name], p) for name, p in self.backbone_hist.named_parameters()]
        on, hi = [a for a, _ in pairs], [b for _, b in pairs]
        self._ema = st = (on, hi, ops.EmaPlan([p.data for p in on], [p.data for p in hi]))
        return st

    def _ema_lists(self):
        on, hi, _ = self._ema_state()
        return [p.data for p in on], [p.data for p in hi]

    @torch.no_grad()
    def _update_history(self, _bump_iter=False):
        """hist <- online*(1-m) + hist*m over every parameter in ONE launch (:158-172).  Whether this is
        step 0 (copy first, then blend) comes from a host mirror of `iter`: the kernels bump the
        buffer through its raw pointer (tensor version untouched), any torch-side write bumps the
        version and costs one re-read.  Steady state: no device->host traffic."""
        _, _, plan = self._ema_state()
        tag = (self.iter.data_ptr(), self.iter._version)
        if self._iter_seen is None or self._iter_seen[0] != tag:
            self._iter_seen = [tag, int(self.iter.item())]
        plan.run(self.mmt, self.iter, bump_iter=_bump_iter, first_iter=self._iter_seen[1] == 0)
        if _bump_iter:
            self._iter_seen[1] += 1

    @torch.no_grad()
    def momentum_anneal_cosine(self, epoch_exact):
        """m = 1 - (1 - m0) (cos(pi e / E) + 1) / 2 in host fp64 (:251-261)."""
        base = self.cfg.CONTRASTIVE.MOMENTUM
        phase = math.cos(math.pi * epoch_exact / self.cfg.SOLVER.MAX_EPOCH) + 1.0
        self.mmt = 1 - (1 - base) * phase * 0.5

    # --------------------------------------------------------- shuffle BN (A6; C1, C2, C3)
    def _shuffle_scope(self):
        """(group, size, rank) the shuffle runs over: the local process group under LOCAL_SHUFFLE_BN
        (:185-197), else WORLD."""
        if self.num_gpus <= 1:
            return None, 1, 0
        if self.cfg.CONTRASTIVE.LOCAL_SHUFFLE_BN:
            return du._LOCAL_PROCESS_GROUP, du.get_local_size(), du.get_local_rank()
        return None, self.cfg.NUM_GPUS * self.cfg.NUM_SHARDS, torch.distributed.get_rank()

    @torch.no_grad()
    def _batch_shuffle(self, x):
        """x: [clip] or [clip, crop].  Returns ([shuffled...], idx_restore) like the reference; the rows
        this rank ends up with are those of `cat_all_gather(x)[perm.view(W, -1)[rank]]`."""
        assert len(x) in (1, 2)
        group, size, rank = self._shuffle_scope()
        bsz = x[0].shape[0]
        # every rank consumes its CPU generator exactly like the reference (:198); rank 0's draw wins (C2)
        perm = torch.randperm(bsz * size)
        if self.num_gpus > 1:
            perm = broadcast_from_rank0(perm, x[0].device)
        # transport of the rows (C1): NVLink peer stores straight into the final positions when the ranks share
        # a box, else an NCCL all-to-all; either way only the rows a rank keeps cross the fabric
        scatters = [None] * len(x)
        if size > 1 and x[0].is_cuda and self._peer_transport_on():
            scatters = [self._peer_scatter(t) for t in x]
        plan = ShufflePlan(perm.numpy(), size, rank, bsz, x[0].device, need_alltoall=any(s is None for s in scatters))
        if x[0].is_cuda and torch.cuda.is_current_stream_capturing():
            # the captured upload node reads the plan's pinned host buffer on every replay
            self._const.setdefault("captured_plans", []).append(plan)
        return [plan.shuffled(t, group, sc, self._status) for t, sc in zip(x, scatters)], plan.restore

    def _peer_transport_on(self):
        """Do the ranks of the shuffle group exchange rows over NVLink peer memory (CUDA IPC) rather than NCCL?
        Decided once, collectively: on when they share one box and a probe exchange can be set up."""
        if self._peer_exchange_on is None:
            group, size, _ = self._shuffle_scope()
            one_box = int(_opt(self.cfg, "NUM_SHARDS", 1)) == 1 or (self.cfg.CONTRASTIVE.LOCAL_SHUFFLE_BN and group is not None)
            on = bool(one_box) and size <= _lib.MAX_PEERS
            if on:
                try:  # PeerExchange raises on every rank together or on none
                    ops.PeerExchange(1, 4, group=group).close()
                except _lib.AvsslError as e:
                    logger.warning("NVLink peer exchange unavailable, using NCCL: %s", e)
                    on = False
            self._peer_exchange_on = on
        return bool(self._peer_exchange_on)

    def _use_peer_exchange(self, x):
        """Can the key rows `x` ([rows, D] fp32) go through the NVLink exchange buffers?"""
        if not (x.is_cuda and x.dim() == 2 and x.dtype == torch.float32 and x.shape[1] % 4 == 0):
            return False
        return self._peer_transport_on()

    def enable_peer_exchange(self, on=True):
        """Force the cross-GPU exchanges (C1 rows, C3 keys) onto NVLink peer memory (`ops.PeerExchange` /
        `ops.PeerScatter`) or back onto NCCL.  Collective: call it on every rank."""
        self._peer_exchange_on = bool(on)
        if not on:
            for ex in self._peer_xchgs.values():
                for one in (ex if isinstance(ex, tuple) else (ex,)):
                    one.close()
            self._peer_xchgs = {}
        return self

    def _peer_xchg(self, rows, dim):
        key = (rows, dim)
        ex = self._peer_xchgs.get(key)
        if ex is None:
            group, _, _ = self._shuffle_scope()
            ex = self._peer_xchgs[key] = ops.PeerExchange(rows, dim, group=group)
        return ex

    def _peer_scatter(self, t):
        """The scatter buffers for tensors shaped like `t` (None when its rows are not 16-byte multiples)."""
        row_bytes = (t.numel() // max(t.shape[0], 1)) * t.element_size()
        if t.shape[0] == 0 or row_bytes % 16 != 0:
            return None
        key = ("scatter", t.shape[0], row_bytes)
        sc = self._peer_xchgs.get(key)
        if sc is None:
            group, _, _ = self._shuffle_scope()
            sc = self._peer_xchgs[key] = ops.PeerScatter(t.shape[0], row_bytes, group=group)
        return sc

    @torch.no_grad()
    def _batch_unshuffle(self, x, idx_restore):
        """Rows of this rank's ORIGINAL batch out of the per-rank results `x` computed on the shuffled
        batch: cat_all_gather(x)[idx_restore[rank]] (:216-230)."""
        group, size, rank = self._shuffle_scope()
        mine = idx_restore[rank, :]
        if size == 1:
            return x[mine]
        if self._use_peer_exchange(x):
            ex = self._peer_xchg(x.shape[0], x.shape[1])
            ex.push(x.contiguous())
            return ex.wait_gather(mine.contiguous(), status=self._status)
        everyone = du.cat_all_gather(x, local=bool(self.cfg.CONTRASTIVE.LOCAL_SHUFFLE_BN))
        return everyone[mine]

    # ---------------------------------------------------------------------- BYOL
    def sim_loss(self, q, k):
        """-mean(sum_c q k) / T for already-normalised q (:243-249)."""
        return ByolSimilarity.apply(q, k, self.T, False)

    # ------------------------------------------------------------------- K4 (+ C9): queue
    def _queue_rows_across_ranks(self, key):
        """C9.  The reference enqueues each rank's own keys and relies on DDP re-broadcasting rank 0's
        buffers before every forward (models/build.py:76-83), so what survives in every queue is rank
        0's keys.  "reference": take rank 0's rows directly (32 KB broadcast instead of 33.5 MB);
        "canonical": all ranks' rows in rank order; "local": this rank's rows."""
        if self.num_gpus <= 1 or self.queue_mode == "local":
            return key
        if self.queue_mode == "canonical":
            return du.cat_all_gather(key.contiguous())
        shared = key.contiguous().clone()
        torch.distributed.broadcast(shared, src=0)
        return shared

    @torch.no_grad()
    def _dequeue_and_enqueue(self, keys, extra_keys=None):
        """queue_x[ptr:ptr+n] = rows; ptr = (ptr + n) wrapping exactly at K (:263-292), `ptr` never
        leaving the device.  keys[0] only, unless MOCO_MULTI_VIEW_QUEUE adds the other views and
        `extra_keys`."""
        batch = [keys[0]]
        if self.cfg.CONTRASTIVE.MOCO_MULTI_VIEW_QUEUE:
            assert len(keys) > 0, "need to have multiple views for adding them to queue"
            batch = list(keys)
            for group in (extra_keys or []):
                batch.extend(group)
        for key in batch:
            rows = self._queue_rows_across_ranks(key.detach())
            assert self.k % int(rows.size(0)) == 0  # :284; `ptr + n <= K` (:285) is checked on the device
            ops.queue_enqueue(self.queue_x, self.ptr, rows.contiguous(), self._status)

    # ---------------------------------------------------------------- key features (A5)
    @torch.no_grad()
    def batch_clips(self, clips):
        """[[pathway tensors] per clip] -> [per pathway: all clips stacked along the batch] (:294-306)."""
        if len(clips) == 1:
            return list(clips[0])
        return [torch.cat([clip[j] for clip in clips], dim=0) for j in range(len(clips[0]))]

    def _normalised_keys(self, feats, restore, defer):
        """Normalize + un-shuffle of one [rows, D] output of the key encoder.  With `defer` the rows stay
        where they are and the head launch picks them up by index: on one GPU even un-normalised (Normalize
        happens inside the head launch, no launch of its own); across GPUs Normalize and the store into
        every rank's NVLink exchange buffer are ONE launch.  Without `defer` this rank's rows come back."""
        shuffled = restore is not None
        _, size, rank = self._shuffle_scope()
        n = feats.shape[0]
        if size == 1:
            if defer:
                return _KeyBatch(table=feats.contiguous(), raw=True, restore=restore, rows=n)
            keys = self.l2_norm(feats)
            return _KeyBatch(local=keys[restore[0, :]] if shuffled else keys)
        # rows must cross ranks for the un-shuffle, and for the queue rows of C9 when the head takes them by index
        crosses = shuffled or (defer and self.queue_mode != "local")
        if crosses and self._use_peer_exchange(feats):
            ex = self._peer_xchg(n, feats.shape[1])
            if defer:  # the head launch normalises and pushes the rows itself (one extra CTA), then waits for everybody's
                return _KeyBatch(xchg=ex, table=feats.contiguous(), raw=True, restore=restore, rank=rank, rows=n, world=size)
            ex.push_normalized(feats.contiguous(), 0.0)
            pending = _KeyBatch(xchg=ex, restore=restore, rank=rank, rows=n, world=size)
            return _KeyBatch(local=ex.wait_gather(self._local_rows(pending, feats.device), status=self._status))
        keys = self.l2_norm(feats)
        if not crosses:
            return _KeyBatch(local=keys)
        everyone = du.cat_all_gather(keys.contiguous(), local=bool(self.cfg.CONTRASTIVE.LOCAL_SHUFFLE_BN))
        if defer:
            return _KeyBatch(table=everyone, restore=restore, rank=rank, rows=n, world=size)
        return _KeyBatch(local=everyone[restore[rank, :]].detach())

    def _local_rows(self, kb, device):
        """Where this rank's original rows sit in the key rows (None = the kernel's default: its own block)."""
        if kb.restore is not None:
            return kb.restore[kb.rank, :].contiguous()
        if kb.xchg is not None or kb.world == 1:
            return None
        return self._row_range(kb.rank * kb.rows, (kb.rank + 1) * kb.rows, device)

    def _queue_rows(self, kb, device):
        """The key rows that go into the queue under `queue_mode` (C9), None = this rank's own block in order."""
        if self.queue_mode == "local" or kb.world == 1:
            return self._local_rows(kb, device)
        if self.queue_mode == "reference":  # rank 0's original rows, on every rank
            return kb.restore[0, :].contiguous() if kb.restore is not None else self._row_range(0, kb.rows, device)
        if kb.restore is not None:          # canonical: every rank's original rows, rank-major
            return kb.restore.reshape(-1).contiguous()
        return self._row_range(0, kb.world * kb.rows, device)

    def _shuffle_ahead(self, groups):
        """Shuffle every key clip NOW, on a side stream, so that the index uploads, the row gathers and
        (W > 1) the all-to-all of the clip rows (C1) run underneath the momentum update that the caller
        launches next on the main stream (north_star: the exchange overlaps the EMA kernel)."""
        dev_t = groups[0][0]
        if not dev_t.is_cuda:
            return [self._batch_shuffle(clip) for clip in groups], None
        main = torch.cuda.current_stream(dev_t.device)
        side = self._cached(("side_stream", dev_t.device), lambda: torch.cuda.Stream(device=dev_t.device, priority=-1))
        side.wait_stream(main)
        with torch.cuda.stream(side):
            done = [self._batch_shuffle(clip) for clip in groups]
        for views, restore in done:
            for t in list(views) + [restore]:
                t.record_stream(main)
        return done, side

    def _encode_keys(self, clip, restore, want_pred, defer=False):
        """One pass of the key encoder over one (possibly batched, already shuffled) clip: forward,
        Normalize, un-shuffle (:335-357).  Returns (_KeyBatch, [predictor keys])."""
        feats, heads = _first_if_list(self.backbone_hist(clip))
        pred = []
        if want_pred:
            pred = [self._normalised_keys(h, restore, False).local for h in heads]
        return self._normalised_keys(feats, restore, defer), pred

    @torch.no_grad()
    def compute_key_feat(self, clips_k, compute_predictor_keys=False, batched_inference=True):
        """Momentum-update the key encoder, count the iteration, and run it over the key clips (:308-371).
        Returns the list of [B, D] keys (and the list of predictor-key lists when asked)."""
        out = self._key_batches(clips_k, compute_predictor_keys, batched_inference, defer=False)
        keys = [kb.local for kb in out[0]]
        return (keys, out[1]) if compute_predictor_keys else keys

    @torch.no_grad()
    def _key_batches(self, clips_k, want_pred, batched_inference, defer):
        assert self.training
        n_clips = len(clips_k)
        assert n_clips > 0
        bsz = clips_k[0][0].shape[0]
        # the reference's size test (:318-319; numel() already includes the batch)
        if n_clips * bsz * clips_k[0][0].numel() > 4 * 64 * 3 * 8 * 224 * 224:
            batched_inference = False
        same_shape = all(view.shape[1:] == ref.shape[1:] for clip in clips_k for view, ref in zip(clip, clips_k[0]))
        batched = bool(batched_inference and same_shape)
        # all key clips through the encoder as one batch, split back per clip afterwards (:321-331, :359-369)
        groups = [self.batch_clips(clips_k)] if batched else list(clips_k)
        restores, side = [None] * len(groups), None
        if self._batch_shuffle_on:
            done, side = self._shuffle_ahead(groups)
            groups, restores = [g for g, _ in done], [r for _, r in done]
        self._update_history(_bump_iter=True)  # :313-314, one launch; the shuffles above run beside it
        if side is not None:
            torch.cuda.current_stream(side.device).wait_stream(side)
        defer = defer and n_clips == 1
        encoded = [self._encode_keys(g, r, want_pred, defer) for g, r in zip(groups, restores)]
        if not batched:
            return [kb for kb, _ in encoded], [pk for _, pk in encoded if pk]
        kb, pred = encoded[0]
        if n_clips == 1:
            return [kb], [pred]
        keys = [_KeyBatch(local=kb.local[i * bsz:(i + 1) * bsz]) for i in range(n_clips)]
        # (the reference slices the LIST of predictor keys here, :367 - an empty or one-element list per clip)
        return keys, [pred[i * bsz:(i + 1) * bsz] for i in range(n_clips)]

    # ------------------------------------------------------------------------ forward
    def forward(self, clips, index=None, time=None, epoch_exact=None, keys=None):
        if epoch_exact is not None and self.momentum_annealing:
            self.momentum_anneal_cosine(epoch_exact)
        if self.type == "mem":
            return self._forward_mem(clips, index, time)
        if self.type == "moco":
            return self._forward_moco(clips, index, time, keys)
        if self.type == "byol":
            return self._forward_byol(clips, index, keys)
        if self.type == "swav":
            return self._forward_swav(clips, index, epoch_exact)
        if self.type == "simclr":
            return self._forward_simclr(clips, index)
        raise NotImplementedError()

    # ---- mem (:379-442): loss computed and dropped, returns (prod, 0.0, True) like the reference
    def _forward_mem(self, clips, index, time):
        n = clips[0].size(0)
        q = self.backbone(clips)
        if index is None:
            return q
        q = self.l2_norm(q)
        if not self.training:
            assert self.knn_mem.duration == 1
            return self.eval_knn(q)
        time *= self.duration - 1
        # negatives from the CPU generator, as the reference draws them (:390-397): a seeded run sees the
        # same indices.  Column 0 is the positive.
        cols = self.k + 1
        clip_ind = torch.randint(0, self.length, size=(n, cols)).to(q.device)
        clip_ind[:, 0] = index.data
        if self.mem_type != "2d":
            time_ind = torch.zeros(size=(n, cols), dtype=int).to(q.device)
        elif self.interp:
            time_ind = torch.empty(n, cols).uniform_(0, self.duration - 1).to(q.device)
        else:
            time_ind = torch.randint(0, self.duration - 1, size=(n, cols)).to(q.device)
        if self.examplar_type == "clip":
            time_ind[:, 0] = time.data
        elif self.examplar_type != "video":
            raise NotImplementedError("unsupported examplar_type {}".format(self.examplar_type))
        interp = bool(self.interp) and self.mem_type == "2d"
        prod = BankDot.apply(q, self.memory.memory, clip_ind, time_ind, self.T, interp, self._status)
        self.nce_loss_fun(prod)  # evaluated and discarded (:436, :442)
        self.memory.update(q, momentum=self.mmt, ind=index, time=time, interp=self.interp, status=self._status)
        self.knn_mem_update(q, index)
        return prod, 0.0, True

    # ---- moco (:443-506)
    def _forward_moco(self, clips, index, time, keys):
        clips_k = None
        if isinstance(clips[0], list):
            clip_q, clips_k = clips[0], list(clips[1:])
            time[:, 0, :], time[:, 1:, :]  # the reference slices `time` here: None raises TypeError as it does there
        else:
            clip_q = clips
        feat_q, _extra = _first_if_list(self.backbone(clip_q))
        if index is None:
            return feat_q
        if not self.training:
            return self.eval_knn(self.l2_norm(feat_q))

        own_keys = keys is None
        B, D = feat_q.shape
        n_k = 1 if own_keys and len(clips_k) == 1 else (len(clips_k) if own_keys else len(keys))
        head_kw = dict(want_logits=self.materialize_logits, impl=self.infonce_impl,
                       workspace=self._head_workspace(B, D, n_k, feat_q.device))
        sweep_stream = self._sweep_ahead(feat_q, n_k, own_keys, head_kw)
        # the fused ring write rides in the head launch when exactly keys[0] is enqueued (the default) and
        # the kernel's vector path applies; anything else goes through _dequeue_and_enqueue afterwards
        can_fuse = (own_keys and not self.cfg.CONTRASTIVE.MOCO_MULTI_VIEW_QUEUE and D % 4 == 0)
        deferred = None
        if own_keys:
            defer = can_fuse and len(clips_k) == 1 and self._tc_head_ok(B, D)
            batches, _ = self._key_batches(clips_k, False, True, defer)
            if batches[0].local is None:
                deferred = batches[0]
            else:
                keys = [kb.local for kb in batches]
        if deferred is not None:
            # the key rows are still in encoder order (exchange buffers, gathered table, or the raw encoder output):
            # the head launch [waits for them,] un-shuffles by index and writes the queue rows `queue_mode` asks for
            n_enq = deferred.rows * (deferred.world if self.queue_mode == "canonical" else 1)
            assert self.k % n_enq == 0  # :284
            head_kw.update(peer_row_idx=self._local_rows(deferred, feat_q.device),
                           enq_row_idx=self._queue_rows(deferred, feat_q.device), enqueue=(self.ptr, self._status))
            if deferred.xchg is not None:
                head_kw.update(peer=deferred.xchg, push_rows=deferred.table)
            else:
                head_kw.update(key_rows=deferred.table, keys_raw=deferred.raw)
            plan = {"keys": None, "kw": head_kw}
            fused = True
        else:
            fused = (can_fuse and tuple(keys[0].shape) == (B, D) and self.k % B == 0
                     and (self.num_gpus <= 1 or self.queue_mode == "local"))
            if fused:
                head_kw["enqueue"] = (self.ptr, self._status)
            plan = {"keys": [k.detach().contiguous() for k in keys], "kw": head_kw}
        if sweep_stream is not None:
            torch.cuda.current_stream(feat_q.device).wait_stream(sweep_stream)
        loss, logits, q = MocoInfoNce.apply(feat_q, self.queue_x, self.T, plan)
        if own_keys and not fused:
            self._dequeue_and_enqueue(keys)
        self.knn_mem_update(q, index)
        return (logits if self.materialize_logits else None), loss

    def _sweep_ahead(self, feat_q, n_keys, own_keys, head_kw):
        """Launch the sweep of q against the queue NOW, on its own stream: it needs neither the keys nor the momentum
        encoder, and the reference's order puts the whole key path (:478 -> :314 momentum update, :338 shuffle, key
        encoder) between the query encoder (:462) and the logits (:490).  The head launch that follows the key path
        then only merges the sweep's partials with the key term.  Returns the stream to wait for, or None."""
        B, D = feat_q.shape
        if not (self.overlap_sweep and feat_q.is_cuda and self._tc_head_ok(B, D)):
            return None
        dev = feat_q.device
        ctas = self.sweep_ctas
        if ctas is None:  # sized against the momentum update that `compute_key_feat` launches next; None: not worth it
            ctas = self._sweep_ctas_beside_ema(B, dev) if own_keys else None
            if ctas is None:
                return None
        logits = torch.empty(n_keys * B, self.k + 1, dtype=torch.float32, device=dev) if self.materialize_logits else None
        fq = feat_q.detach().contiguous()
        main = torch.cuda.current_stream(dev)
        side = self._cached(("sweep_stream", dev), lambda: torch.cuda.Stream(device=dev, priority=-1))
        side.wait_stream(main)
        with torch.cuda.stream(side):
            ops.moco_infonce_sweep(fq, self.queue_x, self.T, head_kw["workspace"], n_keys=n_keys, logits=logits,
                                   impl=self.infonce_impl, sweep_ctas=ctas)
        # (fq, logits and the workspace stay referenced until the head launch, which the main stream orders behind the sweep)
        head_kw.update(swept=ctas, out={"logits": logits} if logits is not None else None)
        return side

    def _sweep_ctas_beside_ema(self, B, device):
        """CTAs for a sweep that shares the GPU with the momentum update, or None when the single-launch head is the
        better choice.  The tcgen05 sweep owns every register of the SMs it runs on; the momentum update is HBM-bound
        and a flat grid of small CTAs that fills whatever is left.  Give the sweep just enough SMs to end well before
        the update does (measured on B200, DESIGN.md 4 K3: ~2 us per 64-row tile beside the update, ~20 us before its
        first CTA is placed); when the update is short compared with the sweep there is nothing to hide it under.
        Across GPUs the single launch stays: there the sweep, placed BEHIND the key encoder, is what hides the key
        exchange (push over NVLink, wait for every rank's rows); with the sweep moved ahead that round trip and the
        ranks' skew sit bare on the critical path (measured at N = 2: 109.4 us per step against 100.9 us)."""
        if self.num_gpus > 1 and torch.distributed.is_available() and torch.distributed.is_initialized() \
                and torch.distributed.get_world_size() > 1:
            return None

        def size():
            sms = ops.sm_count()
            n_params = sum(p.numel() for p in self.backbone_hist.parameters()) if hasattr(self, "backbone_hist") else 0
            ema_us = 12.0 * n_params / 6.5e6                      # 12 bytes per parameter at ~6.5 TB/s
            row_blocks = (B + 127) // 128
            tiles = (self.k + 63) // 64 * row_blocks
            alone_us = 4.0 + 1.2 * -(-tiles // sms)                # the sweep with every SM to itself
            if ema_us < 2.0 * alone_us + 20.0:
                return -1
            tiles_per_cta = max(1, int((0.8 * ema_us - 20.0) / 2.0))
            return int(min(sms, max(row_blocks, -(-tiles // tiles_per_cta))))
        ctas = self._cached(("sweep_ctas", B, device), size)
        return None if ctas < 0 else ctas

    def _head_workspace(self, B, D, n_keys, device):
        """Scratch of the head launch (split partials + barrier words), owned by the module and zero-filled once:
        the step is stream-ordered, and a CUDA-graph capture must not allocate or clear it on every replay."""
        nbytes = ops.moco_infonce_workspace_bytes(B, D, self.k, n_keys)
        return self._cached(("head_ws", B, D, n_keys, device),
                            lambda: torch.zeros(nbytes, dtype=torch.uint8, device=device))

    def _tc_head_ok(self, B, D):
        return self.infonce_impl != _lib.IMPL_SIMT and D in (32, 64, 96, 128)

    # ---- byol (:508-596)
    def _forward_byol(self, clips, index, keys):
        clip_q = clips[0] if isinstance(clips[0], list) else clips
        out = self.backbone(clip_q)
        if not isinstance(out, list):
            raise NotImplementedError("BYOL: predictor is missing")
        feat_q, preds = out[0], out[1:]
        assert len(preds) == 1
        if index is None:
            return feat_q
        if not self.training:
            return self.eval_knn(self.l2_norm(feat_q))
        if keys is None:
            keys = self.compute_key_feat([list(c) for c in clips], compute_predictor_keys=False)

        def sim(pred, key):  # sim_loss(l2_norm(pred), key) in one kernel (K7)
            return ByolSimilarity.apply(pred, key, self.T, True)

        if self.cfg.CONTRASTIVE.SEQUENTIAL:
            # this view's prediction against every other view's key, averaged (:558-562)
            loss = sim(preds[0], keys[0])
            for key in keys[1:]:
                loss = loss + sim(preds[0], key)
            loss = loss / len(keys)
        else:
            # symmetric pair: view 1 predicts key 2, view 2 predicts key 1 (:572-582)
            assert len(clips) == 2
            preds2 = self.backbone(clips[1])[1:]
            assert len(preds2) == 1
            loss = sim(preds[0], keys[1]) + sim(preds2[0], keys[0])
        return self._cached_dummy_logits(len(index), feat_q.device), loss

    # ---- swav (:598-731; the public-code branch is the only live one)
    def _forward_swav(self, clips, index, epoch_exact):
        if not isinstance(clips[0], list):
            emb, _ = self.run_swav_orig_encoder_q(clips)
            if index is None:
                return emb
            if not self.training:
                return self.eval_knn(emb)
        n_crops = len(clips)
        head = getattr(self, "module", self).swav_prototypes
        with torch.no_grad():  # K9: prototype rows back onto the unit sphere, in place (:617-621)
            unit, _ = ops.l2norm_fwd(head.weight.data.contiguous(), 1e-12)
            head.weight.copy_(unit)

        bs = clips[0][0].size(0)
        # the reference encodes, normalises and scores crop by crop (:623-631, run_swav_orig_encoder_q) and concatenates
        # afterwards; the crops differ in resolution, so the backbone runs per crop, but Normalize and the prototype
        # scores are row-wise: ONE Normalize launch and ONE score GEMM over all crops (and one pair of backward GEMMs
        # instead of twelve) give the same rows
        feats = [self.backbone(c) for c in clips]
        embedding = l2norm_lastdim(torch.cat(feats, dim=0), 1e-12)
        scores = head(embedding)
        q_knn = embedding[:bs]

        # the first two crops (the large ones) get codes; every crop is scored against them (:633-679)
        self.swav_crops_for_assign = np.arange(min(2, n_crops))
        queue_live = self.cfg.CONTRASTIVE.SWAV_QEUE_LEN > 0 and epoch_exact >= 15.0
        codes = []
        with torch.no_grad():
            # the assign crops' Sinkhorn problems are independent (one 16-CTA cluster each at cfg5): the second one
            # runs on a side stream beside the first
            side = None
            if scores.is_cuda and self.cfg.NUM_SHARDS <= 1 and len(self.swav_crops_for_assign) > 1:
                main = torch.cuda.current_stream(scores.device)
                side = self._cached(("sinkhorn_stream", scores.device), lambda: torch.cuda.Stream(device=scores.device))
                side.wait_stream(main)
            for slot, crop in enumerate(self.swav_crops_for_assign):
                rows = slice(bs * crop, bs * (crop + 1))
                with torch.cuda.stream(side) if (side is not None and slot == 1) else contextlib.nullcontext():
                    out = scores[rows].detach()
                    if queue_live:
                        out = self._swav_queue_step(slot, embedding[rows].detach(), out, head.weight)
                    if self.cfg.NUM_SHARDS > 1:
                        q = self.distributed_sinkhorn(torch.exp(out / self.swav_eps_sinkhorn).t(), 3)[-bs:]
                    else:  # K10
                        q = ops.sinkhorn(out.contiguous(), self.swav_eps_sinkhorn, 3, keep_last=bs, lane=slot)
                if side is not None and slot == 1:
                    q.record_stream(main)
                codes.append(q)
            if side is not None:
                main.wait_stream(side)
        loss = SwavSwappedCe.apply(scores, torch.stack(codes, 0), n_crops, bs, self.T)  # K11
        self.knn_mem_update(q_knn, index)
        return self._cached_dummy_logits(len(index), scores.device), loss

    def _swav_queue_step(self, slot, emb, out, proto_w):
        """K12 (:642-664): once the feature queue of this crop is full its scores are prepended to the batch's,
        then the batch's embeddings enter the queue, newest first.

        The reference decides "full" by reading the queue's last row back to the host on every step until it is
        non-zero (:651).  The queue starts as zeros and every step pushes `bs` rows at the front, so it is full
        after ceil(L / bs) pushes: a host-side push counter answers the question without touching the device.  The
        counter is unknown after a checkpoint load (or any other torch-side write to the buffer, seen through its
        version counter): one read-back then re-seeds it.
        The buffer keeps the reference's physical layout (newest first) -- it is part of the checkpoint contract --
        so the push is the same front insertion; at the cfg5 size (L = 3840 / 8 ranks, D = 128: 245 KB per crop)
        that is one ~2 us copy, far below the step's other costs, and a ring with a moving head would buy nothing
        while breaking the layout."""
        queue = self.queue_swav[slot]
        bs, rows = emb.shape[0], queue.shape[0]
        if not self.swav_use_the_queue:
            tag = (self.queue_swav.data_ptr(), self.queue_swav._version)
            seen = self._const.get("swav_pushes")
            if seen is None or seen[0] != tag:
                full = bool(torch.any(queue[-1] != 0).item())  # once per (re)start
                seen = [tag, [rows if full else 0] * self.queue_swav.shape[0]]
            self.swav_use_the_queue = seen[1][slot] >= rows
            self._const["swav_pushes"] = seen
        if self.swav_use_the_queue:
            out = torch.cat((queue @ proto_w.t(), out))
        self.queue_swav[slot] = torch.cat((emb, queue[:-bs]))
        seen = self._const.get("swav_pushes")
        if seen is not None:
            seen[1][slot] = min(rows, seen[1][slot] + bs)
            seen[0] = (self.queue_swav.data_ptr(), self.queue_swav._version)  # our own write is not a foreign one
        return out

    # ---- simclr (:733-802; `distributed_loss` is False in the reference, so: gather with gradient)
    def _forward_simclr(self, clips, index):
        clip_q = clips[0] if isinstance(clips[0], list) else clips
        feat_q = self.backbone(clip_q)
        if index is None:
            return self.l2_norm(feat_q)
        if not self.training:
            return self.eval_knn(self.l2_norm(feat_q))
        feat_q2 = self.backbone(clips[1])
        # K6 + C4/C5: normalise, gather, NT-Xent over this rank's rows, gradient with the reference's
        # world-size factor (utils/distributed.py:142-155)
        loss = NtXentRows.apply(feat_q, feat_q2, self.T, self.ntxent_impl, self._simclr_exchanges(feat_q), self._status)
        if self.knn_on:
            with torch.no_grad():
                self.knn_mem_update(self.l2_norm(feat_q.detach()), index)
        return self._cached_dummy_logits(len(index), feat_q.device), loss

    def _simclr_exchanges(self, feat):
        """NVLink exchanges for the two gathers of the SimCLR branch (C4: rows and row sums over WORLD), or None
        when the ranks do not share a box / CUDA IPC is unavailable (NCCL all_gather then)."""
        if self.num_gpus <= 1 or not torch.distributed.is_initialized() or int(_opt(self.cfg, "NUM_SHARDS", 1)) != 1:
            return None
        B, D = feat.shape
        if not (feat.is_cuda and feat.dtype == torch.float32 and D % 4 == 0 and B % 4 == 0):
            return None
        if torch.distributed.get_world_size() > _lib.MAX_PEERS or not self._peer_transport_on():
            return None
        key = ("simclr", B, D)
        ex = self._peer_xchgs.get(key)
        if ex is None:
            ex = self._peer_xchgs[key] = (ops.PeerExchange(2 * B, D), ops.PeerExchange(2, B))
        return ex

    # ------------------------------------------------------------------ SwAV helpers
    def run_swav_orig_encoder_q(self, x):
        """F.normalize(backbone(x)) and its prototype scores (:865-870); the score GEMM is a plain
        library Linear."""
        emb = l2norm_lastdim(self.backbone(x), 1e-12)
        if self.swav_prototypes is None:
            return emb
        return emb, self.swav_prototypes(emb)

    def run_swav_encoder_q(self, im):
        """The non-public-code variant (:848-853): prototypes normalised on the fly."""
        emb = l2norm_lastdim(self.backbone(im), 1e-12)
        w = self.swav_prototypes.weight if isinstance(self.swav_prototypes, nn.Linear) else self.swav_prototypes.t()
        return emb, emb @ l2norm_lastdim(w.contiguous(), 1e-12).t()

    @torch.no_grad()
    def get_code(self, out):
        """Sinkhorn codes of a score matrix (:855-863)."""
        if self.cfg.NUM_SHARDS > 1:
            return self.distributed_sinkhorn(torch.exp(out / self.swav_eps_sinkhorn).t(), 3)
        return ops.sinkhorn(out.contiguous(), self.swav_eps_sinkhorn, 3)

    @torch.no_grad()
    def sinkhorn(self, Q, iters):
        """:872-887.  Takes Q = exp(scores / eps) [B, P] as the reference passes it; the kernel works from
        log Q (API parity only - `forward` hands raw scores to `ops.sinkhorn`)."""
        return ops.sinkhorn(torch.log(Q).contiguous(), 1.0, iters)

    def distributed_sinkhorn(self, Q, nmb_iters):
        """:889-910, multi-node only (NUM_SHARDS > 1): Q is [P, B_local]; row sums are all-reduced."""
        return ops.sinkhorn_distributed(Q, nmb_iters)

    def KLDivLoss(self, out, code):
        """:912-916: -mean_n sum_p code * log softmax(out / T)."""
        return SwavSwappedCe.apply(out, code.unsqueeze(0), 1, out.shape[0], self.T)


def l2_loss(x, y):
    """:919-920."""
    return 2 - 2 * (x * y).sum(dim=-1)


def contrastive_parameter_surgery(model, cfg, epoch_exact, cur_iter):
    """:1083-1116.  SwAV: prototype gradients are dropped during the first epoch.  MoCo: no parameter
    update while the queue still holds its random initialisation (one queue length of samples)."""
    is_head = cfg.MODEL.MODEL_NAME == "ContrastiveModel"
    kind = cfg.CONTRASTIVE.TYPE
    if is_head and kind == "swav" and epoch_exact <= 1.0:
        for name, p in model.named_parameters():
            if "swav_prototypes" in name:
                p.grad = None
    warm_iters = 0
    if is_head and kind == "moco":
        per_iter = cfg.TRAIN.BATCH_SIZE * cfg.NUM_SHARDS
        assert cfg.CONTRASTIVE.QUEUE_LEN % per_iter == 0
        warm_iters = cfg.CONTRASTIVE.QUEUE_LEN // cfg.TRAIN.BATCH_SIZE // cfg.NUM_SHARDS
    update_param = not (cur_iter < warm_iters and epoch_exact < 1)
    if not update_param:
        logger.info("Not updating parameters {}/{}".format(cur_iter, warm_iters))
    return model, update_param


def contrastive_forward(model, cfg, inputs, index, time, epoch_exact, scaler=None):
    """The training loop's entry point (:1119-1171; called from tools/train.py:63-77).

    SEQUENTIAL: every view takes a turn as the query against the keys of the others (moco / byol) or is
    paired with its successor (swav / simclr); each turn back-propagates at once, the summed loss is
    divided by 2 * len(inputs) (so that it matches the symmetric formulation), and MoCo enqueues all
    keys at the end.  Otherwise one forward; the caller back-propagates (`perform_backward`).
    """
    if not cfg.CONTRASTIVE.SEQUENTIAL:
        preds, partial_loss = model(inputs, index, time, epoch_exact, keys=None)
        return model, preds, partial_loss, True

    kind = cfg.CONTRASTIVE.TYPE
    core = getattr(model, "module", model)
    n_views = len(inputs)
    if kind in ("moco", "byol"):
        keys = core.compute_key_feat(inputs, compute_predictor_keys=False, batched_inference=n_views < 2)
    else:
        keys = [None] * n_views
    pairwise = kind in ("swav", "simclr")
    preds, partial_loss = None, None
    for turn in range(n_views - 1 if pairwise else n_views):
        vids = inputs[turn:turn + 2] if pairwise else [inputs[turn]]
        others = keys[:turn] + keys[turn + 1:]
        t_turn = None
        if time is not None:  # this view's time first, then the earlier and the later ones
            t_turn = torch.cat([time[:, turn:turn + 1, :], time[:, :turn, :], time[:, turn + 1:, :]], 1)
        lgt, loss = model(vids, index, t_turn, epoch_exact, keys=others)
        (scaler.scale(loss) if scaler is not None else loss).backward()
        if preds is None:
            preds, partial_loss = lgt, loss.detach()
        else:
            preds = torch.cat([preds, lgt], dim=0)
            partial_loss += loss.detach()
    partial_loss /= n_views * 2.0
    if kind == "moco":
        core._dequeue_and_enqueue(keys)
    return model, preds, partial_loss, False


try:
    MODEL_REGISTRY.register()(ContrastiveModel)
except Exception:  # already registered by the host application
    pass
