"""B200-native `ContrastiveModel` — drop-in for the reference's
`models/contrastive.py` (same class / function names, arguments, return values,
buffer names and error behaviour), with every hot op routed to the hand-written
sm_100a kernels behind `include/avssl_b200.h`.

Reference map (file:line are in /root/reference):
  ContrastiveModel.__init__            models/contrastive.py:37-129
  _update_history (K1)                 :158-172   -> ops.EmaPlan (one launch, device `iter`)
  _batch_shuffle/_batch_unshuffle      :174-230   -> all-to-all exchange (C1) / single all_gather (C3)
  _dequeue_and_enqueue (K4)            :263-292   -> ops.queue_enqueue (device `ptr`, no .item())
  compute_key_feat                     :308-371
  forward: mem / moco / byol / swav / simclr   :373-804
  sinkhorn / distributed_sinkhorn      :872-910   -> ops.sinkhorn (single cooperative kernel)
  Normalize / Memory / Memory1D        :923-1080
  contrastive_parameter_surgery        :1083-1116
  contrastive_forward                  :1119-1171

Differences that are deliberate and documented in DESIGN.md:
  * no host synchronisation on the step path (`iter`, `ptr` stay on the device; the
    reference's asserts on them become a device status word, `check_device_status()`);
  * history parameters are updated in place (the reference rebinds `.data`);
  * `logits` is returned detached (the loss carries the gradient), and can be skipped
    with `materialize_logits = False`;
  * the dummy logits of byol/swav/simclr are a cached device constant (K8);
  * SwAV prototype count is `cfg.CONTRASTIVE.SWAV_NUM_PROTOTYPES` if present, else the
    reference's hard-coded 1000.
The backbones are not part of this package: `_MODEL_TYPES` is filled from the host
application's `models.video_model_builder` when it is importable (i.e. inside the
reference tree), or by the caller.
"""
import logging
import math

import numpy as np
import torch
import torch.nn as nn

from . import _lib, ops
from . import distributed as du
from . import losses

logger = logging.getLogger(__name__)

# Supported model types (models/contrastive.py:20-28); filled lazily, see module doc.
_MODEL_TYPES = {}
try:  # inside the reference tree the video backbones are used unchanged
    from models.video_model_builder import X3D, MViT, ResNet, SlowFast  # type: ignore

    _MODEL_TYPES.update({"slowfast": SlowFast, "slow": ResNet, "c2d": ResNet, "i3d": ResNet,
                         "slow_c2d": ResNet, "x3d": X3D, "mvit": MViT})
except Exception:  # pragma: no cover - standalone use: the caller registers backbones
    pass

try:
    from models.build import MODEL_REGISTRY  # type: ignore
except Exception:  # pragma: no cover
    class _Registry(dict):
        def register(self, obj=None):
            def deco(o):
                self[o.__name__] = o
                return o
            return deco(obj) if obj is not None else deco

        def get(self, name):
            return self[name]

    MODEL_REGISTRY = _Registry()


def _cfg_get(node, name, default):
    try:
        return getattr(node, name)
    except (AttributeError, KeyError):
        return default


# ------------------------------------------------------------------ autograd bridges
class _L2NormFn(torch.autograd.Function):
    """y = x / max(||x||, eps) per row (K2 forward / backward kernels)."""

    @staticmethod
    def forward(ctx, x, eps):
        y, nrm = ops.l2norm_fwd(x.contiguous(), eps)
        ctx.save_for_backward(y, nrm)
        ctx.eps = eps
        return y

    @staticmethod
    def backward(ctx, dy):
        y, nrm = ctx.saved_tensors
        return ops.l2norm_bwd(y, nrm, dy, ctx.eps), None


def _l2norm_rows(x, eps=0.0):
    """Row-normalise the last dim of a >=2-D tensor through the CUDA kernel."""
    shp = x.shape
    y = _L2NormFn.apply(x.reshape(-1, shp[-1]), eps)
    return y.reshape(shp)


class _MocoInfoNceFn(torch.autograd.Function):
    """Fused l2-norm + logits + InfoNCE, forward and backward in one pass (K2+K3)."""

    @staticmethod
    def forward(ctx, feat_q, queue, T, want_logits, impl, enqueue, *keys):
        # enqueue = (ptr, status) folds K4 for keys[0] into the same launch (after the loss
        # has been computed against the old queue, models/contrastive.py:486-503)
        out = ops.moco_infonce(feat_q.detach().contiguous(), [k.detach().contiguous() for k in keys],
                               queue, T, want_logits=want_logits, impl=impl, enqueue=enqueue)
        ctx.save_for_backward(out["dfeat"])
        logits = out["logits"] if want_logits else feat_q.new_empty(0)
        ctx.mark_non_differentiable(logits, out["q"])
        return out["loss"].reshape(()), logits, out["q"]

    @staticmethod
    def backward(ctx, g_loss, g_logits, g_q):
        (dfeat,) = ctx.saved_tensors
        return (dfeat * g_loss,) + (None,) * (len(ctx.needs_input_grad) - 1)


class _ByolSimFn(torch.autograd.Function):
    """-mean(p.k)/T with the predictor l2-norm fused in (K7)."""

    @staticmethod
    def forward(ctx, pred, key, T, normalize):
        loss, dpred = ops.byol_simloss(pred.detach().contiguous(), key.detach().contiguous(), T,
                                       normalize=normalize, want_grad=True)
        ctx.save_for_backward(dpred)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        (dpred,) = ctx.saved_tensors
        return dpred * g, None, None, None


class _NtXentFn(torch.autograd.Function):
    """SimCLR NT-Xent over this rank's rows against all gathered columns (K6 + C4/C5)."""

    @staticmethod
    def forward(ctx, feat1, feat2, T, impl):
        loss, d1, d2 = ops.ntxent(feat1.detach().contiguous(), feat2.detach().contiguous(), T, impl=impl)
        ctx.save_for_backward(d1, d2)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        d1, d2 = ctx.saved_tensors
        return d1 * g, d2 * g, None, None


class _SwavCeFn(torch.autograd.Function):
    """SwAV soft-target cross-entropy over all (assign crop, other crop) pairs (K11)."""

    @staticmethod
    def forward(ctx, output, codes, n_crops, bs, T):
        loss, dout = ops.swav_ce(output.detach().contiguous(), codes, n_crops, bs, T)
        ctx.save_for_backward(dout)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        (dout,) = ctx.saved_tensors
        return dout * g, None, None, None, None


# ------------------------------------------------------------------------ the model
class ContrastiveModel(nn.Module):
    """Contrastive head in its mem / moco / byol / swav / simclr modes."""

    def __init__(self, cfg):
        super(ContrastiveModel, self).__init__()
        self.backbone = _MODEL_TYPES[cfg.MODEL.ARCH](cfg)
        self.type = cfg.CONTRASTIVE.TYPE
        self.T = cfg.CONTRASTIVE.T
        self.dim = cfg.CONTRASTIVE.DIM
        self.length = cfg.CONTRASTIVE.LENGTH
        self.k = cfg.CONTRASTIVE.QUEUE_LEN
        self.mmt = cfg.CONTRASTIVE.MOMENTUM
        self.momentum_annealing = cfg.CONTRASTIVE.MOMENTUM_ANNEALING
        self.duration = 1
        self.cfg = cfg
        self.num_gpus = cfg.NUM_GPUS
        self.l2_norm = Normalize()
        self.knn_num_imgs = 0
        self.knn_on = cfg.CONTRASTIVE.KNN_ON
        self.train_labels = np.zeros((0,), dtype=np.int32)
        self.num_pos = 2
        self.num_crops = self.cfg.DATA.TRAIN_CROP_NUM_TEMPORAL * self.cfg.DATA.TRAIN_CROP_NUM_SPATIAL
        self.nce_loss_fun = losses.get_loss_func("contrastive_loss")(reduction="mean")
        self.softmax = nn.Softmax(dim=1)
        # B200-path knobs (not in the reference)
        self.materialize_logits = True
        self.infonce_impl = _lib.IMPL_AUTO
        self.ntxent_impl = _lib.IMPL_AUTO  # tcgen05 (tf32, both operands rounded to nearest) when D allows; IMPL_SIMT = exact fp32
        self._ema_plan = None
        self._iter_mirror = None
        self._dummy_logits = None
        # C3 over NVLink peer memory instead of NCCL (ops.PeerExchange); opt-in because it needs all
        # ranks of the (local) group on one box: cfg.CONTRASTIVE.PEER_EXCHANGE or enable_peer_exchange().
        self._peer_exchange_on = bool(_cfg_get(cfg.CONTRASTIVE, "PEER_EXCHANGE", False))
        self._peer_xchgs = {}
        self.register_buffer("_status", torch.zeros(1, dtype=torch.int32), persistent=False)

        if self.type == "mem":
            self.mem_type = cfg.CONTRASTIVE.MEM_TYPE
            if self.mem_type == "1d":
                self.memory = Memory1D(self.length, self.duration, self.dim, cfg)
            else:
                self.memory = Memory(self.length, self.duration, self.dim, cfg)
            self.examplar_type = "video"
            self.interp = cfg.CONTRASTIVE.INTERP_MEMORY
        elif self.type == "self":
            pass
        elif self.type == "moco" or self.type == "byol":
            self.backbone_hist = _MODEL_TYPES[cfg.MODEL.ARCH](cfg)
            for p in self.backbone_hist.parameters():
                p.requires_grad = False
            self.register_buffer("ptr", torch.tensor([0]))
            self.ptr.requires_grad = False
            stdv = 1.0 / math.sqrt(self.dim / 3)
            self.register_buffer("queue_x", torch.rand(self.k, self.dim).mul_(2 * stdv).add_(-stdv))
            self.register_buffer("iter", torch.zeros([1], dtype=torch.long))
            self._batch_shuffle_on = (
                False
                if ("sync" in cfg.BN.NORM_TYPE and cfg.BN.NUM_SYNC_DEVICES == cfg.NUM_GPUS)
                or self.type == "byol"
                else True
            )
        elif self.type == "swav":
            self.swav_use_public_code = True
            n_proto = int(_cfg_get(cfg.CONTRASTIVE, "SWAV_NUM_PROTOTYPES", 1000))
            self.swav_prototypes = nn.Linear(self.dim, n_proto, bias=False)
            self.swav_eps_sinkhorn = 0.05
            self.swav_use_the_queue = False
            if self.cfg.CONTRASTIVE.SWAV_QEUE_LEN > 0:
                self.register_buffer(
                    "queue_swav",
                    torch.zeros(2, self.cfg.CONTRASTIVE.SWAV_QEUE_LEN // du.get_world_size(), self.dim))
        elif self.type == "simclr":
            # the reference precomputes float64 pos/neg masks here (:806-846) that its
            # live loss never reads (distributed_loss=False, :748).
            self.pos_mask, self.neg_mask = [], None
        self.simclr_dist_on = cfg.CONTRASTIVE.SIMCLR_DIST_ON

        if self.knn_on:
            self.knn_mem = Memory(self.length, 1, self.dim, cfg)

    # -------------------------------------------------------------- housekeeping
    def _apply(self, fn, *args, **kwargs):
        # .cuda()/.to()/.float() move parameter storage: drop cached pointer tables
        self._ema_plan = None
        self._iter_mirror = None
        self._dummy_logits = None
        return super(ContrastiveModel, self)._apply(fn, *args, **kwargs)

    def enable_peer_exchange(self, on=True):
        """Route the key gather of `_batch_unshuffle` (C3) through NVLink peer stores + epoch flags
        (ops.PeerExchange) instead of an NCCL all_gather.  Collective: call it on every rank; all
        ranks of the shuffle group must sit on one box."""
        self._peer_exchange_on = bool(on)
        if not on:
            for ex in self._peer_xchgs.values():
                ex.close()
            self._peer_xchgs = {}
        return self

    def _peer_xchg(self, rows, dim, local):
        key = (rows, dim, bool(local))
        ex = self._peer_xchgs.get(key)
        if ex is None:
            ex = ops.PeerExchange(rows, dim, group=du._LOCAL_PROCESS_GROUP if local else None)
            self._peer_xchgs[key] = ex
        return ex

    def check_device_status(self):
        """Host check of the device status word that replaces the reference's host
        asserts on `ptr` / bank indices (synchronises; call it off the step path)."""
        flags = int(self._status.item())
        assert not (flags & _lib.DEVFLAG_QUEUE_OVERRUN), "queue overrun: ptr + n > K (models/contrastive.py:285)"
        if flags & _lib.DEVFLAG_BAD_INDEX:
            raise IndexError("memory-bank index out of range")
        return flags

    def _cached_dummy_logits(self, n, device):
        """K8: [n, K+1] zeros with column 0 = 9999 (models/contrastive.py:585-592),
        built once on the device instead of on the CPU every step."""
        d = self._dummy_logits
        if d is None or d.shape[0] != n or d.device != device:
            d = torch.zeros(n, self.k + 1, dtype=torch.float, device=device)
            d[:, 0] = 9999.0
            self._dummy_logits = d
        return d

    # --------------------------------------------------------------------- kNN bank
    @torch.no_grad()
    def knn_mem_update(self, q_knn, index):
        if self.knn_on:
            self.knn_mem.update(q_knn, momentum=1.0, ind=index, time=torch.zeros_like(index), interp=False,
                                status=self._status)

    @torch.no_grad()
    def init_knn_labels(self, train_loader):
        logger.info("initializing knn labels")
        self.num_imgs = len(train_loader.dataset._labels)
        self.train_labels = np.zeros((self.num_imgs,), dtype=np.int32)
        for i in range(self.num_imgs):
            self.train_labels[i] = train_loader.dataset._labels[i]
        self.train_labels = torch.LongTensor(self.train_labels).to(self.knn_mem.memory.device)
        if self.length != self.num_imgs:
            logger.error("Kinetics dataloader size: {} differs from memorybank length {}".format(
                self.num_imgs, self.length))
            self.knn_mem.resize(self.num_imgs, 1, self.dim)

    @torch.no_grad()
    def eval_knn(self, q_knn, knn_k=200):
        # eval-only: stock cuBLAS + topk
        dist = torch.einsum("nc,mc->nm", q_knn.view(q_knn.size(0), -1),
                            self.knn_mem.memory.view(self.knn_mem.memory.size(0), -1))
        yd, yi = dist.topk(knn_k, dim=1, largest=True, sorted=True)
        return yd, yi

    # ------------------------------------------------------------------- K1: EMA
    def _ema_lists(self):
        online = dict(self.backbone.named_parameters())
        o_list, h_list = [], []
        for name, p in self.backbone_hist.named_parameters():
            o_list.append(online[name].data)
            h_list.append(p.data)
        return o_list, h_list

    @torch.no_grad()
    def _update_history(self, _bump_iter=False):
        """Momentum update of the key encoder (models/contrastive.py:158-172) as one
        multi-tensor launch; `iter` is read on the device (no int(self.iter) sync)."""
        if self._ema_plan is None:
            o_list, h_list = self._ema_lists()
            self._ema_plan = ops.EmaPlan(o_list, h_list)
        # host mirror of `iter`: our kernels bump the buffer through its raw pointer, which leaves
        # the tensor's version counter alone; any torch-side write (load_state_dict, zero_(), ...)
        # changes it and forces one re-read.  Steady state: no D2H sync at all.
        key = (self.iter.data_ptr(), self.iter._version)
        if self._iter_mirror is None or self._iter_mirror[0] != key:
            self._iter_mirror = [key, int(self.iter.item())]
        self._ema_plan.run(self.mmt, self.iter, bump_iter=_bump_iter, first_iter=self._iter_mirror[1] == 0)
        if _bump_iter:
            self._iter_mirror[1] += 1

    # ------------------------------------------------------ shuffle BN (A6, C1-C3)
    @torch.no_grad()
    def _batch_shuffle(self, x):
        if len(x) == 2:
            another_crop = True
        else:
            another_crop = False
        if another_crop:
            x, x_crop = x[0], x[1]
        else:
            x = x[0]

        world_size = self.cfg.NUM_GPUS * self.cfg.NUM_SHARDS
        bsz = x.shape[0]
        if self.num_gpus > 1:
            if self.cfg.CONTRASTIVE.LOCAL_SHUFFLE_BN:
                world_size = du.get_local_size()
                gpu_idx = du.get_local_rank()
            else:
                gpu_idx = torch.distributed.get_rank()
            n_total = bsz * world_size
        else:
            n_total = bsz

        idx_randperm = torch.randperm(n_total).to(x.device)  # rank 0's CPU RNG, as the reference
        if self.num_gpus > 1:
            torch.distributed.broadcast(idx_randperm, src=0)
        else:
            gpu_idx = 0
        idx_randperm = idx_randperm.view(world_size, -1)
        if self.num_gpus > 1:
            # C1 as an all-to-all: only the rows this rank keeps cross the fabric
            # (the reference all_gathers the whole batch and discards (W-1)/W of it).
            x = _exchange_rows(x, idx_randperm, gpu_idx, world_size)
            if another_crop:
                x_crop = _exchange_rows(x_crop, idx_randperm, gpu_idx, world_size)
        else:
            x = x[idx_randperm[gpu_idx, :]]
            if another_crop:
                x_crop = x_crop[idx_randperm[gpu_idx, :]]

        idx_restore = torch.argsort(idx_randperm.view(-1))
        idx_restore = idx_restore.view(world_size, -1)
        if another_crop:
            return [x, x_crop], idx_restore
        else:
            return [x], idx_restore

    @torch.no_grad()
    def _batch_unshuffle(self, x, idx_restore):
        if (self.num_gpus > 1 and self._peer_exchange_on and x.is_cuda and x.dim() == 2
                and x.dtype == torch.float32 and x.shape[1] % 4 == 0):
            # C3 without a collective kernel: push this rank's keys into every peer's buffer, then
            # wait + select idx_restore[rank] in one small launch (bit-identical to the path below)
            local = bool(self.cfg.CONTRASTIVE.LOCAL_SHUFFLE_BN)
            gpu_idx = du.get_local_rank() if local else torch.distributed.get_rank()
            ex = self._peer_xchg(x.shape[0], x.shape[1], local)
            ex.push(x.contiguous())
            return ex.wait_gather(idx_restore[gpu_idx, :].contiguous(), status=self._status)
        if self.num_gpus > 1:
            if self.cfg.CONTRASTIVE.LOCAL_SHUFFLE_BN:
                x = du.cat_all_gather(x, local=True)
                gpu_idx = du.get_local_rank()
            else:
                x = du.cat_all_gather(x)
                gpu_idx = torch.distributed.get_rank()
        else:
            gpu_idx = 0
        idx = idx_restore[gpu_idx, :]
        x = x[idx]
        return x

    # ---------------------------------------------------------------------- BYOL
    def sim_loss(self, q, k):
        """models/contrastive.py:243-249 (q already normalised)."""
        return _ByolSimFn.apply(q, k, self.T, False)

    @torch.no_grad()
    def momentum_anneal_cosine(self, epoch_exact):
        self.mmt = (1 - (1 - self.cfg.CONTRASTIVE.MOMENTUM)
                    * (math.cos(math.pi * epoch_exact / self.cfg.SOLVER.MAX_EPOCH) + 1.0) * 0.5)

    # ----------------------------------------------------------------- K4: queue
    @torch.no_grad()
    def _dequeue_and_enqueue(self, keys, extra_keys=None):
        if not self.cfg.CONTRASTIVE.MOCO_MULTI_VIEW_QUEUE:
            keys_queue_update = [keys[0]]
        else:
            assert len(keys) > 0, "need to have multiple views for adding them to queue"
            keys_queue_update = []
            keys_queue_update += keys
            if extra_keys:
                keys_queue_update += [item for sublist in extra_keys for item in sublist]
        for key in keys_queue_update:
            num_items = int(key.size(0))
            assert self.k % num_items == 0
            # `assert ptr + num_items <= self.k` (:285) is evaluated on the device
            ops.queue_enqueue(self.queue_x, self.ptr, key.detach().contiguous(), self._status)

    @torch.no_grad()
    def batch_clips(self, clips):
        clips_batched = [None] * len(clips[0])
        for i, clip in enumerate(clips):
            for j, view in enumerate(clip):
                if i == 0:
                    clips_batched[j] = view
                else:
                    clips_batched[j] = torch.cat([clips_batched[j], view], dim=0)
                del view
        return clips_batched

    @torch.no_grad()
    def compute_key_feat(self, clips_k, compute_predictor_keys=False, batched_inference=True):
        assert self.training
        # momentum update key encoder + `self.iter += 1` (:313-314), one launch
        self._update_history(_bump_iter=True)
        n_clips = len(clips_k)
        bsz = clips_k[0][0].shape[0]
        if n_clips * bsz * clips_k[0][0].numel() > 4 * 64 * 3 * 8 * 224 * 224:
            batched_inference = False  # hack to avoid oom on large inputs
        assert n_clips > 0
        if batched_inference and all(
            [clips_k[i][j].shape[1:] == clips_k[0][j].shape[1:]
             for i in range(len(clips_k)) for j in range(len(clips_k[i]))]):
            clips_k = [self.batch_clips(clips_k)]
            batched = True
        else:
            batched = False

        keys, pred_keys = [], []
        for k in range(0, len(clips_k)):
            clip_k = clips_k[k]
            if self._batch_shuffle_on:
                clip_k, idx_restore = self._batch_shuffle(clip_k)
            hist_feat = self.backbone_hist(clip_k)
            if isinstance(hist_feat, list):
                hist_time = hist_feat[1:]
                hist_feat = hist_feat[0]
                if compute_predictor_keys:
                    tks = []
                    for tk in hist_time:
                        tk = self.l2_norm(tk)
                        if self._batch_shuffle_on:
                            tk = self._batch_unshuffle(tk, idx_restore).detach()
                        tks.append(tk)
                    pred_keys.append(tks)
            x_hist = self.l2_norm(hist_feat)
            if self._batch_shuffle_on:
                x_hist = self._batch_unshuffle(x_hist, idx_restore).detach()
            keys.append(x_hist)
        if batched:
            assert len(keys) == 1, "batched input uses single clip"
            batched_key = keys[0]
            if compute_predictor_keys:
                batched_pred_key = pred_keys[0]
            keys, pred_keys = [], []
            for k in range(0, n_clips):
                keys.append(batched_key[k * bsz:(k + 1) * bsz])
                if compute_predictor_keys:
                    pred_keys.append(batched_pred_key[k * bsz:(k + 1) * bsz])
        if compute_predictor_keys:
            return keys, pred_keys
        else:
            return keys

    # -------------------------------------------------------------------- forward
    def forward(self, clips, index=None, time=None, epoch_exact=None, keys=None):
        if epoch_exact is not None and self.momentum_annealing:
            self.momentum_anneal_cosine(epoch_exact)

        if self.type == "mem":
            return self._forward_mem(clips, index, time)
        elif self.type == "moco":
            return self._forward_moco(clips, index, time, keys)
        elif self.type == "byol":
            return self._forward_byol(clips, index, keys)
        elif self.type == "swav":
            return self._forward_swav(clips, index, epoch_exact)
        elif self.type == "simclr":
            return self._forward_simclr(clips, index)
        else:
            raise NotImplementedError()

    # models/contrastive.py:379-442
    def _forward_mem(self, clips, index, time):
        batch_size = clips[0].size(0)
        q = self.backbone(clips)
        if index is None:
            return q
        q = self.l2_norm(q)
        if not self.training:
            assert self.knn_mem.duration == 1
            return self.eval_knn(q)
        time *= self.duration - 1
        # negatives come from the CPU generator exactly as in the reference (:390-397),
        # so a seeded run draws the same indices (RNG parity).
        clip_ind = torch.randint(0, self.length, size=(batch_size, self.k + 1)).to(q.device)
        clip_ind.select(1, 0).copy_(index.data)
        if self.mem_type == "2d":
            if self.interp:
                time_ind = torch.empty(batch_size, self.k + 1).uniform_(0, self.duration - 1).to(q.device)
            else:
                time_ind = torch.randint(0, self.duration - 1, size=(batch_size, self.k + 1)).to(q.device)
        else:
            time_ind = torch.zeros(size=(batch_size, self.k + 1), dtype=int).to(q.device)
        if self.examplar_type == "clip":
            time_ind.select(1, 0).copy_(time.data)
        elif self.examplar_type == "video":
            pass
        else:
            raise NotImplementedError("unsupported examplar_type {}".format(self.examplar_type))
        # K14: q . bank[ind] / T without the [B, K+1, D] gather
        prod = _MemDotFn.apply(q, self.memory.memory, clip_ind, time_ind, self.T,
                               bool(self.interp) and self.mem_type == "2d", self._status)
        loss = self.nce_loss_fun(prod)  # computed and dropped, as in the reference (:436,442)
        del loss
        self.memory.update(q, momentum=self.mmt, ind=index, time=time, interp=self.interp, status=self._status)
        self.knn_mem_update(q, index)
        return prod, 0.0, True

    # models/contrastive.py:443-506
    def _forward_moco(self, clips, index, time, keys):
        if isinstance(clips[0], list):
            n_clips = len(clips)
            ind_clips = np.arange(n_clips)
            clip_q = clips[ind_clips[0]]
            clips_k = [clips[i] for i in ind_clips[1:]]
            time_q = time[:, ind_clips[0], :]  # noqa: F841 (kept: raises like the reference when time is None)
            time_k = (time[:, ind_clips[1:], :] if keys is None else time[:, ind_clips[0] + 1:, :])  # noqa: F841
        else:
            clip_q = clips

        feat_q = self.backbone(clip_q)
        extra_projs = []
        if isinstance(feat_q, list):
            extra_projs = feat_q[1:]
            feat_q = feat_q[0]
            extra_projs = [self.l2_norm(feat) for feat in extra_projs]  # noqa: F841

        if index is None:
            return feat_q
        if not self.training:
            return self.eval_knn(self.l2_norm(feat_q))

        if keys is None:
            keys = self.compute_key_feat(clips_k, compute_predictor_keys=False)
            auto_enqueue_keys = True
        else:
            auto_enqueue_keys = False

        # K2+K3: q = l2norm(feat_q); logits = [q.k, q.queue^T]/T; InfoNCE fwd + bwd
        # K4 rides in the same launch when exactly keys[0] is enqueued (the default,
        # CONTRASTIVE.MOCO_MULTI_VIEW_QUEUE off): the ring write happens behind a grid-wide
        # barrier after the last read of the queue.
        enqueue_here = self.training and auto_enqueue_keys
        fused = (enqueue_here and not self.cfg.CONTRASTIVE.MOCO_MULTI_VIEW_QUEUE
                 and keys[0].shape == feat_q.shape and self.k % int(keys[0].size(0)) == 0)
        loss, logits, q = _MocoInfoNceFn.apply(feat_q, self.queue_x, self.T, self.materialize_logits,
                                               self.infonce_impl, (self.ptr, self._status) if fused else None,
                                               *keys)
        if not self.materialize_logits:
            logits = None
        if enqueue_here and not fused:
            self._dequeue_and_enqueue(keys)
        self.knn_mem_update(q, index)
        return logits, loss

    # models/contrastive.py:508-596
    def _forward_byol(self, clips, index, keys):
        clips_key = [None] * len(clips)
        for i, clip in enumerate(clips):
            p = []
            for path in clip:
                p.append(path)
            clips_key[i] = p
        if isinstance(clips[0], list):
            n_clips = len(clips)
            clip_q = clips[0]
        else:
            clip_q = clips

        feat_q = self.backbone(clip_q)
        if isinstance(feat_q, list):
            predictors = feat_q[1:]
            feat_q = feat_q[0]
        else:
            raise NotImplementedError("BYOL: predictor is missing")
        assert len(predictors) == 1
        if index is None:
            return feat_q
        if not self.training:
            return self.eval_knn(self.l2_norm(feat_q))

        if keys is None:
            keys = self.compute_key_feat(clips_key, compute_predictor_keys=False)

        # sim_loss(l2_norm(pred), key) with the normalisation fused into the kernel (K7)
        if self.cfg.CONTRASTIVE.SEQUENTIAL:
            loss_reg = _ByolSimFn.apply(predictors[0], keys[0], self.T, True)
            for i in range(1, len(keys)):
                loss_reg = loss_reg + _ByolSimFn.apply(predictors[0], keys[i], self.T, True)
            loss_reg = loss_reg / len(keys)
        else:
            loss_q1 = _ByolSimFn.apply(predictors[0], keys[1], self.T, True)
            assert len(clips) == 2
            clip_q2 = clips[1]
            feat_q2 = self.backbone(clip_q2)
            predictors2 = feat_q2[1:]
            assert len(predictors2) == 1
            loss_q2 = _ByolSimFn.apply(predictors2[0], keys[0], self.T, True)
            loss_reg = loss_q1 + loss_q2

        dummy_logits = self._cached_dummy_logits(len(index), feat_q.device)
        return dummy_logits, loss_reg

    # models/contrastive.py:598-731 (public-code branch, the only live one)
    def _forward_swav(self, clips, index, epoch_exact):
        if not isinstance(clips[0], list):
            proj_1, _ = self.run_swav_orig_encoder_q(clips)
            if index is None:
                return proj_1
            if not self.training:
                return self.eval_knn(proj_1)
        n_clips = len(clips)

        # K9: prototype rows to unit l2, in place (:617-621)
        with torch.no_grad():
            m = getattr(self, "module", self)
            w, _ = ops.l2norm_fwd(m.swav_prototypes.weight.data.contiguous(), 1e-12)
            m.swav_prototypes.weight.copy_(w)

        bs = clips[0][0].size(0)
        output, embedding = [], []
        for i, clip_q in enumerate(clips):
            x = self.run_swav_orig_encoder_q(clip_q)
            embedding.append(x[0])
            output.append(x[1])
        q_knn = embedding[0]
        embedding = torch.cat(embedding, dim=0)
        output = torch.cat(output, dim=0)

        swav_extra_crops = n_clips - 2
        self.swav_crops_for_assign = np.arange(n_clips - swav_extra_crops)
        codes = []
        for i, crop_id in enumerate(self.swav_crops_for_assign):
            with torch.no_grad():
                out = output[bs * crop_id:bs * (crop_id + 1)].detach()
                if self.cfg.CONTRASTIVE.SWAV_QEUE_LEN > 0 and epoch_exact >= 15.0:
                    # (:651-653) one host sync, only while the queue is still filling
                    if self.swav_use_the_queue or not torch.all(self.queue_swav[i, -1, :] == 0):
                        self.swav_use_the_queue = True
                        out = torch.cat((torch.mm(self.queue_swav[i], m.swav_prototypes.weight.t()), out))
                    # K12: FIFO shift by bs, newest first (:659-664)
                    self.queue_swav[i] = torch.cat(
                        (embedding[crop_id * bs:(crop_id + 1) * bs].detach(), self.queue_swav[i, :-bs]))
                # K10: Q = exp(out/eps)^T, 3 Sinkhorn-Knopp iterations, last bs rows
                if self.cfg.NUM_SHARDS > 1:
                    q = self.distributed_sinkhorn(torch.exp(out / self.swav_eps_sinkhorn).t(), 3)[-bs:]
                else:
                    q = ops.sinkhorn(out.contiguous(), self.swav_eps_sinkhorn, 3, keep_last=bs)
            codes.append(q)
        # K11: all (assign crop, other crop) soft-target cross-entropies, fwd + bwd
        loss_swav = _SwavCeFn.apply(output, torch.stack(codes, 0), n_clips, bs, self.T)
        self.knn_mem_update(q_knn, index)
        dummy_logits = self._cached_dummy_logits(len(index), output.device)
        return dummy_logits, loss_swav

    # models/contrastive.py:733-802 (live branch: distributed_loss=False, gather with gradient)
    def _forward_simclr(self, clips, index):
        if isinstance(clips[0], list):
            clip_q = clips[0]
        else:
            clip_q = clips
        feat_q = self.backbone(clip_q)
        if index is None:
            return self.l2_norm(feat_q)
        if not self.training:
            return self.eval_knn(self.l2_norm(feat_q))
        feat_q2 = self.backbone(clips[1])
        # K6 (+C4/C5): l2-norm, all_gather, NT-Xent rows of this rank, gradient incl. the
        # reference's world-size factor (utils/distributed.py:142-155)
        loss = _NtXentFn.apply(feat_q, feat_q2, self.T, self.ntxent_impl)
        with torch.no_grad():
            q_knn = self.l2_norm(feat_q.detach())
        self.knn_mem_update(q_knn, index)
        dummy_logits = self._cached_dummy_logits(len(index), feat_q.device)
        return dummy_logits, loss

    def _simclr_precompute_pos_neg_mask_multi(self):
        """models/contrastive.py:806-846 builds masks only the dead branch (:749-768) reads."""
        self.pos_mask, self.neg_mask = [], None

    # ------------------------------------------------------------------ SwAV helpers
    def run_swav_encoder_q(self, im):
        """models/contrastive.py:848-853 (non-public-code variant; prototypes as a matrix)."""
        proj = self.backbone(im)
        proj = _l2norm_rows(proj, 1e-12)
        w = self.swav_prototypes.weight.t() if isinstance(self.swav_prototypes, nn.Linear) else self.swav_prototypes
        protos = _l2norm_rows(w.t().contiguous(), 1e-12).t()
        out = proj @ protos
        return proj, out

    @torch.no_grad()
    def get_code(self, out):
        """models/contrastive.py:855-863."""
        if self.cfg.NUM_SHARDS > 1:
            return self.distributed_sinkhorn(torch.exp(out / self.swav_eps_sinkhorn).t(), 3)
        return ops.sinkhorn(out.contiguous(), self.swav_eps_sinkhorn, 3)

    def run_swav_orig_encoder_q(self, x):
        """models/contrastive.py:865-870: F.normalize (eps 1e-12) + bias-free prototype Linear."""
        x = self.backbone(x)
        x = _l2norm_rows(x, 1e-12)
        if self.swav_prototypes is not None:
            return x, self.swav_prototypes(x)  # plain library GEMM (cuBLAS), autograd as usual
        return x

    @torch.no_grad()
    def sinkhorn(self, Q, iters):
        """models/contrastive.py:872-887.  Q: [B, P] = exp(scores / eps), as the reference
        passes it; the kernel works on log Q so the exp is undone here (API parity only —
        the forward path hands raw scores to ops.sinkhorn directly)."""
        return ops.sinkhorn(torch.log(Q).contiguous(), 1.0, iters)

    def distributed_sinkhorn(self, Q, nmb_iters):
        """models/contrastive.py:889-910 (multi-node only, NUM_SHARDS > 1).  Q: [P, B_local].
        Same kernel; the row sums are all-reduced between the two halves of an iteration."""
        return ops.sinkhorn_distributed(Q, nmb_iters)

    def KLDivLoss(self, out, code):
        """models/contrastive.py:912-916."""
        return _SwavCeFn.apply(out, code.unsqueeze(0), 1, out.shape[0], self.T)


def l2_loss(x, y):
    return 2 - 2 * (x * y).sum(dim=-1)


class _MemDotFn(torch.autograd.Function):
    """prod = q . bank[ind, time] / T (K14).  Backward recomputes the gather."""

    @staticmethod
    def forward(ctx, q, bank, ind, time, T, interp, status):
        prod = ops.membank_gather_dot(bank, q.detach().contiguous(), ind, time, T, interp=interp, status=status)
        ctx.save_for_backward(bank, ind, time)
        ctx.T, ctx.interp = T, interp
        return prod

    @staticmethod
    def backward(ctx, g):
        # rarely used (the reference drops the mem-mode loss): plain torch gather-matmul
        bank, ind, time = ctx.saved_tensors
        B = ind.shape[0]
        b3 = bank if bank.dim() == 3 else bank.unsqueeze(1)
        if ctx.interp:
            t0 = time.floor().long().clamp(0, b3.shape[1] - 1)
            t1 = (t0 + 1).clamp(0, b3.shape[1] - 1)
            w1 = 1 - (time - t0).reshape(-1, 1).float()
            sel = b3[ind.reshape(-1), t0.reshape(-1)] * (1 - w1) + b3[ind.reshape(-1), t1.reshape(-1)] * w1
        else:
            sel = b3[ind.reshape(-1), time.long().reshape(-1)]
        sel = sel.view(B, -1, b3.shape[-1])
        dq = torch.einsum("nk,nkc->nc", g, sel) / ctx.T
        return dq, None, None, None, None, None, None


class Normalize(nn.Module):
    """models/contrastive.py:923-934: x / (sum x^2)^(1/2) along `dim`, no eps."""

    def __init__(self, power=2, dim=1):
        super(Normalize, self).__init__()
        self.dim = dim
        self.power = power

    def forward(self, x):
        if self.power != 2:
            raise NotImplementedError("Normalize: only power=2 has a CUDA kernel")
        d = self.dim if self.dim >= 0 else x.dim() + self.dim
        if d == x.dim() - 1:
            return _l2norm_rows(x, 0.0)
        xt = x.transpose(d, -1).contiguous()
        return _l2norm_rows(xt, 0.0).transpose(d, -1)


class Memory(nn.Module):
    """models/contrastive.py:937-1039: [length, duration, dim] bank."""

    def __init__(self, length, duration, dim, cfg):
        super(Memory, self).__init__()
        self.length = length
        self.duration = duration
        self.dim = dim
        stdv = 1.0 / math.sqrt(dim / 3)
        self.register_buffer("memory", torch.rand(length, duration, dim).mul_(2 * stdv).add_(-stdv))
        self.device = self.memory.device
        self.l2_norm = Normalize(dim=1)
        self.l2_norm2d = Normalize(dim=2)
        self.num_gpus = cfg.NUM_GPUS

    def resize(self, length, duration, dim):
        self.length = length
        self.duration = duration
        self.dim = dim
        stdv = 1.0 / math.sqrt(dim / 3)
        dev = self.memory.device
        del self.memory
        self.memory = torch.rand(length, duration, dim).mul_(2 * stdv).add_(-stdv).to(dev)

    def get(self, ind, time, interp=False):
        """Row gather (optionally time-interpolated), models/contrastive.py:966-987."""
        batch_size = ind.size(0)
        with torch.no_grad():
            if interp:
                t0 = time.floor().long()
                t0 = torch.clamp(t0, 0, self.memory.shape[1] - 1)
                t1 = torch.clamp(t0 + 1, 0, self.memory.shape[1] - 1)
                mem_t0 = self.memory[ind.view(-1), t0.view(-1), :]
                mem_t1 = self.memory[ind.view(-1), t1.view(-1), :]
                w_t1 = 1 - (time - t0).view(-1, 1).float()
                selected_mem = mem_t0 * (1 - w_t1) + mem_t1 * w_t1
            else:
                selected_mem = self.memory[ind.view(-1), time.long().view(-1), :]
        return selected_mem.view(batch_size, -1, self.dim)

    def update(self, mem, momentum, ind, time, interp=False, status=None):
        """models/contrastive.py:989-1036: all_gather (C8) then the fused
        gather-lerp-normalise-scatter kernel (K5)."""
        if self.num_gpus > 1:
            mem, ind, time = du.all_gather([mem, ind, time])
        with torch.no_grad():
            ops.membank_update(self.memory, mem.detach().reshape(mem.size(0), -1).contiguous(), ind, time,
                               momentum, interp=interp, status=status)

    def forward(self, inputs):
        pass


class Memory1D(nn.Module):
    """models/contrastive.py:1042-1080: [length, dim] bank."""

    def __init__(self, length, duration, dim, cfg):
        super(Memory1D, self).__init__()
        assert duration == 1
        self.length = length
        self.duration = duration
        self.dim = dim
        stdv = 1.0 / math.sqrt(dim / 3)
        self.register_buffer("memory", torch.rand(length, dim).mul_(2 * stdv).add_(-stdv))
        self.l2_norm = Normalize(dim=1)
        self.num_gpus = cfg.NUM_GPUS

    @torch.no_grad()
    def get(self, ind, time, interp=False):
        batch_size = ind.size(0)
        if len(ind.shape) == 1:
            return torch.index_select(self.memory, 0, ind.view(-1)).view(batch_size, self.dim)
        else:
            return torch.index_select(self.memory, 0, ind.view(-1)).view(batch_size, -1, self.dim)

    @torch.no_grad()
    def update(self, mem, momentum, ind, time, interp=False, status=None):
        if self.num_gpus > 1:
            mem, ind, time = du.all_gather([mem, ind, time])
        mem = mem.view(mem.size(0), -1)
        ops.membank_update(self.memory, mem.detach().contiguous(), ind.long(), None, momentum, interp=False,
                           status=status)


def _exchange_rows(x, idx_randperm, gpu_idx, world_size):
    """C1 as an all-to-all: rank r needs rows idx_randperm[r] of the
    rank-major concatenation; row g lives on rank g // B at local offset g % B.
    Every rank knows the whole permutation, so all split sizes are computed locally.
    Bit-identical to `cat_all_gather(x)[idx_randperm[r]]`."""
    import torch.distributed as dist
    B = x.shape[0]
    perm = idx_randperm.cpu()  # [W, B]; 8 bytes per clip, needed on the host for the split sizes
    owner = perm // B
    send_rows, send_sizes = [], []
    me = gpu_idx
    for dst in range(world_size):
        sel = perm[dst][owner[dst] == me] % B  # my rows that dst wants, in dst's order
        send_rows.append(sel)
        send_sizes.append(int(sel.numel()))
    recv_sizes = [int((owner[me] == src).sum()) for src in range(world_size)]
    send_idx = torch.cat(send_rows).to(x.device)
    send_buf = x.index_select(0, send_idx).contiguous()
    recv_buf = torch.empty((B,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    dist.all_to_all_single(recv_buf, send_buf, output_split_sizes=recv_sizes, input_split_sizes=send_sizes)
    # recv_buf is grouped by source rank; put the rows into this rank's take order
    pos = torch.argsort(owner[me], stable=True)  # positions in take order, grouped by source
    inv = torch.empty_like(pos)
    inv[pos] = torch.arange(B)
    return recv_buf.index_select(0, inv.to(x.device))


def contrastive_parameter_surgery(model, cfg, epoch_exact, cur_iter):
    """models/contrastive.py:1083-1116."""
    if cfg.MODEL.MODEL_NAME == "ContrastiveModel" and cfg.CONTRASTIVE.TYPE == "swav" and epoch_exact <= 1.0:
        for name, p in model.named_parameters():
            if "swav_prototypes" in name:
                p.grad = None

    iters_noupdate = 0
    if cfg.MODEL.MODEL_NAME == "ContrastiveModel" and cfg.CONTRASTIVE.TYPE == "moco":
        assert cfg.CONTRASTIVE.QUEUE_LEN % (cfg.TRAIN.BATCH_SIZE * cfg.NUM_SHARDS) == 0
        iters_noupdate = cfg.CONTRASTIVE.QUEUE_LEN // cfg.TRAIN.BATCH_SIZE // cfg.NUM_SHARDS

    if cur_iter < iters_noupdate and epoch_exact < 1:
        logger.info("Not updating parameters {}/{}".format(cur_iter, iters_noupdate))
        update_param = False
    else:
        update_param = True
    return model, update_param


def contrastive_forward(model, cfg, inputs, index, time, epoch_exact, scaler=None):
    """models/contrastive.py:1119-1171."""
    if cfg.CONTRASTIVE.SEQUENTIAL:
        perform_backward = False
        mdl = getattr(model, "module", model)
        keys = (
            mdl.compute_key_feat(inputs, compute_predictor_keys=False,
                                 batched_inference=True if len(inputs) < 2 else False)
            if cfg.CONTRASTIVE.TYPE == "moco" or cfg.CONTRASTIVE.TYPE == "byol"
            else [None] * len(inputs)
        )
        for k, vid in enumerate(inputs):
            other_keys = keys[:k] + keys[k + 1:]
            time_cur = None if time is None else torch.cat(
                [time[:, k:k + 1, :], time[:, :k, :], time[:, k + 1:, :]], 1)  # q, kpre, kpost
            vids = [vid]
            if cfg.CONTRASTIVE.TYPE == "swav" or cfg.CONTRASTIVE.TYPE == "simclr":
                if k < len(inputs) - 1:
                    vids = inputs[k:k + 2]
                else:
                    break
            lgt_k, loss_k = model(vids, index, time_cur, epoch_exact, keys=other_keys)
            if scaler is not None:
                scaler.scale(loss_k).backward()
            else:
                loss_k.backward()
            if k == 0:
                preds, partial_loss = lgt_k, loss_k.detach()
            else:
                preds = torch.cat([preds, lgt_k], dim=0)
                partial_loss += loss_k.detach()
        partial_loss /= len(inputs) * 2.0  # to have same loss as symm model
        if cfg.CONTRASTIVE.TYPE == "moco":
            mdl._dequeue_and_enqueue(keys)
    else:
        perform_backward = True
        preds, partial_loss = model(inputs, index, time, epoch_exact, keys=None)
    return model, preds, partial_loss, perform_backward


try:
    MODEL_REGISTRY.register()(ContrastiveModel)
except Exception:  # already registered by the host application
    pass
