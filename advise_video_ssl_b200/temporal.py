"""The AdViSe `TemporalModel`'s own momentum / BYOL-cosine path on the B200 kernels
(SURVEY.md §8(f) rank 1): drop-in bodies for

    TemporalModel._update_history   models/temporal_modeling.py:217-238  -> K1 (one launch)
    TemporalModel.contrast_forward  models/temporal_modeling.py:354-375  -> K2 + K7

`TemporalModel` itself (CLIP spatial encoder, temporal transformer, heads) stays the reference's
code.  A maintainer binds the two methods:

    from advise_video_ssl_b200 import temporal
    TemporalModel._update_history = temporal.update_history
    TemporalModel.contrast_forward = temporal.contrast_forward

or mixes `TemporalContrastMixin` in front of `nn.Module`.  Same attribute names are read
(`temporal_encoder[_hist]`, `head_projector[_hist]`, `head_predictor`, `mmt`, `T`, `init_flag`);
there is no CPU fallback.
"""
import logging

import torch

from . import ops
from .autograd import ByolSimilarity

logger = logging.getLogger(__name__)


def _pairs(self):
    """(online, hist) parameter pairs in the order the reference walks them (:220-237)."""
    enc = dict(self.temporal_encoder.named_parameters())
    proj = dict(self.head_projector.named_parameters())
    online, hist = [], []
    for name, p in self.temporal_encoder_hist.named_parameters():
        online.append(enc[name].data)
        hist.append(p.data)
    for name, p in self.head_projector_hist.named_parameters():
        online.append(proj[name].data)
        hist.append(p.data)
    return online, hist


@torch.no_grad()
def update_history(self):
    """hist <- online*(1-m) + hist*m over temporal_encoder + head_projector in ONE multi-tensor
    launch (the reference issues ~3 kernels and an allocation per tensor); on the first call the
    history is first replaced by the online weights (`init_flag`, :226-232).  Bit-exact."""
    online, hist = _pairs(self)
    st = self.__dict__.get("_avssl_temporal_ema")
    if st is None or not st[0].matches(online, hist):
        plan = ops.EmaPlan(online, hist)
        st = (plan, torch.zeros(1, dtype=torch.int64, device=plan.device))
        self.__dict__["_avssl_temporal_ema"] = st
    first = not hasattr(self, "init_flag")
    if first:
        setattr(self, "init_flag", True)
        logger.info("EMA Models Initializing.")
    st[0].run(self.mmt, st[1], bump_iter=False, first_iter=first)


def contrast_forward(self, feats, keys):
    """Symmetric BYOL cosine loss (:354-375): the two l2-norms, the row dots, the mean and the
    gradient w.r.t. the predictor output come from K2 + K7 (one launch per pair for the loss and
    d loss / d q); the three heads stay torch modules."""
    assert len(feats) == len(keys) == 2  # HACK in the reference too: only 2 positive samples
    keys = keys[::-1]
    loss = 0.0
    for feat, key in zip(feats, keys):
        feat = self.head_projector(feat)
        q = self.head_predictor(feat)
        with torch.no_grad():
            k = self.head_projector_hist(key)
            k, _ = ops.l2norm_fwd(k.float().contiguous(), eps=0.0)       # Normalize(dim=1), no eps
        loss = loss + ByolSimilarity.apply(q.float(), k, self.T, True)  # -mean(l2(q).k)/T, q normalised inside
    return loss / len(feats) + 1.0 / self.T


class TemporalContrastMixin:
    """`class TemporalModel(TemporalContrastMixin, nn.Module)` picks both methods up."""
    _update_history = update_history
    contrast_forward = contrast_forward
