"""Generates tests/golden/temporal.npz from the UNMODIFIED reference
(`/root/reference/models/temporal_modeling.py:217-238` `_update_history`, `:354-375`
`contrast_forward`), imported through ref_shim in the authoring container.

TemporalModel itself cannot be constructed here (it loads a CLIP backbone), but the two methods
only touch a handful of attributes, so they are called unbound on a plain holder object that
carries small stand-in modules with the same parameter naming: that executes the reference's own
code for the path.  Run:  python tests/golden/make_golden_temporal.py
"""
import os
import sys

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

ref_shim.install()
import models.temporal_modeling as tm  # noqa: E402  (reference)
from models.contrastive import Normalize  # noqa: E402  (reference)

B, E, D, T, M = 16, 48, 32, 0.2, 0.99


def make_modules(seed):
    g = torch.Generator().manual_seed(seed)

    def lin(i, o):
        l = nn.Linear(i, o)
        with torch.no_grad():
            l.weight.copy_(torch.randn(o, i, generator=g) * 0.2)
            l.bias.copy_(torch.randn(o, generator=g) * 0.1)
        return l
    enc = nn.Sequential(lin(E, E), nn.LayerNorm(E), lin(E, E))
    proj = lin(E, D)
    return enc, proj


class Holder(nn.Module):
    """Carries exactly the attributes the two reference methods read."""

    def __init__(self):
        super().__init__()
        self.temporal_encoder, self.head_projector = make_modules(1)
        self.temporal_encoder_hist, self.head_projector_hist = make_modules(2)  # deliberately different
        self.temporal_encoder_hist.eval().requires_grad_(False)
        self.head_projector_hist.eval().requires_grad_(False)
        g = torch.Generator().manual_seed(3)
        self.head_predictor = nn.Linear(D, D)
        with torch.no_grad():
            self.head_predictor.weight.copy_(torch.randn(D, D, generator=g) * 0.2)
            self.head_predictor.bias.copy_(torch.randn(D, generator=g) * 0.1)
        self.l2_norm = Normalize(dim=1)
        self.mmt, self.T = M, T


def params(mod_pairs):
    out = {}
    for tag, mod in mod_pairs:
        for n, p in mod.named_parameters():
            out["%s/%s" % (tag, n)] = p.detach().clone().numpy()
    return out


def main():
    torch.manual_seed(0)
    h = Holder()
    out = {"B": B, "E": E, "D": D, "T": T, "m": M}
    for k, v in params([("enc", h.temporal_encoder), ("proj", h.head_projector), ("pred", h.head_predictor)]).items():
        out["online0/" + k] = v
    for k, v in params([("enc", h.temporal_encoder_hist), ("proj", h.head_projector_hist)]).items():
        out["hist_init/" + k] = v
    g = torch.Generator().manual_seed(9)
    for step in range(3):
        if step > 0:  # the online weights move between steps (as an optimiser would)
            with torch.no_grad():
                for p in list(h.temporal_encoder.parameters()) + list(h.head_projector.parameters()):
                    p.add_(torch.randn(p.shape, generator=g) * 0.05)
            for k, v in params([("enc", h.temporal_encoder), ("proj", h.head_projector)]).items():
                out["online%d/%s" % (step, k)] = v
        tm.TemporalModel._update_history(h)           # reference code
        for k, v in params([("enc", h.temporal_encoder_hist), ("proj", h.head_projector_hist)]).items():
            out["hist%d/%s" % (step, k)] = v
    out["n_steps"] = 3

    feats = [torch.randn(B, E, generator=g).requires_grad_(True) for _ in range(2)]
    keys = [torch.randn(B, E, generator=g) for _ in range(2)]
    loss = tm.TemporalModel.contrast_forward(h, feats, keys)   # reference code
    loss.backward()
    out["loss"] = loss.detach().numpy()
    for i in range(2):
        out["feat%d" % i] = feats[i].detach().numpy()
        out["key%d" % i] = keys[i].numpy()
        out["dfeat%d" % i] = feats[i].grad.numpy()
    for tag, mod in (("proj", h.head_projector), ("pred", h.head_predictor)):
        for n, p in mod.named_parameters():
            out["grad/%s/%s" % (tag, n)] = p.grad.numpy()
    np.savez_compressed(os.path.join(HERE, "temporal.npz"), **out)
    print("wrote temporal.npz: loss %.6f, %d arrays" % (float(loss), len(out)))


if __name__ == "__main__":
    main()
