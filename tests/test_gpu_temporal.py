"""-m gpu: the TemporalModel momentum / BYOL-cosine path (SURVEY.md §8(f) rank 1,
models/temporal_modeling.py:217-238, 354-375) on K1 / K2 / K7 against golden vectors produced by
the reference's own two methods (tests/golden/make_golden_temporal.py)."""
import pytest
import torch
import torch.nn as nn

from helpers import rel_err

pytestmark = pytest.mark.gpu


def _names(g):
    return [k[len("hist_init/"):] for k in g.keys() if k.startswith("hist_init/")]


class Holder(nn.Module):
    """Same attribute layout as the reference's TemporalModel for this path."""

    def __init__(self, g, mixin):
        super().__init__()
        E, D = int(g.scalar("E")), int(g.scalar("D"))

        def enc():
            return nn.Sequential(nn.Linear(E, E), nn.LayerNorm(E), nn.Linear(E, E))
        self.temporal_encoder, self.head_projector = enc(), nn.Linear(E, D)
        self.temporal_encoder_hist = enc().eval().requires_grad_(False)
        self.head_projector_hist = nn.Linear(E, D).eval().requires_grad_(False)
        self.head_predictor = nn.Linear(D, D)
        self.mmt, self.T = g.scalar("m"), g.scalar("T")


def _load(mod, g, prefix, tag):
    with torch.no_grad():
        for n, p in mod.named_parameters():
            p.copy_(g["%s/%s/%s" % (prefix, tag, n)])


def test_temporal_update_history_bit_exact_and_contrast_loss(golden):
    from advise_video_ssl_b200 import temporal
    g = golden("temporal")

    class Model(temporal.TemporalContrastMixin, Holder):
        pass

    m = Model(g, None)
    _load(m.temporal_encoder, g, "online0", "enc")
    _load(m.head_projector, g, "online0", "proj")
    _load(m.head_predictor, g, "online0", "pred")
    _load(m.temporal_encoder_hist, g, "hist_init", "enc")
    _load(m.head_projector_hist, g, "hist_init", "proj")
    m = m.cuda()
    n_steps = int(g.scalar("n_steps"))
    for s in range(n_steps):
        if s > 0:
            _load(m.temporal_encoder, g, "online%d" % s, "enc")
            _load(m.head_projector, g, "online%d" % s, "proj")
        m._update_history()
        assert hasattr(m, "init_flag")
        for tag, mod in (("enc", m.temporal_encoder_hist), ("proj", m.head_projector_hist)):
            for n, p in mod.named_parameters():
                assert torch.equal(p.detach().cpu(), g["hist%d/%s/%s" % (s, tag, n)]), (s, tag, n)

    feats = [g["feat%d" % i].cuda().requires_grad_(True) for i in range(2)]
    keys = [g["key%d" % i].cuda() for i in range(2)]
    loss = m.contrast_forward(feats, keys)
    loss.backward()
    # north_star tolerance for fp32 loss / gradients is 1e-3 relative; the kernels are well inside
    assert rel_err(loss, g["loss"]) < 1e-5
    for i in range(2):
        assert rel_err(feats[i].grad, g["dfeat%d" % i]) < 1e-4
    for tag, mod in (("proj", m.head_projector), ("pred", m.head_predictor)):
        for n, p in mod.named_parameters():
            assert rel_err(p.grad, g["grad/%s/%s" % (tag, n)]) < 1e-4, (tag, n)
    for p in list(m.temporal_encoder_hist.parameters()) + list(m.head_projector_hist.parameters()):
        assert p.grad is None


def test_temporal_functions_bind_onto_a_foreign_class(golden):
    """The INTEGRATION.md recipe: assign the two functions onto the reference's class."""
    from advise_video_ssl_b200 import temporal
    g = golden("temporal")

    class Foreign(Holder):
        pass

    Foreign._update_history = temporal.update_history
    Foreign.contrast_forward = temporal.contrast_forward
    m = Foreign(g, None).cuda()
    before = [p.detach().clone() for p in m.temporal_encoder.parameters()]
    m._update_history()  # first call: history := online, then blended with itself
    for b, h in zip(before, m.temporal_encoder_hist.parameters()):
        expect = b * (1.0 - m.mmt) + b * m.mmt
        assert torch.equal(h, expect)
    with pytest.raises(AssertionError):
        m.contrast_forward([torch.randn(4, int(g.scalar("E"))).cuda()], [torch.randn(4, int(g.scalar("E"))).cuda()])
