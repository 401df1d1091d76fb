"""Projection tail: the step immediately before the contrastive head.

The reference's backbones end in a projection MLP (`models/head_helper.py:20-68`, `MLPHead`: Linear -> BN1d -> ReLU
-> ... -> Linear) whose raw output the contrastive model normalises first thing (`models/contrastive.py:462`, `:350`,
`:757`, `:850`).  `LinearNormalize` is a drop-in for that LAST `nn.Linear` — same `weight` / `bias` parameters, so
checkpoints, the optimiser's parameter list and the momentum encoder's `named_parameters()` order are untouched —
that emits unit rows from the GEMM's epilogue: the raw projection never reaches HBM, and the backward launch applies
the gradient of the normalisation where it reads the upstream gradient.

    fuse_projection_tail(model)        # ContrastiveModel, a backbone, a head or an MLPHead / nn.Sequential

The contrastive head normalises its input again; on unit rows that is the identity up to one rounding (||q|| = 1 ± ulp)
and its gradient is the same tangent-space projection applied twice, so losses and gradients agree with the unfused
path to fp32 rounding (`tests/test_gpu_projtail.py`).

Not fused (the tail is left as it is): heads with predictors (BYOL: the predictors consume the RAW projection,
`models/head_helper.py:216-220`), layers wider than 256 outputs or with `in_features % 4 != 0`.
"""
import torch
import torch.nn as nn

from . import ops
from .autograd import LinearNormalize as _LinearNormalizeFn


class LinearNormalize(nn.Module):
    """`nn.Linear` + `Normalize(power=2, dim=1)` in one kernel.  Parameters are named and shaped like nn.Linear's."""

    def __init__(self, in_features, out_features, bias=True, eps=0.0, normalize=True):
        super().__init__()
        if not ops.linear_l2norm_supported(in_features, out_features):
            raise ValueError("LinearNormalize: unsupported shape (in %d, out %d): needs out <= 256 and in %% 4 == 0"
                             % (in_features, out_features))
        self.in_features, self.out_features = in_features, out_features
        self.eps, self.normalize = float(eps), bool(normalize)
        ref = nn.Linear(in_features, out_features, bias=bias)  # same default initialisation as the layer it replaces
        self.weight = ref.weight
        if bias:
            self.bias = ref.bias
        else:
            self.register_parameter("bias", None)

    @classmethod
    def from_linear(cls, linear, eps=0.0, normalize=True):
        """Wraps an existing nn.Linear: the SAME Parameter objects (optimiser state and EMA tables stay valid)."""
        m = cls.__new__(cls)
        nn.Module.__init__(m)
        if not ops.linear_l2norm_supported(linear.in_features, linear.out_features):
            raise ValueError("LinearNormalize: unsupported shape (in %d, out %d)" % (linear.in_features, linear.out_features))
        m.in_features, m.out_features = linear.in_features, linear.out_features
        m.eps, m.normalize = float(eps), bool(normalize)
        m.weight = linear.weight
        if linear.bias is not None:
            m.bias = linear.bias
        else:
            m.register_parameter("bias", None)
        for attr in ("xavier_init",):  # the init hook the reference sets on its Linear layers (head_helper.py:39,58)
            if hasattr(linear, attr):
                setattr(m, attr, getattr(linear, attr))
        return m

    def forward(self, x):
        assert x.shape[-1] == self.in_features, "LinearNormalize expects [..., %d]" % self.in_features
        if x.is_cuda and (x.dim() != 2 or x.dtype != torch.float32):
            # fully-convolutional inference (a [N, T, H, W, C] input that the head averages AFTER the projection,
            # models/head_helper.py:222-228) and autocast inputs keep the plain Linear: the contrastive model
            # normalises whatever the backbone returns, so an un-normalised output is always valid
            return torch.nn.functional.linear(x, self.weight.to(x.dtype), None if self.bias is None else self.bias.to(x.dtype))
        return _LinearNormalizeFn.apply(x, self.weight, self.bias, self.eps, self.normalize)

    def extra_repr(self):
        return "in_features=%d, out_features=%d, bias=%s, normalize=%s" % (
            self.in_features, self.out_features, self.bias is not None, self.normalize)


def _tail_of(module):
    """(parent, key) of the last nn.Linear of a projection: `module` is an nn.Linear slot owner, an MLPHead-like
    module with a `.projection` nn.Sequential, or the Sequential itself.  None when there is nothing to fuse."""
    proj = getattr(module, "projection", None)
    if isinstance(proj, (nn.Linear, LinearNormalize)):          # SSL.NUM_MLP_LAYERS == 1 (head_helper.py:135-136)
        return module, "projection"
    if proj is not None and not isinstance(proj, nn.Sequential):  # ResNetBasicHead.projection = MLPHead
        return _tail_of(proj)
    seq = proj if isinstance(proj, nn.Sequential) else (module if isinstance(module, nn.Sequential) else None)
    if seq is not None and len(seq) and isinstance(seq[len(seq) - 1], (nn.Linear, LinearNormalize)):
        return seq, str(len(seq) - 1)
    return None


def _heads(root):
    """Modules that own a projection tail under `root`: `root` itself, or every submodule named `head`."""
    if _tail_of(root) is not None:
        return [root]
    return [m for name, m in root.named_modules() if name.split(".")[-1] == "head" and _tail_of(m) is not None]


def fuse_projection_tail(root, eps=0.0):
    """Replaces the last Linear of every projection MLP under `root` by `LinearNormalize` (in place, sharing the
    Parameters).  `root`: a ContrastiveModel (both encoders are converted), a backbone, a head, an MLPHead or an
    nn.Sequential.  Returns the number of layers converted; heads with predictors and unsupported shapes are skipped."""
    done = 0
    for head in _heads(root):
        if len(getattr(head, "predictors", ())) > 0:
            continue
        parent, key = _tail_of(head)
        lin = parent._modules[key]
        if isinstance(lin, LinearNormalize) or not ops.linear_l2norm_supported(lin.in_features, lin.out_features):
            continue
        if lin.weight.dtype != torch.float32:
            continue
        parent._modules[key] = LinearNormalize.from_linear(lin, eps=eps)
        done += 1
    return done
