"""Step-path helpers next to the contrastive head, same names and
semantics as the reference's `models/optimizer.py`:

    get_grad_norm_(parameters, norm_type=2.0)   models/optimizer.py:375-397

The reference launches one `torch.norm` per parameter, stacks the results, takes the norm of the
stack and — in `utils/solver.py:109-111` — immediately copies it to the host.  Here the 2-norm is ONE
multi-tensor launch over the gradient storages (ops.MultiTensorNorm, same chunk-table machinery as
the momentum update) and the result stays on the device until the caller asks for it.
"""
import torch

from . import ops

_plans = {}        # storage-pointer tuple -> ops.MultiTensorNorm (insertion-ordered: oldest first)
_MAX_PLANS = 8     # gradients, parameters (LARS) and a few sub-lists; storages are stable across steps


def _plan_for(tensors):
    key = tuple((t.data_ptr(), t.numel()) for t in tensors)  # a recycled address with another size is a new list
    plan = _plans.pop(key, None)
    if plan is None:
        plan = ops.MultiTensorNorm(tensors)
        while len(_plans) >= _MAX_PLANS:
            _plans.pop(next(iter(_plans)))  # drop the least recently used plan
    _plans[key] = plan                      # (re)insert as most recently used
    return plan


def get_grad_norm_(parameters, norm_type=2.0):
    """Total gradient norm, `torch.norm(torch.stack([torch.norm(p.grad, t) for p in parameters]), t)`.
    Returns a 0-dim device tensor (the reference's callers do `.cpu()` themselves)."""
    if isinstance(parameters, torch.Tensor):
        parameters = [parameters]
    parameters = [p for p in parameters if p.grad is not None]
    norm_type = float(norm_type)
    if len(parameters) == 0:
        return torch.tensor(0.0)
    grads = [p.grad.detach() for p in parameters]
    if norm_type != 2.0 or any(g.dtype != torch.float32 or not g.is_contiguous() for g in grads):
        # other p-norms are not on the step path (the reference only ever calls it with 2.0)
        device = grads[0].device
        return torch.norm(torch.stack([torch.norm(g, norm_type).to(device) for g in grads]), norm_type)
    total, _ = _plan_for(grads).run()
    return total.reshape(()).clone()


def per_parameter_norms(tensors):
    """||x_t||_2 for every tensor in one launch (LARS.step's param_norm / grad_norm,
    models/optimizer.py:351-352).  Returns a device tensor [len(tensors)]."""
    tensors = [t.detach() for t in tensors]
    _, per = _plan_for(tensors).run()
    return per.clone()
