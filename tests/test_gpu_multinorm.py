"""-m gpu: multi-tensor L2 norm (get_grad_norm_, models/optimizer.py:375-397)
against the golden produced by the reference function and against the oracle at full size."""
import pytest
import torch

from helpers import rel_err
from oracle import contrastive_oracle as O

pytestmark = pytest.mark.gpu

RTOL = 2e-6  # fp32 sums of squares in a different (fixed) order than ATen's; far inside north_star's 1e-3


def test_get_grad_norm_golden(golden):
    from advise_video_ssl_b200 import optimizer
    g = golden("gradnorm")
    n = int(g.scalar("n"))
    params = []
    for i in range(n):
        p = torch.nn.Parameter(torch.zeros_like(g["grad%d" % i]).cuda())
        p.grad = g["grad%d" % i].cuda()
        params.append(p)
    params.append(torch.nn.Parameter(torch.zeros(5).cuda()))  # no grad: skipped
    total = optimizer.get_grad_norm_(params)
    assert total.is_cuda and total.dim() == 0
    assert rel_err(total, g["total"]) < RTOL
    per = optimizer.per_parameter_norms([p.grad for p in params[:-1]])
    assert rel_err(per, g["per_tensor"]) < RTOL
    again = optimizer.get_grad_norm_(params)  # cached plan, deterministic
    assert torch.equal(total, again)
    assert optimizer.get_grad_norm_([params[-1]]).item() == 0.0
    # other p-norms take the stock path, like the reference
    assert rel_err(optimizer.get_grad_norm_(params, 1.0), O.grad_norm([p.grad.cpu() for p in params[:-1]], 1.0)) < 1e-5


def test_multi_norm_unaligned_ragged_and_full_size():
    from advise_video_ssl_b200 import ops
    torch.manual_seed(4)
    base = torch.randn(3 * 4096 + 77, device="cuda")
    views = [base[1:4098], base[4099:4100], base[4100:4100], base[4100:]]  # misaligned, 1-element, EMPTY, ragged tail
    plan = ops.MultiTensorNorm(views)
    total, per = plan.run()
    ref = [torch.norm(v.cpu().double()) for v in views]
    assert rel_err(per, torch.stack(ref)) < RTOL
    assert rel_err(total, torch.norm(torch.stack(ref))) < RTOL
    # BASELINE cfg2 size: the Slow-R50 parameter list (36.1 M elements); scaling property ||a x|| = |a| ||x||
    import json, os
    shapes = [tuple(s) for _, s in json.load(open(os.path.join(os.path.dirname(__file__), "golden", "slow_r50_param_shapes.json")))["slow_r50_moco_dim128"]["shapes"]]
    xs = [torch.randn(s, device="cuda") for s in shapes]
    plan = ops.MultiTensorNorm(xs)
    t1 = plan.run()[0].clone()
    ref = torch.sqrt(sum((x.double() ** 2).sum() for x in xs))
    assert rel_err(t1, ref) < RTOL
    for x in xs:
        x.mul_(-2.0)
    t2 = plan.run()[0]
    assert rel_err(t2, 2.0 * t1) < 1e-6
    assert plan.algorithmic_bytes == 4 * 36095168


def test_multi_norm_no_cpu_fallback():
    from advise_video_ssl_b200 import ops
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.MultiTensorNorm([torch.randn(4)])
