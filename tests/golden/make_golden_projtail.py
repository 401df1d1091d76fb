"""Generates tests/golden/projtail.npz from the UNMODIFIED reference: `MLPHead`
(`/root/reference/models/head_helper.py:20-68`, Linear -> BN1d -> ReLU -> Linear -> BN1d -> ReLU -> Linear) followed
by `Normalize` (`/root/reference/models/contrastive.py:923-934`), forward and backward, imported through ref_shim.
Two cases: the cfg2 projection (dim 128, with bias, BN on) at a ragged batch, and a dim-256 tail without bias.
Run:  python tests/golden/make_golden_projtail.py"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

ref_shim.install()
import models.head_helper as rh  # noqa: E402  (reference)
import models.contrastive as rc  # noqa: E402  (reference)

CASES = {  # name: (B, dim_in, mlp_dim, dim_out, num_layers, bn_on, bias)
    "a": (37, 48, 136, 128, 3, True, True),
    "b": (70, 40, 100, 256, 2, False, False),
}


def main():
    out = {}
    for name, (B, dim_in, mlp_dim, dim_out, layers, bn_on, bias) in CASES.items():
        torch.manual_seed(5 + len(name) + B)
        head = rh.MLPHead(dim_in, dim_out, mlp_dim, layers, bn_on=bn_on, bias=bias).train()
        norm = rc.Normalize(power=2, dim=1)
        h = torch.randn(B, dim_in, requires_grad=True)
        G = torch.randn(B, dim_out)
        taps = {}
        last = head.projection[len(head.projection) - 1]
        hook = last.register_forward_hook(lambda m, i, o: taps.update(x=i[0].detach().clone(), y=o.detach().clone()))
        q = norm(head(h))
        hook.remove()
        (q * G).sum().backward()
        out.update({name + "_cfg": np.array([B, dim_in, mlp_dim, dim_out, layers, int(bn_on), int(bias)], dtype=np.int64),
                    name + "_h": h.detach().numpy(), name + "_G": G.numpy(), name + "_q": q.detach().numpy(),
                    name + "_tail_x": taps["x"].numpy(), name + "_tail_y": taps["y"].numpy(),
                    name + "_dh": h.grad.numpy()})
        for k, v in head.state_dict().items():
            out["%s_sd_%s" % (name, k)] = v.detach().numpy()
        for k, p in head.named_parameters():
            out["%s_grad_%s" % (name, k)] = p.grad.numpy()
    np.savez_compressed(os.path.join(HERE, "projtail.npz"), **out)
    print("wrote projtail.npz:", {k: v.shape for k, v in out.items() if k.endswith("_q")})


if __name__ == "__main__":
    main()
