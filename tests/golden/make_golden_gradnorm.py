"""Generates tests/golden/gradnorm.npz from the UNMODIFIED reference `get_grad_norm_`
(`/root/reference/models/optimizer.py:375-397`), imported through ref_shim.
Run:  python tests/golden/make_golden_gradnorm.py"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

ref_shim.install()
import models.optimizer as ro  # noqa: E402  (reference)

SHAPES = [(4097,), (33, 65), (8192,), (1,), (5000,), (3, 4096), (7, 3, 3, 3, 3)]


def main():
    g = torch.Generator().manual_seed(21)
    params = []
    out = {"n": len(SHAPES)}
    for i, s in enumerate(SHAPES):
        p = torch.nn.Parameter(torch.zeros(s))
        p.grad = torch.randn(s, generator=g) * (0.1 + i)
        params.append(p)
        out["grad%d" % i] = p.grad.numpy()
    extra = torch.nn.Parameter(torch.zeros(5))  # grad None: skipped by the reference (:378)
    out["total"] = ro.get_grad_norm_(params + [extra]).numpy()
    out["per_tensor"] = np.array([float(torch.norm(p.grad)) for p in params], dtype=np.float32)
    out["empty"] = ro.get_grad_norm_([extra]).numpy()
    np.savez_compressed(os.path.join(HERE, "gradnorm.npz"), **out)
    print("wrote gradnorm.npz: total", float(out["total"]))


if __name__ == "__main__":
    main()
