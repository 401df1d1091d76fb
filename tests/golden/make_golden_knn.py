"""Generates tests/golden/knn.npz from the UNMODIFIED reference: `ContrastiveModel.eval_knn`
(`/root/reference/models/contrastive.py:232-241`) and the eval branch of `forward` that calls it (:469-474), imported
through ref_shim.  The bank holds unit rows (what `knn_mem_update` stores, :131-140).
Run:  python tests/golden/make_golden_knn.py"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

rc = ref_shim.load_reference()
ref_shim.register_stub_backbone(rc)


def main():
    N, D, L = 12, 32, 1500
    cfg = ref_shim.make_cfg(CONTRASTIVE__TYPE="moco", CONTRASTIVE__T=0.1, CONTRASTIVE__DIM=D, CONTRASTIVE__QUEUE_LEN=64,
                            CONTRASTIVE__LENGTH=L, CONTRASTIVE__KNN_ON=True)
    torch.manual_seed(33)
    model = rc.ContrastiveModel(cfg).eval()
    g = torch.Generator().manual_seed(34)
    bank = torch.nn.functional.normalize(torch.randn(L, D, generator=g), dim=1)
    with torch.no_grad():
        model.knn_mem.memory.copy_(bank.view(L, 1, D))
    x = torch.randn(N, D, generator=g)
    q = torch.nn.functional.normalize(x, dim=1)
    out = {"N": N, "D": D, "L": L, "bank": bank.numpy(), "q": q.numpy(), "x": x.numpy(),
           "W": model.backbone.proj.weight.detach().numpy()}
    for k in (200, 5):
        yd, yi = model.eval_knn(q, knn_k=k)
        out["yd%d" % k], out["yi%d" % k] = yd.numpy(), yi.numpy()
    # the eval branch of forward (:469-474): backbone -> Normalize -> eval_knn, default knn_k = 200
    yd, yi = model([[x], [x]], torch.arange(N), torch.zeros(N, 2, 1))
    out["fwd_yd"], out["fwd_yi"] = yd.detach().numpy(), yi.numpy()
    np.savez_compressed(os.path.join(HERE, "knn.npz"), **out)
    print("wrote knn.npz:", out["yd200"].shape, out["yi200"].dtype, out["fwd_yd"].shape)


if __name__ == "__main__":
    main()
