"""Developer probe (GPU box): element-wise relative gradient error of the tensor-core and CUDA-core kernels on entries
above a floor (a fraction of the largest |reference| entry), to back the element-wise bounds asserted in the tests.
Lives under tests/ because it imports the oracle; not collected by pytest.  Prints one JSON line per case."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402
from advise_video_ssl_b200 import _lib, ops  # noqa: E402
from oracle import contrastive_oracle as O  # noqa: E402
import recipes  # noqa: E402


def elementwise(a, ref):
    a, ref = a.double().cpu(), ref.double().cpu()
    mx = ref.abs().max().item()
    out = {"max_norm": (a - ref).abs().max().item() / mx}
    for frac in (0.01, 0.05, 0.1):
        m = ref.abs() >= frac * mx
        out["floor_%g" % frac] = ((a - ref).abs()[m] / ref.abs()[m]).max().item()
        out["kept_%g" % frac] = round(m.float().mean().item(), 3)
    return out


def golden(name):
    z = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    return {k: torch.from_numpy(np.array(z[k])) for k in z.files if z[k].dtype.kind not in "US"}


g = golden("moco_cfg1")
r = recipes.moco_cfg1()
hist = O.ema_update([r["W"]], [r["W"].clone()], r["m"], 0)[0]
keys = [O.l2_normalize(F.linear(r["xk"], hist))]
for name, impl in (("simt", _lib.IMPL_SIMT), ("tc3x", _lib.IMPL_TC3X), ("tc1x", _lib.IMPL_TC1X)):
    out = ops.moco_infonce(g["featq"].cuda(), [k.cuda() for k in keys], r["queue"].cuda(), r["T"], False, impl)
    print(json.dumps({"case": "moco cfg1 (B=64, K=65536, D=128) dfeat vs the reference golden", "impl": name,
                      **elementwise(out["dfeat"], g["dfeatq"])}), flush=True)

for B, D, T in ((512, 256, 0.1), (100, 128, 0.1), (256, 64, 0.2)):
    gen = torch.Generator().manual_seed(B + D)
    f1, f2 = torch.randn(B, D, generator=gen), torch.randn(B, D, generator=gen)
    a, b = f1.double().requires_grad_(True), f2.double().requires_grad_(True)
    loss = O.ntxent(O.l2_normalize(a), O.l2_normalize(b), T)
    loss.backward()
    for name, impl in (("simt", _lib.IMPL_SIMT), ("tc", _lib.IMPL_AUTO)):
        l, d1, d2 = ops.ntxent(f1.cuda(), f2.cuda(), T, impl=impl)
        print(json.dumps({"case": "simclr B=%d D=%d dfeat1 vs fp64 autograd of the oracle" % (B, D), "impl": name,
                          "loss_rel": abs(l.item() - loss.item()) / abs(loss.item()), **elementwise(d1, a.grad)}), flush=True)
