"""Developer probe (GPU box): run the tcgen05 InfoNCE kernels on a few shapes and print
their errors against the CPU oracle, plus a quick timing.  Lives under tests/ because it imports the
oracle (test infrastructure); it is not collected by pytest."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from advise_video_ssl_b200 import ops  # noqa: E402
from oracle import contrastive_oracle as O  # noqa: E402


def run(B, D, K, T, impl, peaked=False, nk=1):
    torch.manual_seed(1)
    feat = torch.randn(B, D)
    queue = O.l2_normalize(torch.randn(K, D))
    q = O.l2_normalize(feat)
    if peaked:
        idx = torch.randint(0, K, (B, 3))
        for i in range(B):
            queue[idx[i]] = O.l2_normalize(q[i:i + 1] + 0.05 * torch.randn(3, D))
    keys = [O.l2_normalize(q + 0.3 * torch.randn(B, D)) for _ in range(nk)]
    cl, cdf, _ = O.moco_head_closed_form(feat, keys, queue, T)
    out = ops.moco_infonce(feat.cuda(), [k.cuda() for k in keys], queue.cuda(), T, True, impl)
    torch.cuda.synchronize()
    lg = O.moco_logits(q, keys, queue, T)
    le = abs(out["loss"].item() - cl.item()) / abs(cl.item())
    ge = (out["dfeat"].double().cpu() - cdf).abs().max().item() / cdf.abs().max().item()
    lge = (out["logits"].cpu() - lg).abs().max().item()
    print("impl %d B=%d D=%d K=%d T=%.2f peaked=%d: loss %.6f ref %.6f rel %.2e | grad rel %.2e | logits abs %.2e"
          % (impl, B, D, K, T, peaked, out["loss"].item(), cl.item(), le, ge, lge), flush=True)


if __name__ == "__main__":
    impls = [int(a) for a in sys.argv[1:]] or [3, 2]
    for impl in impls:
        run(64, 128, 64, 0.1, impl)
        run(64, 128, 1024, 0.1, impl)
        run(5, 32, 130, 0.07, impl, nk=2)
        run(130, 64, 777, 0.07, impl)
        run(64, 128, 65536, 0.1, impl)
        run(64, 128, 8192, 0.07, impl, peaked=True)
        # timing
        feat = torch.randn(64, 128).cuda()
        key = torch.nn.functional.normalize(torch.randn(64, 128)).cuda()
        queue = torch.nn.functional.normalize(torch.randn(65536, 128)).cuda()
        for want in (True, False):
            for _ in range(5):
                ops.moco_infonce(feat, [key], queue, 0.1, want, impl)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(50):
                ops.moco_infonce(feat, [key], queue, 0.1, want, impl)
            e1.record()
            torch.cuda.synchronize()
            print("impl %d logits=%s: %.2f us per call (queue L2-resident)" % (impl, want, e0.elapsed_time(e1) * 20), flush=True)
