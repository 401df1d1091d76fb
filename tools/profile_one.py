"""Developer probe (GPU box): runs ONE kernel family a few times so that `ncu --set full -k regex:<name>` can capture it.
Usage: python tools/profile_one.py k11|knn|projtail|sinkhorn|byol"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from advise_video_ssl_b200 import ops  # noqa: E402

dev = torch.device("cuda")
what = sys.argv[1] if len(sys.argv) > 1 else "k11"
torch.manual_seed(0)
if what == "k11":
    Bc, crops, P = 256, 6, 3000
    scores = torch.randn(crops * Bc, P, device=dev) * 0.1
    codes = torch.softmax(torch.randn(2, Bc, P, device=dev), -1)
    for _ in range(3):
        ops.swav_ce(scores, codes, crops, Bc, 0.1)
elif what == "knn":
    N, M, D, k = 64, 239975, 128, 200
    q = torch.nn.functional.normalize(torch.randn(N, D, device=dev), dim=1)
    bank = torch.nn.functional.normalize(torch.randn(M, D, device=dev), dim=1)
    for _ in range(3):
        ops.knn_similarity_topk(q, bank, k)
elif what == "projtail":
    B, Kin, Dout = 64, 2048, 128
    x = torch.randn(B, Kin, device=dev).relu_()
    W = torch.randn(Dout, Kin, device=dev) / Kin ** 0.5
    b = torch.randn(Dout, device=dev)
    G = torch.randn(B, Dout, device=dev)
    for _ in range(3):
        q, nrm = ops.linear_l2norm_fwd(x, W, b)
        ops.linear_l2norm_bwd(x, W, q, nrm, G)
elif what == "sinkhorn":
    scores = torch.randn(256, 3000, device=dev) * 0.1
    for _ in range(3):
        ops.sinkhorn(scores, 0.05, 3)
elif what == "byol":
    pred, key = torch.randn(64, 256, device=dev), torch.nn.functional.normalize(torch.randn(64, 256, device=dev), dim=1)
    for _ in range(3):
        ops.byol_simloss(pred, key, 0.1)
torch.cuda.synchronize()
print("ok", what)
