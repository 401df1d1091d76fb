"""Developer probe (GPU box): launch the InfoNCE kernels a few times (for ncu)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from advise_video_ssl_b200 import ops
impl = int(sys.argv[1]) if len(sys.argv) > 1 else 2
want = (sys.argv[2] == "1") if len(sys.argv) > 2 else False
feat = torch.randn(64, 128).cuda()
key = torch.nn.functional.normalize(torch.randn(64, 128)).cuda()
queue = torch.nn.functional.normalize(torch.randn(65536, 128)).cuda()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(6):
    flush.zero_()
    ops.moco_infonce(feat, [key], queue, 0.1, want, impl)
torch.cuda.synchronize()
print("done")
