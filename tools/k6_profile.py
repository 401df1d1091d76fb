import os, sys
sys.path.insert(0, os.getcwd())
import torch
from advise_video_ssl_b200._lib import lib, check
dev = torch.device("cuda")
B, W, D, T = 512, 8, 256, 0.1
N = B * W
gathered = torch.nn.functional.normalize(torch.randn(W, 2, B, D, device=dev), dim=-1)
out = torch.empty(2 * N, D, device=dev); out_r = torch.empty(2 * N, D, device=dev, dtype=torch.float16)
rows = torch.cat([torch.arange(0, B, dtype=torch.int32, device=dev), torch.arange(N, N + B, dtype=torch.int32, device=dev)])
n_loc = 2 * B
ws = torch.zeros(lib.avssl_ntxent_workspace_bytes(2 * N, D, n_loc), dtype=torch.uint8, device=dev)
z = torch.empty(n_loc, device=dev); zall = torch.ones(2 * N, device=dev); nrm = torch.ones(n_loc, device=dev)
loss = torch.empty(1, device=dev); dfe = torch.empty(n_loc, D, device=dev)
st = torch.cuda.current_stream().cuda_stream
for _ in range(4):
    check(lib.avssl_ntxent_prepare(gathered.data_ptr(), W, B, D, out.data_ptr(), out_r.data_ptr(), st), "prepare")
    check(lib.avssl_ntxent_rowsum(out.data_ptr(), out_r.data_ptr(), rows.data_ptr(), 0, N, 2 * N, D, n_loc, T, z.data_ptr(), ws.data_ptr(), ws.numel(), 0, st), "rowsum")
    check(lib.avssl_ntxent_grad(out.data_ptr(), out_r.data_ptr(), rows.data_ptr(), 0, N, zall.data_ptr(), nrm.data_ptr(), 2 * N, D, n_loc, T, float(W), loss.data_ptr(), dfe.data_ptr(), ws.data_ptr(), ws.numel(), 0, st), "grad")
torch.cuda.synchronize()
print("done")
