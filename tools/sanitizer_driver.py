"""Compact driver for compute-sanitizer (GPU box): one small invocation of every kernel family that
synchronises through shared memory, mbarriers, cluster barriers or global flags -- K3 (tcgen05 InfoNCE with
the fused enqueue, indexed / raw keys and the fused key push), K10 (Sinkhorn: cluster and cooperative paths),
K6 (NT-Xent on tcgen05), K1 (EMA) and the peer exchange / scatter kernels on a one-rank exchange.

    compute-sanitizer --tool memcheck  python tools/sanitizer_driver.py
    compute-sanitizer --tool racecheck python tools/sanitizer_driver.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from advise_video_ssl_b200 import ops, _lib  # noqa: E402

dev = torch.device("cuda")
g = torch.Generator().manual_seed(0)


def rnd(*shape):
    return torch.randn(*shape, generator=g).to(dev)


# K1
online = [rnd(257, 33), rnd(4096 * 2 + 4), rnd(128)]
hist = [torch.zeros_like(o) for o in online]
it = torch.zeros(1, dtype=torch.int64, device=dev)
ops.EmaPlan(online, hist).run(0.99, it, bump_iter=True)

# K3 (+K4, indexed raw keys, fused push)
B, D, K, T = 64, 128, 2048, 0.1
queue = torch.nn.functional.normalize(rnd(K, D))
ptr = torch.zeros(1, dtype=torch.int64, device=dev)
status = torch.zeros(1, dtype=torch.int32, device=dev)
feat, key = rnd(B, D), torch.nn.functional.normalize(rnd(B, D))
ops.moco_infonce(feat, [key], queue, T, True, _lib.IMPL_TC3X, enqueue=(ptr, status))
ops.moco_infonce(feat, [key], queue, T, False, _lib.IMPL_TC1X)
ops.moco_infonce(feat, [key], queue, T, True, _lib.IMPL_SIMT, enqueue=(ptr, status))
perm = torch.randperm(B, generator=g).to(dev)
ops.moco_infonce(feat, None, queue, T, True, _lib.IMPL_AUTO, enqueue=(ptr, status), key_rows=rnd(B, D), keys_raw=True,
                 peer_row_idx=perm, enq_row_idx=perm)
x = ops.PeerExchange(B, D)
ops.moco_infonce(feat, None, queue, T, True, _lib.IMPL_AUTO, enqueue=(ptr, status), peer=x, push_rows=rnd(B, D),
                 peer_row_idx=perm, enq_row_idx=perm)
x.push_normalized(rnd(B, D), 0.0)
x.wait_gather(perm, status=status)
x.push(key)
x.wait_gather_all()
sc = ops.PeerScatter(B, 3 * 8 * 16 * 4)
sc.exchange(rnd(B, 3, 8, 16), None, status=status, dest_pos_host=torch.argsort(perm.cpu()).numpy())
sc.exchange(rnd(B, 3, 8, 16), torch.argsort(perm), status=status)

# K6
f1, f2 = rnd(96, 128), rnd(96, 128)
ops.ntxent(f1, f2, 0.1, gather=False, impl=_lib.IMPL_AUTO)
ops.ntxent(rnd(40, 256), rnd(40, 256), 0.2, gather=False, impl=_lib.IMPL_AUTO)
ops.ntxent(f1, f2, 0.1, gather=False, impl=_lib.IMPL_SIMT)

# K10 / K11
scores = rnd(64, 300) * 0.1
ops.sinkhorn(scores, 0.05, 3)                      # cluster path (fits distributed shared memory)
big = rnd(2048, 3000) * 0.1
ops.sinkhorn(big, 0.05, 3, keep_last=64)           # cooperative-grid path
codes = torch.softmax(rnd(2, 32, 300), -1)
ops.swav_ce(rnd(4 * 32, 300) * 0.1, codes, 4, 32, 0.1)

# K5 / K7 / multi-norm
bank = torch.nn.functional.normalize(rnd(500, 1, 64), dim=-1)
ops.membank_update(bank, rnd(24, 64), torch.randint(0, 500, (24,), generator=g).to(dev), None, 0.5, status=status)
ops.byol_simloss(rnd(64, 256), torch.nn.functional.normalize(rnd(64, 256)), 0.1)
ops.MultiTensorNorm(online).run()
torch.cuda.synchronize()
x.close()
sc.close()
assert int(status.item()) == 0
print("sanitizer driver done")
