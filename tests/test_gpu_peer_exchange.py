"""-m gpu tests of the NVLink peer-memory key exchange (C3) on one GPU: with world_size 1 the
exchange targets its own buffer, which exercises the push / epoch / slot / wait protocol, the
EMA-fused push and the InfoNCE-fused wait bit for bit.  The two-GPU check against NCCL's
all_gather is tools/peer_exchange_check.py (torchrun) and runs inside bench.py at N > 1."""
import pytest
import torch
import torch.nn.functional as F

from oracle import contrastive_oracle as O

pytestmark = pytest.mark.gpu


def _ops():
    from advise_video_ssl_b200 import ops
    return ops


def test_push_wait_gather_bit_exact_over_epochs():
    ops = _ops()
    B, D = 64, 128
    x = ops.PeerExchange(B, D)
    try:
        g = torch.Generator().manual_seed(3)
        for step in range(5):  # both payload slots, monotonic flags
            rows = torch.randn(B, D, generator=g).cuda()
            x.push(rows)
            own = x.wait_gather()
            assert torch.equal(own, rows), step
            perm = torch.randperm(B, generator=g).cuda()
            sel = x.wait_gather(perm)  # idx_restore-style row select (models/contrastive.py:226-229)
            assert torch.equal(sel, rows[perm]), step
            assert torch.equal(x.wait_gather_all(), rows), step
    finally:
        x.close()


def test_wait_gather_flags_bad_index():
    ops = _ops()
    x = ops.PeerExchange(8, 32)
    try:
        rows = torch.randn(8, 32).cuda()
        x.push(rows)
        status = torch.zeros(1, dtype=torch.int32, device="cuda")
        idx = torch.tensor([0, 8, 3, -1], dtype=torch.int64, device="cuda")
        out = torch.zeros(4, 32, device="cuda")
        x.wait_gather(idx, out=out, status=status)
        assert int(status.item()) & 2
        assert torch.equal(out[0], rows[0]) and torch.equal(out[2], rows[3])
        assert out[1].abs().sum().item() == 0 and out[3].abs().sum().item() == 0
    finally:
        x.close()


def test_rejects_wrong_block_shape():
    ops = _ops()
    x = ops.PeerExchange(8, 32)
    try:
        with pytest.raises(ValueError):
            x.push(torch.randn(4, 32).cuda())
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            x.push(torch.randn(8, 32))
    finally:
        x.close()


@pytest.mark.parametrize("first", [True, False, None])
def test_ema_fused_push_matches_plain_ema(first):
    ops = _ops()
    torch.manual_seed(5)
    shapes = [(257, 33), (4096 * 3 + 1,), (128,), (8192,)]
    online = [torch.randn(s).cuda() for s in shapes]
    hist_a = [torch.randn(s).cuda() for s in shapes]
    hist_b = [h.clone() for h in hist_a]
    it0 = 0 if first in (True, None) else 3
    it_a = torch.full((1,), it0, dtype=torch.int64, device="cuda")
    it_b = it_a.clone()
    B, D = 64, 128
    x = ops.PeerExchange(B, D)
    try:
        rows = torch.randn(B, D).cuda()
        ops.EmaPlan(online, hist_a).run(0.99, it_a, bump_iter=True, first_iter=first)
        ops.EmaPlan(online, hist_b).run(0.99, it_b, bump_iter=True, first_iter=first, push=(x, rows))
        for a, b in zip(hist_a, hist_b):
            assert torch.equal(a, b)
        assert int(it_a.item()) == int(it_b.item()) == it0 + 1
        assert torch.equal(x.wait_gather(), rows)
        ref = O.ema_update([o.cpu() for o in online], [torch.zeros_like(o).cpu() for o in online], 0.99, 0)
        assert len(ref) == len(online)
    finally:
        x.close()


@pytest.mark.parametrize("B,D,K", [(64, 128, 4096), (32, 64, 1024), (100, 128, 2000)])
def test_infonce_fused_wait_matches_plain_launch(B, D, K):
    ops = _ops()
    torch.manual_seed(11)
    f = torch.randn(B, D).cuda()
    k = F.normalize(torch.randn(B, D)).cuda()
    queue_a = F.normalize(torch.randn(K, D)).cuda()
    queue_b = queue_a.clone()
    queue0 = queue_a.cpu().clone()
    enq = K % B == 0
    ptr_a = torch.tensor([K - B if enq else 0], dtype=torch.int64, device="cuda")
    ptr_b = ptr_a.clone()
    st = torch.zeros(1, dtype=torch.int32, device="cuda")
    x = ops.PeerExchange(B, D)
    try:
        a = ops.moco_infonce(f, [k], queue_a, 0.1, enqueue=(ptr_a, st) if enq else None)
        x.push(k)
        b = ops.moco_infonce(f, None, queue_b, 0.1, enqueue=(ptr_b, st) if enq else None, peer=x)
        for name in ("loss", "dfeat", "q", "lse", "logits"):
            assert torch.equal(a[name], b[name]), name
        assert torch.equal(queue_a, queue_b) and torch.equal(ptr_a, ptr_b) and int(st.item()) == 0
        if enq:
            assert int(ptr_b.item()) == 0 and torch.equal(queue_b[K - B:], k)
        # row select inside the launch: key row i = gathered[perm[i]]
        perm = torch.randperm(B).cuda()
        x.push(k)
        c = ops.moco_infonce(f, None, queue0.cuda(), 0.1, peer=x, peer_row_idx=perm)
        d = ops.moco_infonce(f, [k[perm].contiguous()], queue0.cuda(), 0.1)
        for name in ("loss", "dfeat", "lse", "logits"):
            assert torch.equal(c[name], d[name]), name
        # against the oracle (north_star tolerance 1e-3; the kernel is far inside)
        fc = f.cpu().clone().requires_grad_(True)
        _, _, loss = O.moco_head(fc, [k.cpu()], queue0, 0.1)
        assert abs(b["loss"].item() - loss.item()) < 1e-3 * abs(loss.item())
    finally:
        x.close()


def test_fused_wait_needs_tensor_core_kernel():
    ops = _ops()
    from advise_video_ssl_b200 import _lib
    x = ops.PeerExchange(16, 48)  # D = 48 has no tcgen05 variant
    try:
        x.push(torch.randn(16, 48).cuda())
        with pytest.raises(_lib.AvsslError, match="avssl_peer_wait_gather"):
            ops.moco_infonce(torch.randn(16, 48).cuda(), None, torch.randn(256, 48).cuda(), 0.1, peer=x)
        k = x.wait_gather()
        ops.moco_infonce(torch.randn(16, 48).cuda(), [k], torch.randn(256, 48).cuda(), 0.1)
    finally:
        x.close()


@pytest.mark.parametrize("shape,dtype", [((64, 3, 4, 8), torch.float32), ((33, 40), torch.float16), ((5, 2), torch.int64),
                                         ((16, 3, 8, 112, 112), torch.float32)])
def test_scatter_rows_bit_exact_over_epochs(shape, dtype):
    """C1 as a scatter (models/contrastive.py:186-207): with one rank the rows land at argsort(perm) of its own
    buffer, i.e. the result is x[perm]; several epochs exercise both slots, and a CUDA graph replays it."""
    ops = _ops()
    B = shape[0]
    row_bytes = int(torch.empty(shape[1:], dtype=dtype).numel() * torch.empty(0, dtype=dtype).element_size())
    sc = ops.PeerScatter(B, row_bytes)
    status = torch.zeros(1, dtype=torch.int32, device="cuda")
    try:
        g = torch.Generator().manual_seed(11)
        for step in range(4):
            x = (torch.randn(shape, generator=g) * 100).to(dtype).cuda()
            perm = torch.randperm(B, generator=g)
            out = sc.exchange(x, torch.argsort(perm).cuda(), status=status)
            assert torch.equal(out, x[perm.cuda()]), step
        # graph replay: fixed destination although the slot alternates
        x = (torch.randn(shape, generator=g) * 100).to(dtype).cuda()
        dest = torch.argsort(perm).cuda()
        out = torch.empty_like(x)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            sc.exchange(x, dest, out=out, status=status)
        for rep in range(3):
            x.copy_((torch.randn(shape, generator=g) * 100).to(dtype))
            graph.replay()
            assert torch.equal(out, x[perm.cuda()]), rep
        bad = dest.clone()
        bad[0] = B
        sc.exchange(x, bad, out=out, status=status)
        assert int(status.item()) == 2
    finally:
        sc.close()
    with pytest.raises(ValueError):
        ops.PeerScatter(4, 24)  # rows must be 16-byte multiples


@pytest.mark.parametrize("B,D,K", [(64, 128, 4096), (32, 64, 1024)])
def test_head_launch_pushes_its_own_keys(B, D, K):
    """C3 push fused into the head launch: an extra CTA normalises this rank's raw key rows and stores them into the
    exchange buffers while the others sweep; same bits as Normalize -> push -> head(wait), over both payload slots."""
    ops = _ops()
    from advise_video_ssl_b200 import _lib
    g = torch.Generator().manual_seed(B + D)
    T = 0.1
    queue_c = O.l2_normalize(torch.randn(K, D, generator=g))
    x1, x2 = ops.PeerExchange(B, D), ops.PeerExchange(B, D)
    try:
        q1, q2 = queue_c.clone().cuda(), queue_c.clone().cuda()
        p1 = torch.zeros(1, dtype=torch.int64, device="cuda")
        p2 = torch.zeros(1, dtype=torch.int64, device="cuda")
        status = torch.zeros(1, dtype=torch.int32, device="cuda")
        for step in range(3):
            feat = torch.randn(B, D, generator=g).cuda()
            raw = (torch.randn(B, D, generator=g) * 4).cuda()
            perm = torch.randperm(B, generator=g).cuda()
            a = ops.moco_infonce(feat, None, q1, T, True, _lib.IMPL_AUTO, enqueue=(p1, status), peer=x1, push_rows=raw,
                                 peer_row_idx=perm, enq_row_idx=perm)
            x2.push_normalized(raw, 0.0)
            b = ops.moco_infonce(feat, None, q2, T, True, _lib.IMPL_AUTO, enqueue=(p2, status), peer=x2,
                                 peer_row_idx=perm, enq_row_idx=perm)
            for name in ("loss", "dfeat", "q", "logits", "lse"):
                assert torch.equal(a[name], b[name]), (step, name)
            assert torch.equal(q1, q2) and torch.equal(p1, p2)
            assert torch.equal(x1.wait_gather_all(), ops.l2norm_fwd(raw, 0.0)[0])
        assert int(status.item()) == 0
    finally:
        x1.close()
        x2.close()
