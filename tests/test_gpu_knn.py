"""-m gpu parity tests of the kNN evaluation path (second half): eval_knn
(`models/contrastive.py:232-241`) = similarities from the head's tcgen05 mainloop + the exact top-k of `csrc/knn.cu`.

The top-k itself is integer / selection work: values must be the matrix entries bit for bit, and the order is
(value descending, index ascending) -- compared with a stable CPU sort.  The similarity values carry the tolerance of
the 3-term tf32 split (5e-6 absolute on unit rows), so against the oracle's fp32 GEMM the sorted values agree to 2e-5
and every returned index points at a value within 2e-5 of the one reported."""
import pytest
import torch

from helpers import make_cfg, register_backbones
from oracle import contrastive_oracle as O

pytestmark = pytest.mark.gpu


def _stable_topk(mat, k):
    """(value desc, index asc): what `topk_rows` promises; a stable descending sort keeps equal values in index order."""
    v, i = torch.sort(mat.cpu(), dim=1, descending=True, stable=True)
    return v[:, :k], i[:, :k]


@pytest.mark.parametrize("N,M,k", [(3, 1000, 5), (64, 70001, 200), (5, 16384 * 2 + 17, 256), (2, 200, 200), (7, 300, 200),
                                   (1, 40000, 1), (4, 600000, 200), (2, 100000, 1024), (9, 33, 32), (3, 16385, 200)])
def test_topk_rows_is_exact(N, M, k):
    from advise_video_ssl_b200 import ops
    g = torch.Generator().manual_seed(N * 31 + M + k)
    mat = torch.randn(N, M, generator=g)
    yd, yi = ops.topk_rows(mat.cuda(), k)
    rv, ri = _stable_topk(mat, k)
    assert yi.dtype == torch.int64 and tuple(yd.shape) == (N, k)
    assert torch.equal(yd.cpu(), rv) and torch.equal(yi.cpu(), ri)
    tv, ti = mat.cuda().topk(k, dim=1, largest=True, sorted=True)  # the reference's own operator (:240)
    assert torch.equal(yd, tv)
    distinct = (tv[:, 1:] != tv[:, :-1]).all(dim=1) if k > 1 else torch.ones(N, dtype=torch.bool, device="cuda")
    if k < M:  # the boundary value must not be tied with the first element left out either
        kth1 = mat.cuda().topk(k + 1, dim=1).values[:, -1]
        distinct &= kth1 != tv[:, -1]
    assert torch.equal(yi[distinct], ti[distinct])                   # torch leaves the order of ties unspecified


@pytest.mark.parametrize("N,M,k", [(4, 5000, 200), (3, 40000, 64), (2, 256, 200)])
def test_topk_rows_ties_go_to_the_smaller_index(N, M, k):
    from advise_video_ssl_b200 import ops
    g = torch.Generator().manual_seed(M + k)
    mat = torch.randint(-3, 4, (N, M), generator=g).float()  # seven distinct values incl. +-0: ties everywhere
    mat[0].fill_(0.25)                                        # a constant row: the answer is indices 0 .. k-1
    yd, yi = ops.topk_rows(mat.cuda(), k)
    rv, ri = _stable_topk(mat, k)
    assert torch.equal(yd.cpu(), rv) and torch.equal(yi.cpu(), ri)
    assert torch.equal(yi[0].cpu(), torch.arange(k))


def test_topk_rows_strided_rows_and_scale():
    from advise_video_ssl_b200 import ops
    g = torch.Generator().manual_seed(11)
    N, M, k, D = 6, 20001, 200, 32
    full = torch.randn(N, M + 1, generator=g).cuda()
    q = torch.randn(N, D, generator=g).cuda() * 3
    yd, yi = ops.topk_rows(full[:, 1:], k, q_scale_rows=q)      # the layout of logits[:, 1:]
    rv, ri = _stable_topk(full[:, 1:], k)
    _, nrm = ops.l2norm_fwd(q)
    assert torch.equal(yi.cpu(), ri)
    assert torch.equal(yd, rv.cuda() * nrm[:, None])
    with pytest.raises(RuntimeError):
        ops.topk_rows(full[:, :100], 101)                         # k > M, as torch.topk
    with pytest.raises(RuntimeError):
        ops.topk_rows(full.cpu(), 5)                              # no CPU fallback


@pytest.mark.parametrize("N,M,D,scale", [(64, 20037, 128, 1.0), (17, 5000, 64, 3.7), (130, 70000, 128, 1.0), (8, 3000, 256, 1.0)])
def test_knn_similarity_topk_against_oracle(N, M, D, scale):
    from advise_video_ssl_b200 import ops
    g = torch.Generator().manual_seed(N + M + D)
    q = O.l2_normalize(torch.randn(N, D, generator=g)) * scale
    bank = O.l2_normalize(torch.randn(M, D, generator=g))
    k = 200
    rd, ri = O.knn_topk(q, bank, k)
    yd, yi = ops.knn_similarity_topk(q.cuda(), bank.cuda(), k)
    tol = 2e-5 * scale
    assert (yd.cpu() - rd).abs().max().item() < tol
    dist = q @ bank.t()
    assert (dist.gather(1, yi.cpu()) - yd.cpu()).abs().max().item() < tol
    # sorted, indices distinct per row, and identical to the oracle's wherever the neighbouring values are not near-ties
    assert bool((yd[:, 1:] <= yd[:, :-1]).all())
    assert all(len(set(r.tolist())) == k for r in yi.cpu())
    gap_ok = torch.ones(N, k, dtype=torch.bool)
    gap_ok[:, 1:] &= (rd[:, :-1] - rd[:, 1:]) > 2 * tol
    gap_ok[:, :-1] &= (rd[:, :-1] - rd[:, 1:]) > 2 * tol
    assert torch.equal(yi.cpu()[gap_ok], ri[gap_ok])


def _same_neighbours(yd, yi, rd, ri, tol):
    """Sorted values within tol; indices identical wherever the neighbouring reference values are further apart."""
    assert (yd.cpu() - rd).abs().max().item() < tol
    ok = torch.ones_like(ri, dtype=torch.bool)
    gap = (rd[:, :-1] - rd[:, 1:]) > 2 * tol
    ok[:, 1:] &= gap
    ok[:, :-1] &= gap
    assert torch.equal(yi.cpu()[ok], ri[ok]) and ok.float().mean().item() > 0.9


def test_eval_knn_golden(golden):
    """knn.npz: the reference's own eval_knn (k = 200 and 5) and the eval branch of forward, on a 1500-row bank."""
    from advise_video_ssl_b200 import ops
    g = golden("knn")
    for k in (200, 5):
        yd, yi = ops.knn_similarity_topk(g["q"].cuda(), g["bank"].cuda(), k)
        assert yi.dtype == torch.int64
        _same_neighbours(yd, yi, g["yd%d" % k], g["yi%d" % k], 2e-5)
    C = register_backbones()
    N, D, L = int(g.scalar("N")), int(g.scalar("D")), int(g.scalar("L"))
    cfg = make_cfg(CONTRASTIVE__TYPE="moco", CONTRASTIVE__T=0.1, CONTRASTIVE__DIM=D, CONTRASTIVE__QUEUE_LEN=64,
                   CONTRASTIVE__KNN_ON=True, CONTRASTIVE__LENGTH=L)
    model = C.ContrastiveModel(cfg).cuda().eval()
    with torch.no_grad():
        model.backbone.proj.weight.copy_(g["W"])
        model.knn_mem.memory.copy_(g["bank"].view(L, 1, D))
    x = g["x"].cuda()
    yd, yi = model([[x], [x]], torch.arange(N).cuda(), torch.zeros(N, 2, 1).cuda())
    _same_neighbours(yd, yi, g["fwd_yd"], g["fwd_yi"], 2e-5)


def test_eval_knn_through_the_module():
    """ContrastiveModel in eval mode returns (yd, yi) of the kNN bank (:232-241, :469-474)."""
    C = register_backbones()
    D, L = 128, 4099
    cfg = make_cfg(CONTRASTIVE__TYPE="moco", CONTRASTIVE__T=0.1, CONTRASTIVE__DIM=D, CONTRASTIVE__QUEUE_LEN=256,
                   CONTRASTIVE__KNN_ON=True, CONTRASTIVE__LENGTH=L, MODEL__ARCH="identity")
    torch.manual_seed(1)
    model = C.ContrastiveModel(cfg).cuda().eval()
    bank = O.l2_normalize(torch.randn(L, D))
    with torch.no_grad():
        model.knn_mem.memory.copy_(bank.view(L, 1, D))
    x = torch.randn(40, D)
    yd, yi = model([[x.cuda()]], torch.arange(40).cuda(), torch.zeros(40, 2, 1).cuda())
    rd, ri = O.knn_topk(O.l2_normalize(x), bank, 200)
    assert tuple(yd.shape) == (40, 200) and yi.dtype == torch.int64
    assert (yd.cpu() - rd).abs().max().item() < 2e-5
    assert (yi.cpu() == ri).float().mean().item() > 0.99   # near-ties may swap neighbours
    yd2, yi2 = model.eval_knn(O.l2_normalize(x).cuda(), knn_k=5)
    assert (yd2.cpu() - rd[:, :5]).abs().max().item() < 2e-5
