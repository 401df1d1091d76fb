"""CPU ORACLE — test infrastructure, not product code.

A functional CPU restatement (torch CPU tensor ops, fp32 or fp64) of the
contrastive hot path of the reference `models/contrastive.py`,
`models/losses.py:15-25` and `utils/distributed.py:79-155`.  Every function cites
the reference lines it follows.  It exists to CHECK the CUDA path:

  * only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` /
    `--impl reference` legs of `bench.py` may import it;
  * the product package (`advise_video_ssl_b200/`) never imports it and has no
    CPU fallback.

Parity status: PINNED.  The reference ships no tests or golden vectors
(SURVEY.md §4), so the pin is the unmodified reference itself, imported in the
authoring container through `tests/golden/ref_shim.py` and run on seeded inputs
by `tests/golden/make_golden.py`; the resulting vectors are committed under
`tests/golden/*.npz` and `tests/test_oracle_golden.py` checks every function
below against them (bit-exact where the op order is identical, which is
everywhere except where noted).

The reference is a torch program, so the restatement uses the same torch CPU
operators in the same order: that is what makes it bit-identical to the
reference executed on CPU, and what makes it a fair `cpu_baseline` (same BLAS,
same threads).  Gradients come from torch autograd exactly as in the reference;
closed-form fp64 gradients are provided separately as an independent check.
"""
import math

import torch


# ----------------------------------------------------------------------------- K2
def l2_normalize(x, dim=1):
    """Normalize.forward, models/contrastive.py:923-934 (power 2, no eps)."""
    nrm = x.pow(2).sum(dim, keepdim=True).pow(1.0 / 2)
    return x.div(nrm)


# ----------------------------------------------------------------------------- K1
def ema_update(online, hist, m, it):
    """_update_history, models/contrastive.py:158-172.

    online / hist: lists of tensors in named_parameters() order.  At it == 0 the
    history is first overwritten with the online weights (:167-169) and THEN
    blended (:171-172), so step 0 yields w*(1-m) + w*m, not w.
    Returns the new list (the reference rebinds .data, it never updates in place).
    """
    if it == 0:
        hist = [o.clone() for o in online]
    return [o * (1.0 - m) + h * m for o, h in zip(online, hist)]


def momentum_cosine(m0, epoch_exact, max_epoch):
    """momentum_anneal_cosine, models/contrastive.py:251-261 (host fp64)."""
    return 1 - (1 - m0) * (math.cos(math.pi * epoch_exact / max_epoch) + 1.0) * 0.5


# ----------------------------------------------------------------------------- K3
def moco_logits(q, keys, queue, T):
    """MoCo score computation, models/contrastive.py:486-498.

    q [B,D] normalised queries; keys: list of [B,D]; queue [K,D].
    Returns logits [len(keys)*B, K+1] already divided by T.
    """
    neg = torch.einsum("nc,kc->nk", [q, queue.clone().detach()])
    blocks = []
    for key in keys:
        pos = torch.einsum("nc,nc->n", [q, key]).unsqueeze(-1)
        blocks.append(torch.cat([pos, neg], dim=1))
    logits = blocks[0] if len(blocks) == 1 else torch.cat(blocks, dim=0)
    return torch.div(logits, T)


def info_nce(logits):
    """ContrastiveLoss.forward, models/losses.py:20-25: CE against class 0, mean."""
    tgt = torch.zeros(logits.shape[0], dtype=torch.long)
    return torch.nn.functional.cross_entropy(logits, tgt, reduction="mean")


def moco_head(feat_q, keys, queue, T):
    """l2-norm + logits + loss, models/contrastive.py:462-500. Differentiable in feat_q."""
    q = l2_normalize(feat_q)
    logits = moco_logits(q, keys, queue, T)
    return q, logits, info_nce(logits)


def moco_head_closed_form(feat_q, keys, queue, T):
    """Independent fp64 closed form of loss and d loss / d feat_q.

    dq_i = sum_k [ sum_j p_kij queue_j + p_ki0 key_ki - key_ki ] / (T n_rows)
    df_i = (dq_i - (dq_i . q_i) q_i) / ||f_i||
    """
    f = feat_q.double()
    Q = queue.double()
    nrm = f.pow(2).sum(1, keepdim=True).sqrt()
    q = f / nrm
    s = q @ Q.t() / T
    n_rows = len(keys) * f.shape[0]
    loss = 0.0
    dq = torch.zeros_like(q)
    for key in keys:
        k = key.double()
        s0 = (q * k).sum(1, keepdim=True) / T
        m = torch.maximum(s.max(1, keepdim=True).values, s0)
        e = torch.exp(s - m)
        e0 = torch.exp(s0 - m)
        Z = e.sum(1, keepdim=True) + e0
        loss = loss + (torch.log(Z) + m - s0).sum()
        dq = dq + ((e / Z) @ Q + (e0 / Z) * k - k) / (T * n_rows)
    loss = loss / n_rows
    df = (dq - (dq * q).sum(1, keepdim=True) * q) / nrm
    return loss, df, q


# ----------------------------------------------------------------------------- K4
def enqueue(queue, ptr, keys, K, multi_view=False, extra_keys=None):
    """_dequeue_and_enqueue, models/contrastive.py:263-292.  In place on `queue`.

    Only keys[0] is written unless MOCO_MULTI_VIEW_QUEUE (:266-279).  The pointer
    wraps only when it lands exactly on K (:290-291).  Returns the new int ptr.
    """
    ptr = int(ptr)
    todo = [keys[0]]
    if multi_view:
        assert len(keys) > 0
        todo = list(keys)
        if extra_keys:
            todo += [k for sub in extra_keys for k in sub]
    for key in todo:
        n = int(key.size(0))
        assert K % n == 0
        assert ptr + n <= K
        queue[ptr:ptr + n, :] = key
        ptr += n
        if ptr == K:
            ptr = 0
    return ptr


# ------------------------------------------------ TemporalModel path
def temporal_ema_update(online, hist, m, first):
    """TemporalModel._update_history, models/temporal_modeling.py:217-238.

    online / hist: name-matched parameter lists of temporal_encoder + head_projector and their
    *_hist twins.  On the first call (`init_flag` absent, :226-232) the history is overwritten
    with the online weights and THEN blended (:234-237), exactly like ContrastiveModel at iter 0.
    """
    if first:
        hist = [o.clone() for o in online]
    return [o * (1.0 - m) + h * m for o, h in zip(online, hist)]


def temporal_contrast_loss(qs_raw, ks_raw, T):
    """TemporalModel.contrast_forward, models/temporal_modeling.py:354-375, after the heads:
    qs_raw[i] = head_predictor(head_projector(feat_i)), ks_raw[i] = head_projector_hist(key_i)
    with the keys already reversed (:356).  loss = mean_i( -mean_n(l2(q).l2(k)) / T ) + 1/T."""
    loss = 0.0
    for q, k in zip(qs_raw, ks_raw):
        q = l2_normalize(q)
        k = l2_normalize(k)
        sim = torch.einsum("nc,nc->n", [q, k])
        sim = sim / T
        loss = loss + (-sim.mean())
    return loss / len(qs_raw) + 1.0 / T


def grad_norm(grads, norm_type=2.0):
    """get_grad_norm_, models/optimizer.py:375-397 (p-norm branch): norm of the stacked
    per-tensor norms.  An empty list gives tensor(0.0) (:380-381)."""
    if len(grads) == 0:
        return torch.tensor(0.0)
    return torch.norm(torch.stack([torch.norm(g.detach(), norm_type) for g in grads]), norm_type)


# ----------------------------------------------------------------------------- K7
def byol_sim_loss(p, k, T):
    """sim_loss, models/contrastive.py:243-249: -mean_n(sum_c p k) / T."""
    sim = torch.einsum("nc,nc->n", [p, k])
    sim = sim / T
    return -sim.mean()


def byol_pair_loss(pred1_raw, pred2_raw, key1, key2, T):
    """Symmetric BYOL branch, models/contrastive.py:572-582:
    sim_loss(l2(pred1), key2) + sim_loss(l2(pred2), key1)."""
    return byol_sim_loss(l2_normalize(pred1_raw), key2, T) + byol_sim_loss(
        l2_normalize(pred2_raw), key1, T)


def dummy_logits(n, K):
    """K8, models/contrastive.py:585-592: [n, K+1], column 0 = 9999, rest 0."""
    return torch.cat((9999.0 * torch.ones((n, 1), dtype=torch.float),
                      torch.zeros((n, K), dtype=torch.float)), dim=1)


# ----------------------------------------------------------------------------- K6
def ntxent(q, q2, T):
    """SimCLR live branch, models/contrastive.py:775-792 (q, q2 normalised [N,D],
    already gathered across ranks).  Raw exp, diagonal removed by mask, no max
    subtraction."""
    out = torch.cat([q, q2], dim=0)
    sim = torch.exp(torch.mm(out, out.t().contiguous()) / T)
    keep = (torch.ones_like(sim) - torch.eye(out.shape[0])).bool()
    sim = sim.masked_select(keep).view(out.shape[0], -1)
    pos = torch.exp(torch.sum(q * q2, dim=-1) / T)
    pos = torch.cat([pos, pos], dim=0)
    return (-torch.log(pos / sim.sum(dim=-1))).mean()


def ntxent_closed_form(q, q2, T):
    """fp64 closed form of the NT-Xent loss and d loss/d out (out=[q;q2]).

    G_r = [ sum_{c!=r} e^{s_rc} (1/Z_r + 1/Z_c) o_c  - 2 o_{r+} ] / (2N T)
    """
    o = torch.cat([q, q2], 0).double()
    n2 = o.shape[0]
    n = n2 // 2
    e = torch.exp(o @ o.t() / T)
    e.fill_diagonal_(0.0)
    Z = e.sum(1)
    partner = torch.cat([torch.arange(n, n2), torch.arange(0, n)])
    spos = (o * o[partner]).sum(1) / T
    loss = (torch.log(Z) - spos).mean()
    w = e * (1.0 / Z[:, None] + 1.0 / Z[None, :])
    g = (w @ o - 2.0 * o[partner]) / (n2 * T)
    return loss, g, Z


# ---------------------------------------------------------------------------- K10
def sinkhorn(Q, iters):
    """sinkhorn, models/contrastive.py:872-887.  Q: [B, P] = exp(scores/eps).
    Returns [B, P] float32 with every row (sample) summing to 1."""
    Q = Q.t()
    Q = Q / torch.sum(Q)
    r = torch.ones(Q.shape[0], dtype=Q.dtype) / Q.shape[0]
    c = torch.ones(Q.shape[1], dtype=Q.dtype) / Q.shape[1]
    for _ in range(iters):
        Q = Q * (r / torch.sum(Q, dim=1)).unsqueeze(1)
        Q = Q * (c / torch.sum(Q, dim=0)).unsqueeze(0)
    Q = Q / torch.sum(Q, dim=0, keepdim=True)
    return Q.t().float()


def distributed_sinkhorn_emulated(Q_parts, iters):
    """distributed_sinkhorn, models/contrastive.py:889-910, with the all_reduce
    (utils/distributed.py:90-106, SUM) emulated over a list of per-rank [P, B_r]
    matrices.  Returns the list of per-rank [B_r, P] codes."""
    W = len(Q_parts)
    tot = sum(torch.sum(Q) for Q in Q_parts)
    Q_parts = [Q / tot for Q in Q_parts]
    P = Q_parts[0].shape[0]
    r = torch.ones(P, dtype=Q_parts[0].dtype) / P
    cs = [torch.ones(Q.shape[1], dtype=Q.dtype) / (W * Q.shape[1]) for Q in Q_parts]
    cur = sum(torch.sum(Q, dim=1) for Q in Q_parts)
    for _ in range(iters):
        u = cur
        Q_parts = [Q * (r / u).unsqueeze(1) for Q in Q_parts]
        Q_parts = [Q * (c / torch.sum(Q, dim=0)).unsqueeze(0) for Q, c in zip(Q_parts, cs)]
        cur = sum(torch.sum(Q, dim=1) for Q in Q_parts)
    return [(Q / torch.sum(Q, dim=0, keepdim=True)).t().float() for Q in Q_parts]


# ----------------------------------------------------------------------- K9/K11/K12
def swav_renorm_prototypes(W):
    """models/contrastive.py:617-621: rows of the prototype matrix to unit l2."""
    return torch.nn.functional.normalize(W.clone(), dim=1, p=2)


def swav_scores(feat, W):
    """run_swav_orig_encoder_q, models/contrastive.py:865-870: F.normalize (eps
    1e-12) then the bias-free prototype Linear."""
    x = torch.nn.functional.normalize(feat, dim=1, p=2)
    return x, torch.nn.functional.linear(x, W)


def swav_queue_push(queue_i, emb):
    """models/contrastive.py:659-664: FIFO shift by bs, newest rows first."""
    bs = emb.shape[0]
    out = queue_i.clone()
    out[bs:] = queue_i[:-bs]
    out[:bs] = emb
    return out


def swav_loss(output, bs, n_crops, T, eps=0.05, iters=3, queue=None, W=None,
              embedding=None, use_queue=False):
    """SwAV public-code branch, models/contrastive.py:631-679.

    output [n_crops*bs, P] prototype scores (differentiable), two assign crops
    (crops_for_assign = arange(n_crops - (n_crops-2)) = [0, 1]).  When `queue`
    ([2, L, D]) is given and use_queue, queue scores are prepended (:647-658) and
    the queue is pushed (:659-664); returns (loss, codes list, new queue).
    NOTE log(softmax()) not log_softmax.
    """
    loss = 0
    codes = []
    new_queue = None if queue is None else queue.clone()
    assign = list(range(n_crops - (n_crops - 2)))
    for i, crop in enumerate(assign):
        with torch.no_grad():
            out = output[bs * crop: bs * (crop + 1)]
            if queue is not None:
                if use_queue:
                    out = torch.cat((torch.mm(new_queue[i], W.t()), out))
                new_queue[i] = swav_queue_push(new_queue[i], embedding[crop * bs:(crop + 1) * bs])
            q = torch.exp(out / eps).t()
            q = sinkhorn(q.t(), iters)[-bs:]
        codes.append(q)
        sub = 0
        for v in [c for c in range(n_crops) if c != crop]:
            p = torch.softmax(output[bs * v: bs * (v + 1)] / T, dim=1)
            sub = sub - torch.mean(torch.sum(q * torch.log(p), dim=1))
        loss = loss + sub / (n_crops - 1)
    loss = loss / len(assign)
    return loss, codes, new_queue


# ------------------------------------------------------------------------ K5 / K14
def membank_get(memory, ind, time, interp=False):
    """Memory.get, models/contrastive.py:966-987.  memory [L, dur, D]."""
    bsz = ind.size(0)
    if interp:
        t0 = time.floor().long().clamp(0, memory.shape[1] - 1)
        t1 = (t0 + 1).clamp(0, memory.shape[1] - 1)
        m0 = memory[ind.view(-1), t0.view(-1), :]
        m1 = memory[ind.view(-1), t1.view(-1), :]
        w1 = 1 - (time - t0).view(-1, 1).float()
        sel = m0 * (1 - w1) + m1 * w1
    else:
        sel = memory[ind.view(-1), time.long().view(-1), :]
    return sel.view(bsz, -1, memory.shape[2])


def membank_update(memory, mem, momentum, ind, time, interp=False):
    """Memory.update, models/contrastive.py:989-1036 (post all_gather).  In place.
    Duplicate (ind,time) pairs: the CPU index_put applies rows in order, so the
    LAST occurrence wins; every update is computed from the OLD bank rows."""
    if interp:
        t0 = time.floor().long().clamp(0, memory.shape[1] - 1)
        t1 = (t0 + 1).clamp(0, memory.shape[1] - 1)
        m0 = memory[ind.view(-1), t0.view(-1), :]
        m1 = memory[ind.view(-1), t1.view(-1), :]
        w1 = 1 - (time - t0).view(-1, 1).float()
        w0 = 1 - w1
        u0 = l2_normalize(mem * w0 * momentum + m0 * (1 - momentum), dim=1)
        u1 = l2_normalize(mem * w1 * momentum + m1 * (1 - momentum), dim=1)
        memory[ind.view(-1), t0.view(-1), :] = u0.squeeze()
        memory[ind.view(-1), t1.view(-1), :] = u1.squeeze()
    else:
        mem = mem.view(mem.size(0), 1, -1)
        old = membank_get(memory, ind, time, interp=False)
        upd = l2_normalize(mem * momentum + old * (1 - momentum), dim=2)
        memory[ind.view(-1), time.long().view(-1), :] = upd.squeeze()
    return memory


def memory1d_update(memory, mem, momentum, ind):
    """Memory1D.update, models/contrastive.py:1066-1080 (post all_gather). In place."""
    mem = mem.view(mem.size(0), -1)
    ind = ind.long()
    old = torch.index_select(memory, 0, ind.view(-1))
    upd = l2_normalize(old * (1 - momentum) + mem * momentum, dim=1)
    memory.index_copy_(0, ind, upd)
    return memory


def mem_mode_prod(q, memory, clip_ind, time_ind, T, one_d):
    """mem branch, models/contrastive.py:429-434: prod[n,k] = q_n . bank[ind[n,k]] / T."""
    B = q.shape[0]
    if one_d:
        k = torch.index_select(memory, 0, clip_ind.view(-1)).view(B, -1, memory.shape[-1])
    else:
        k = membank_get(memory, clip_ind, time_ind, False)
    return torch.div(torch.einsum("nc,nkc->nk", q, k), T)


def knn_topk(q, bank, k=200):
    """eval_knn, models/contrastive.py:232-241."""
    d = torch.einsum("nc,mc->nm", q.view(q.size(0), -1), bank.view(bank.size(0), -1))
    return d.topk(k, dim=1, largest=True, sorted=True)


# ---------------------------------------------------------------- projection tail
def projection_tail(x, weight, bias=None):
    """Last Linear of MLPHead (models/head_helper.py:52-58, nn.Linear: x W^T + b) followed by the head's
    Normalize (models/contrastive.py:923-934, applied to the backbone output at :462 / :350 / :757)."""
    return l2_normalize(torch.nn.functional.linear(x, weight, bias), dim=1)


# ---------------------------------------------------------------- shuffle (A6, C1-C3)
def shuffle_plan(perm, world):
    """_batch_shuffle index math, models/contrastive.py:203-211: rank r keeps the
    rows perm.view(W,-1)[r] of the gathered batch; idx_restore = argsort(perm)."""
    take = perm.view(world, -1)
    restore = torch.argsort(perm.view(-1)).view(world, -1)
    return take, restore


def shuffle_emulated(x_parts, perm):
    """_batch_shuffle over a list of per-rank tensors (all_gather emulated by cat)."""
    W = len(x_parts)
    allx = torch.cat(x_parts, 0)
    take, restore = shuffle_plan(perm, W)
    return [allx[take[r]] for r in range(W)], restore


def unshuffle_emulated(y_parts, restore):
    """_batch_unshuffle, models/contrastive.py:216-230."""
    ally = torch.cat(y_parts, 0)
    return [ally[restore[r]] for r in range(len(y_parts))]


def allgather_with_gradient_bwd(grad_full_per_rank, rank):
    """AllGatherWithGradient.backward, utils/distributed.py:142-155: SUM over ranks
    of the full gathered gradient, then this rank's row slice."""
    tot = sum(grad_full_per_rank)
    W = len(grad_full_per_rank)
    mb = tot.size(0) // W
    return tot[rank * mb:(rank + 1) * mb]


# --------------------------------------------------- whole MoCo head step (cfg1/cfg2)
class MoCoHeadStep:
    """One full MoCo contrastive-head step on CPU — the `cpu_baseline` workload:
    EMA (K1) -> l2norm + logits + InfoNCE fwd/bwd (K2,K3) -> enqueue (K4),
    in the reference's order (models/contrastive.py:443-506, :308-316)."""

    def __init__(self, online, queue, T, m, dtype=torch.float32):
        self.online = [o.to(dtype) for o in online]
        self.hist = [torch.zeros_like(o) for o in self.online]
        self.queue = queue.to(dtype).clone()
        self.K = queue.shape[0]
        self.T, self.m = T, m
        self.ptr = 0
        self.it = 0

    def step(self, feat_q, keys):
        self.hist = ema_update(self.online, self.hist, self.m, self.it)
        self.it += 1
        f = feat_q.detach().clone().requires_grad_(True)
        q, logits, loss = moco_head(f, keys, self.queue, self.T)
        loss.backward()
        self.ptr = enqueue(self.queue, self.ptr, keys, self.K)
        return loss.detach(), f.grad, logits.detach(), q.detach()
