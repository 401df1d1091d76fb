"""Tensor-level wrappers over the C-ABI (include/avssl_b200.h).

PyTorch is used for device memory and streams only: every function here takes CUDA
tensors, extracts raw pointers and the current stream and calls the library.
CPU tensors are rejected — there is no fallback path.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from ._lib import lib, check

_f32 = torch.float32


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _req(t, name, dtype=_f32):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError("%s must be a CUDA tensor: the contrastive hot path has no CPU fallback" % name)
    if t.dtype != dtype:
        raise TypeError("%s must be %s, got %s" % (name, dtype, t.dtype))
    if not t.is_contiguous():
        raise ValueError("%s must be contiguous" % name)
    return t


def moco_infonce_workspace_bytes(B, D, K, n_keys=1):
    return int(lib.avssl_moco_infonce_workspace_bytes(B, D, K, n_keys))


def sm_count():
    n = lib.avssl_device_sm_count()
    if n < 0:
        raise _lib.AvsslError(_lib.last_error())
    return n


# ----------------------------------------------------------------------------- K1
_CHUNK_DT = np.dtype([("online", np.uint64), ("hist", np.uint64), ("n", np.uint32), ("flags", np.uint32)])
assert _CHUNK_DT.itemsize == ctypes.sizeof(_lib.EmaChunk) == 24


def ema_plan_table(online_ptrs, hist_ptrs, numels):
    """Host-side chunk table (numpy structured array) for a parameter list."""
    n = len(numels)
    numel = np.ascontiguousarray(numels, dtype=np.int64)
    optr = np.ascontiguousarray(online_ptrs, dtype=np.uint64)
    hptr = np.ascontiguousarray(hist_ptrs, dtype=np.uint64)
    n_chunks = lib.avssl_ema_plan_chunks(numel.ctypes.data, n)
    if n_chunks < 0:
        raise ValueError("bad parameter list for the EMA plan")
    table = np.zeros(max(n_chunks, 1), dtype=_CHUNK_DT)
    check(lib.avssl_ema_plan_fill(optr.ctypes.data, hptr.ctypes.data, numel.ctypes.data, n,
                                  table.ctypes.data, n_chunks), "avssl_ema_plan_fill")
    return table[:n_chunks]


class EmaPlan:
    """Device pointer table for the multi-tensor momentum update (K1).

    Built once from the online / history parameter lists (named_parameters()
    order, models/contrastive.py:164-172) and reused every step; rebuild when any
    parameter storage moves (the owning module invalidates it from `_apply`).
    """

    def __init__(self, online, hist):
        if len(online) != len(hist):
            raise ValueError("online and history parameter lists differ in length")
        dev = None
        for o, h in zip(online, hist):
            _req(o, "online parameter")
            _req(h, "history parameter")
            if o.shape != h.shape:
                raise ValueError("online/history shape mismatch: %s vs %s" % (tuple(o.shape), tuple(h.shape)))
            dev = o.device if dev is None else dev
            if o.device != dev or h.device != dev:
                raise ValueError("all EMA tensors must live on one device")
        self.device = dev if dev is not None else torch.device("cuda", torch.cuda.current_device())
        self.n_tensors = len(online)
        self.n_params = int(sum(o.numel() for o in online))
        self.ptr_key = tuple((o.data_ptr(), h.data_ptr()) for o, h in zip(online, hist))
        table = ema_plan_table([o.data_ptr() for o in online], [h.data_ptr() for h in hist],
                               [o.numel() for o in online])
        self.n_chunks = int(table.shape[0])
        host = torch.from_numpy(table.view(np.uint8).reshape(-1).copy()) if self.n_chunks else torch.zeros(24, dtype=torch.uint8)
        self.table = host.to(self.device)
        self.done = torch.zeros(1, dtype=torch.int32, device=self.device)

    def matches(self, online, hist):
        return self.ptr_key == tuple((o.data_ptr(), h.data_ptr()) for o, h in zip(online, hist))

    def run(self, m, iter_buf, bump_iter=False, first_iter=None, push=None):
        """hist <- online*(1-m) + hist*m, bit-exact with the reference's fp32 ops.
        first_iter: True/False when the caller mirrors `iter` on the host, None to let the
        kernel read it from the device.
        push=(PeerExchange, rows): the same launch also pushes `rows` to every peer (C3)."""
        _req(iter_buf, "iter", torch.int64)
        m = float(m)
        fi = -1 if first_iter is None else (1 if first_iter else 0)
        if push is not None:
            xchg, rows = push
            xchg.check_rows(rows)
            check(lib.avssl_ema_multi_tensor_push(self.table.data_ptr(), self.n_chunks, m, 1.0 - m,
                                                  iter_buf.data_ptr(), fi, 1 if bump_iter else 0,
                                                  self.done.data_ptr(), ctypes.addressof(xchg.desc),
                                                  rows.data_ptr(), _stream()), "avssl_ema_multi_tensor_push")
            return
        check(lib.avssl_ema_multi_tensor(self.table.data_ptr(), self.n_chunks, m, 1.0 - m,
                                         iter_buf.data_ptr(), fi, 1 if bump_iter else 0,
                                         self.done.data_ptr(), _stream()), "avssl_ema_multi_tensor")

    @property
    def algorithmic_bytes(self):
        return 12 * self.n_params


class MultiTensorNorm:
    """Device plan for the multi-tensor L2 norm (get_grad_norm_, models/optimizer.py:375-397;
    LARS.step's per-parameter norms, :351-352): one launch over all tensors, no host sync.

        plan = MultiTensorNorm(grads)      # rebuilt only when a tensor's storage moves
        total, per_tensor = plan.run()     # device tensors [1], [n]
    """

    def __init__(self, tensors):
        dev = None
        for t in tensors:
            _req(t, "tensor")
            dev = t.device if dev is None else dev
            if t.device != dev:
                raise ValueError("all tensors must live on one device")
        if dev is None:
            if not torch.cuda.is_available():
                raise RuntimeError("MultiTensorNorm needs a CUDA device: the contrastive hot path has no CPU fallback")
            dev = torch.device("cuda", torch.cuda.current_device())
        self.device = dev
        self.n_tensors = len(tensors)
        self.n_elems = int(sum(t.numel() for t in tensors))
        self.ptr_key = tuple(t.data_ptr() for t in tensors)
        ptrs = [t.data_ptr() for t in tensors]
        numels = [t.numel() for t in tensors]
        table = ema_plan_table(ptrs, ptrs, numels)
        self.n_chunks = int(table.shape[0])
        chunk = int(lib.avssl_ema_chunk_elems())
        first = np.zeros(self.n_tensors + 1, dtype=np.int32)
        first[1:] = np.cumsum([(n + chunk - 1) // chunk for n in numels], dtype=np.int64)
        host = torch.from_numpy(table.view(np.uint8).reshape(-1).copy()) if self.n_chunks else torch.zeros(24, dtype=torch.uint8)
        self.table = host.to(dev)
        self.first = torch.from_numpy(first).to(dev)
        self.ws = torch.zeros(int(lib.avssl_multi_l2norm_workspace_bytes(self.n_chunks, self.n_tensors)), dtype=torch.uint8, device=dev)
        self.per_tensor = torch.zeros(max(self.n_tensors, 1), dtype=_f32, device=dev)
        self.total = torch.zeros(1, dtype=_f32, device=dev)

    def matches(self, tensors):
        return self.ptr_key == tuple(t.data_ptr() for t in tensors)

    def run(self):
        check(lib.avssl_multi_l2norm(self.table.data_ptr(), self.n_chunks, self.first.data_ptr(), self.n_tensors,
                                     self.per_tensor.data_ptr(), self.total.data_ptr(), self.ws.data_ptr(),
                                     self.ws.numel(), _stream()), "avssl_multi_l2norm")
        return self.total, self.per_tensor[:self.n_tensors]

    @property
    def algorithmic_bytes(self):
        return 4 * self.n_elems


# ------------------------------------------------------------------------------ C3
class PeerExchange:
    """Cross-GPU exchange of [rows_per_rank, D] fp32 blocks over NVLink peer memory (C3).

    Replaces `cat_all_gather(keys)` + row select of `_batch_unshuffle`
    (models/contrastive.py:216-230) for the ranks of one box with plain peer stores and epoch
    flags: no collective kernel, no side stream, CUDA-graph capturable.  Collective set-up:
    every rank of `group` must construct it (one all_gather of the 64-byte IPC handles), and
    every rank must then issue the same sequence of pushes.

        x = PeerExchange(B, D)                 # once
        x.push(keys)                           # or EmaPlan.run(..., push=(x, keys))
        k = x.wait_gather(idx_restore_rank)    # or moco_infonce(..., peer=x)

    world_size 1 (or no process group) works too: the exchange then targets its own buffer.
    """

    def __init__(self, rows_per_rank, D, group=None, device=None, timeout_ms=None):
        import os
        import torch.distributed as dist
        if not torch.cuda.is_available():
            raise RuntimeError("PeerExchange needs a CUDA device: the contrastive hot path has no CPU fallback")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        on = dist.is_available() and dist.is_initialized()
        self.world = dist.get_world_size(group) if on else 1
        self.rank = dist.get_rank(group) if on else 0
        self.rows_per_rank, self.D = int(rows_per_rank), int(D)
        if self.world > _lib.MAX_PEERS:
            raise ValueError("PeerExchange supports at most %d ranks (one NVLink domain)" % _lib.MAX_PEERS)
        nbytes = self._buffer_bytes()
        if nbytes == 0:
            raise ValueError("bad exchange geometry world=%d rows=%d D=%d" % (self.world, rows_per_rank, D))
        self.nbytes = int(nbytes)
        base = ctypes.c_void_p()
        handle = (ctypes.c_ubyte * _lib.IPC_HANDLE_BYTES)()
        self._own, self._opened = None, []
        err = None
        with torch.cuda.device(self.device):
            # Every rank reaches every collective below whatever happens locally, and all ranks raise
            # together if any of them failed: a half-built exchange would hang the first push.
            try:
                check(lib.avssl_peer_alloc(self.nbytes, ctypes.byref(base), ctypes.addressof(handle)), "avssl_peer_alloc")
                self._own = base.value
            except Exception as e:  # noqa: BLE001
                err = "rank %d: %s" % (self.rank, e)
            bases = [None] * self.world
            bases[self.rank] = self._own
            if self.world > 1:
                handles = [None] * self.world
                dist.all_gather_object(handles, (self.rank, bytes(handle) if err is None else None, err), group=group)
                errs = [h[2] for h in handles if h[2]]
                if not errs:
                    for r, h, _ in handles:
                        if r == self.rank:
                            continue
                        p = ctypes.c_void_p()
                        buf = (ctypes.c_ubyte * _lib.IPC_HANDLE_BYTES).from_buffer_copy(h)
                        try:
                            check(lib.avssl_peer_open(ctypes.addressof(buf), ctypes.byref(p)), "avssl_peer_open(rank %d)" % r)
                        except Exception as e:  # noqa: BLE001
                            err = "rank %d: %s" % (self.rank, e)
                            break
                        self._opened.append(p.value)
                        bases[r] = p.value
                    # doubles as the barrier: every buffer is mapped everywhere before the first push
                    flags = [None] * self.world
                    dist.all_gather_object(flags, err, group=group)
                    errs = [f for f in flags if f]
                if errs:
                    self.close()
                    raise _lib.AvsslError("PeerExchange set-up failed (CUDA IPC between the ranks of one box is "
                                          "required): " + "; ".join(errs))
            elif err is not None:
                raise _lib.AvsslError(err)
        self.desc = _lib.PeerXchg()
        for r in range(self.world):
            self.desc.base[r] = bases[r]
        self.desc.world, self.desc.rank = self.world, self.rank
        self.desc.rows_per_rank, self.desc.D = self.rows_per_rank, self.D
        # bound of the consumer's spin on a peer's flag (0 = forever); on expiry the kernel sets
        # DEVFLAG_PEER_TIMEOUT in the status word it was given and carries on instead of hanging
        if timeout_ms is None:
            timeout_ms = int(os.environ.get("AVSSL_PEER_TIMEOUT_MS", "30000"))
        self.desc.timeout_ms = max(0, int(timeout_ms))

    def _buffer_bytes(self):
        return lib.avssl_peer_xchg_bytes(self.world, self.rows_per_rank, self.D)

    def check_rows(self, rows):
        _req(rows, "rows")
        if tuple(rows.shape) != (self.rows_per_rank, self.D):
            raise ValueError("exchange block is [%d, %d], got %s" % (self.rows_per_rank, self.D, tuple(rows.shape)))

    def push(self, rows):
        """Store `rows` into every rank's buffer and publish the new epoch (one small launch)."""
        self.check_rows(rows)
        check(lib.avssl_peer_push_rows(ctypes.addressof(self.desc), rows.data_ptr(), _stream()), "avssl_peer_push_rows")

    def push_normalized(self, feat, eps=0.0, keep=None):
        """Normalize + push in one launch: every rank receives feat / max(||feat||, eps), the
        bits `l2norm_fwd` then `push` would have moved.  `keep` ([rows, D], optional) receives
        this rank's normalised rows."""
        self.check_rows(feat)
        if keep is not None:
            self.check_rows(keep)
        check(lib.avssl_l2norm_push_rows(ctypes.addressof(self.desc), feat.data_ptr(), float(eps),
                                         keep.data_ptr() if keep is not None else None, _stream()),
              "avssl_l2norm_push_rows")

    def wait_gather(self, row_idx=None, out=None, status=None):
        """Wait for every rank's push of the current epoch, then return gathered[row_idx]
        (row_idx None: this rank's own block).  gathered = cat_all_gather order."""
        n_out = self.rows_per_rank if row_idx is None else int(row_idx.numel())
        if row_idx is not None:
            _req(row_idx, "row_idx", torch.int64)
        if out is None:
            out = torch.empty(n_out, self.D, dtype=_f32, device=self.device)
        _req(out, "out")
        check(lib.avssl_peer_wait_gather(ctypes.addressof(self.desc), row_idx.data_ptr() if row_idx is not None else None,
                                         n_out, out.data_ptr(), status.data_ptr() if status is not None else None,
                                         _stream()), "avssl_peer_wait_gather")
        return out

    def wait_gather_all(self, out=None):
        """All world*rows_per_rank rows in rank order (what cat_all_gather returns)."""
        idx = torch.arange(self.world * self.rows_per_rank, dtype=torch.int64, device=self.device)
        return self.wait_gather(idx, out=out)

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001  (interpreter shutdown, lost context)
            pass

    def close(self):
        """Unmap the peers' buffers and free this rank's.  Collective in spirit: call it on every
        rank once no push or wait is in flight."""
        if getattr(self, "_own", None) is None and not getattr(self, "_opened", None):
            return
        torch.cuda.synchronize(self.device)
        for p in self._opened:
            lib.avssl_peer_close(p)
        if self._own is not None:
            lib.avssl_peer_free(self._own)
        self._own, self._opened = None, []


class PeerScatter(PeerExchange):
    """C1 over NVLink peer memory: the shuffle-BN exchange of the key encoder's input rows
    (models/contrastive.py:174-214) as a scatter -- every rank writes each of its rows once, straight
    into its final position in the destination rank's buffer; nothing is gathered first or reordered
    on arrival.  Rows are opaque (any dtype, `row_bytes` a multiple of 16).

        sc = PeerScatter(B, row_bytes)                       # collective set-up, once
        shuffled = sc.exchange(x, dest_pos)                  # dest_pos = argsort(perm)[rank*B:(rank+1)*B]
                                                             # == cat_all_gather(x)[perm.view(W, -1)[rank]]
    """

    def __init__(self, rows_per_rank, row_bytes, group=None, device=None, timeout_ms=None):
        if row_bytes % 16 != 0:
            raise ValueError("row_bytes must be a multiple of 16, got %d" % row_bytes)
        self.row_bytes = int(row_bytes)
        super().__init__(rows_per_rank, row_bytes // 4, group=group, device=device, timeout_ms=timeout_ms)

    def _buffer_bytes(self):
        return lib.avssl_peer_scatter_bytes(self.rows_per_rank, self.row_bytes)

    def _check(self, t, name):
        if not t.is_cuda or not t.is_contiguous():
            raise ValueError("%s must be a contiguous CUDA tensor" % name)
        if t.shape[0] != self.rows_per_rank or t.numel() * t.element_size() != self.rows_per_rank * self.row_bytes:
            raise ValueError("%s must hold %d rows of %d bytes, got %s %s" % (name, self.rows_per_rank, self.row_bytes,
                                                                            tuple(t.shape), t.dtype))

    def exchange(self, x, dest_pos, out=None, status=None, dest_pos_host=None):
        """Scatter this rank's rows to their final positions, wait for everybody's, return the received rows
        (one launch; collective: every rank of the group makes the call).  `dest_pos_host` (numpy int64, the
        same values): small batches then carry the positions in the kernel parameters and `dest_pos` may be None."""
        self._check(x, "x")
        host_ptr = None
        if dest_pos_host is not None:
            dest_pos_host = np.ascontiguousarray(dest_pos_host, dtype=np.int64)
            if dest_pos_host.size != self.rows_per_rank:
                raise ValueError("dest_pos_host needs %d entries" % self.rows_per_rank)
            host_ptr = dest_pos_host.ctypes.data
        if dest_pos is not None:
            _req(dest_pos, "dest_pos", torch.int64)
            if dest_pos.numel() != self.rows_per_rank:
                raise ValueError("dest_pos needs %d entries" % self.rows_per_rank)
        elif host_ptr is None:
            raise ValueError("need dest_pos or dest_pos_host")
        if out is None:
            out = torch.empty_like(x, memory_format=torch.contiguous_format)
        self._check(out, "out")
        check(lib.avssl_peer_scatter_exchange(ctypes.addressof(self.desc), x.data_ptr(),
                                              dest_pos.data_ptr() if dest_pos is not None else None, host_ptr,
                                              out.data_ptr(), status.data_ptr() if status is not None else None,
                                              _stream()), "avssl_peer_scatter_exchange")
        return out


# ---------------------------------------------------------------------------- K2+K3
_workspaces = {}


def _workspace(device, nbytes, op):
    """Zero-once scratch per (op, device, stream).  Every op keeps its own: the kernels leave
    their barrier / counter words in an op-specific state between launches (the cooperative
    Sinkhorn's generation word never returns to zero, the InfoNCE grid barrier expects zero),
    so two ops must never see each other's words."""
    key = (op, device.index, torch.cuda.current_stream(device).cuda_stream)
    ws = _workspaces.get(key)
    if (ws is None or ws.numel() < nbytes) and torch.cuda.is_current_stream_capturing():
        # a CUDA-graph capture runs on its own stream: reuse the scratch of the (stream-ordered) warm-up
        # rather than allocating and zero-filling a new one inside the graph on every replay
        for (o, d, _), cand in _workspaces.items():
            if o == op and d == device.index and cand.numel() >= nbytes:
                return cand
    if ws is None or ws.numel() < nbytes:
        ws = torch.zeros(nbytes, dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


def moco_infonce_sweep(feat_q, queue, T, workspace, n_keys=1, logits=None, impl=_lib.IMPL_AUTO, sweep_ctas=0):
    """First launch of the two-launch head (avssl_moco_infonce_sweep): q against the queue on the current stream --
    per-CTA softmax partials into `workspace`, logits[:, 1:] into `logits` ([n_keys*B, K+1] or None).  Needs neither
    the keys nor the momentum encoder; `moco_infonce(..., swept=sweep_ctas)` with the same tensors finishes the head
    on a stream ordered behind this one.  `sweep_ctas`: 0 = every SM, fewer leaves SMs to a kernel running beside it."""
    _req(feat_q, "feat_q")
    _req(queue, "queue")
    _req(workspace, "workspace", torch.uint8)
    B, D = feat_q.shape
    K = queue.shape[0]
    if logits is not None:
        _req(logits, "logits")
        if tuple(logits.shape) != (n_keys * B, K + 1):
            raise ValueError("logits must be [%d, %d], got %s" % (n_keys * B, K + 1, tuple(logits.shape)))
    check(lib.avssl_moco_infonce_sweep(feat_q.data_ptr(), queue.data_ptr(), B, D, K, float(T), int(n_keys),
                                       logits.data_ptr() if logits is not None else None, workspace.data_ptr(),
                                       workspace.numel(), int(impl), int(sweep_ctas), _stream()), "avssl_moco_infonce_sweep")


def moco_infonce(feat_q, keys, queue, T, want_logits=True, impl=_lib.IMPL_AUTO, out=None, enqueue=None, workspace=None,
                 peer=None, peer_row_idx=None, enq_row_idx=None, key_rows=None, keys_raw=False, push_rows=None, swept=None):
    """Fused l2-norm + logits + InfoNCE forward/backward (K2+K3).

    Returns dict(loss[1], dfeat[B,D], q[B,D], lse[n_keys*B], logits[n_keys*B,K+1] or None).
    `out` may carry preallocated tensors under the same names (CUDA-graph replay).
    `enqueue=(ptr, status)` additionally performs K4 for keys[0] after the loss has been
    computed against the old queue: queue[ptr:ptr+B] = keys[0], ptr advanced on the device
    (same launch with the tcgen05 kernels); `status` may be None.
    `workspace`: optional caller-owned uint8 tensor of `moco_infonce_workspace_bytes()` bytes,
    zero-filled once (CUDA-graph capture: no allocation or memset inside the captured region).
    `peer`: a PeerExchange whose current epoch holds the keys (`keys` must be None): the launch
    waits for the exchange itself, after its sweep over the queue, and takes key row i from
    gathered[peer_row_idx[i]] (default: this rank's own block) -- C3 fused into K3.
    `enq_row_idx` (peer only, int64): the rows of the gathered buffer that the fused enqueue writes,
    ptr advancing by their count (C9: rank 0's rows on every rank = the reference's effective
    semantics under DDP's buffer broadcast; all rows = canonical MoCo).  Default: this rank's block.
    `push_rows` (peer only, [rows_per_rank, D]): this rank's RAW key-encoder output; one extra CTA of the
    launch normalises it and performs this step's push into every rank's buffer, so the caller pushes nothing.
    `key_rows` ([n, D], `keys` must be None): keys that live on this device but are still in the key
    encoder's row order -- query row i meets key_rows[peer_row_idx[i]] and the queue receives
    key_rows[enq_row_idx[e]] (the un-shuffle folded into the launch); with `keys_raw` they are the
    encoder's raw output and Normalize is applied inside the kernel as well.
    `swept` (int): `moco_infonce_sweep(..., sweep_ctas=swept)` has already run on these tensors (`workspace`, and
    `out["logits"]` when logits are wanted): this call only merges its partials with the key term.
    """
    if swept is not None:
        if workspace is None or (want_logits and "logits" not in (out or {})):
            raise ValueError("swept: pass the workspace (and out['logits']) the sweep wrote")
        impl = int(impl) | _lib.head_swept(swept)
    _req(feat_q, "feat_q")
    _req(queue, "queue")
    if feat_q.dim() != 2 or queue.dim() != 2 or queue.shape[1] != feat_q.shape[1]:
        raise ValueError("feat_q [B,D] and queue [K,D] expected, got %s and %s" % (tuple(feat_q.shape), tuple(queue.shape)))
    B, D = feat_q.shape
    K = queue.shape[0]
    if peer is not None or key_rows is not None:
        if keys is not None or (peer is not None and key_rows is not None):
            raise ValueError("pass exactly one of keys, peer and key_rows")
        if peer_row_idx is not None:
            _req(peer_row_idx, "peer_row_idx", torch.int64)
            if peer_row_idx.numel() != B:
                raise ValueError("peer_row_idx needs %d entries" % B)
        if key_rows is not None:
            _req(key_rows, "key_rows")
            if key_rows.dim() != 2 or key_rows.shape[1] != D:
                raise ValueError("key_rows must be [n, %d], got %s" % (D, tuple(key_rows.shape)))
        keys = []
    n_keys = 1 if (peer is not None or key_rows is not None) else len(keys)
    if not 1 <= n_keys <= _lib.MAX_KEYS:
        raise ValueError("need 1..%d key tensors, got %d" % (_lib.MAX_KEYS, n_keys))
    for k in keys:
        _req(k, "key")
        if tuple(k.shape) != (B, D):
            raise ValueError("key shape %s != %s" % (tuple(k.shape), (B, D)))
    dev = feat_q.device
    out = out or {}
    loss = out.get("loss") if "loss" in out else torch.empty(1, dtype=_f32, device=dev)
    dfeat = out.get("dfeat") if "dfeat" in out else torch.empty(B, D, dtype=_f32, device=dev)
    q = out.get("q") if "q" in out else torch.empty(B, D, dtype=_f32, device=dev)
    lse = out.get("lse") if "lse" in out else torch.empty(n_keys * B, dtype=_f32, device=dev)
    logits = None
    if want_logits:
        logits = out.get("logits") if "logits" in out else torch.empty(n_keys * B, K + 1, dtype=_f32, device=dev)
    nbytes = lib.avssl_moco_infonce_workspace_bytes(B, D, K, n_keys)
    if workspace is not None:
        ws = _req(workspace, "workspace", torch.uint8)
        if ws.numel() < nbytes:
            raise ValueError("workspace has %d bytes, need %d" % (ws.numel(), nbytes))
    else:
        ws = _workspace(dev, nbytes, "moco_infonce")
    if peer is not None or key_rows is not None:
        ptr, status = enqueue if enqueue is not None else (None, None)
        n_enq = B
        if enq_row_idx is not None:
            _req(enq_row_idx, "enq_row_idx", torch.int64)
            n_enq = int(enq_row_idx.numel())
        if ptr is not None:
            _req(ptr, "ptr", torch.int64)
            assert K % n_enq == 0, "queue length %d is not a multiple of the key batch %d" % (K, n_enq)
        if status is not None:
            _req(status, "status", torch.int32)
    if key_rows is not None:
        check(lib.avssl_moco_infonce_fwd_bwd_enqueue_indexed(
            feat_q.data_ptr(), key_rows.data_ptr(), int(key_rows.shape[0]), 1 if keys_raw else 0,
            peer_row_idx.data_ptr() if peer_row_idx is not None else None,
            enq_row_idx.data_ptr() if enq_row_idx is not None else None, n_enq,
            queue.data_ptr(), ptr.data_ptr() if ptr is not None else None,
            status.data_ptr() if status is not None else None, B, D, K, float(T),
            q.data_ptr(), loss.data_ptr(), dfeat.data_ptr(), lse.data_ptr(),
            logits.data_ptr() if logits is not None else None, ws.data_ptr(), ws.numel(), int(impl), _stream()),
            "avssl_moco_infonce_fwd_bwd_enqueue_indexed")
        return {"loss": loss, "dfeat": dfeat, "q": q, "lse": lse, "logits": logits}
    if peer is not None:
        if push_rows is not None:
            peer.check_rows(push_rows)
        check(lib.avssl_moco_infonce_fwd_bwd_enqueue_peer(
            feat_q.data_ptr(), ctypes.addressof(peer.desc), peer_row_idx.data_ptr() if peer_row_idx is not None else None,
            enq_row_idx.data_ptr() if enq_row_idx is not None else None, n_enq,
            push_rows.data_ptr() if push_rows is not None else None,
            queue.data_ptr(), ptr.data_ptr() if ptr is not None else None,
            status.data_ptr() if status is not None else None, B, D, K, float(T),
            q.data_ptr(), loss.data_ptr(), dfeat.data_ptr(), lse.data_ptr(),
            logits.data_ptr() if logits is not None else None, ws.data_ptr(), ws.numel(), int(impl), _stream()),
            "avssl_moco_infonce_fwd_bwd_enqueue_peer")
        return {"loss": loss, "dfeat": dfeat, "q": q, "lse": lse, "logits": logits}
    key_ptrs = (ctypes.c_void_p * n_keys)(*[k.data_ptr() for k in keys])
    if enqueue is None:
        check(lib.avssl_moco_infonce_fwd_bwd(
            feat_q.data_ptr(), ctypes.addressof(key_ptrs), n_keys, queue.data_ptr(), B, D, K, float(T),
            q.data_ptr(), loss.data_ptr(), dfeat.data_ptr(), lse.data_ptr(),
            logits.data_ptr() if logits is not None else None, ws.data_ptr(), ws.numel(), int(impl), _stream()),
            "avssl_moco_infonce_fwd_bwd")
    else:
        ptr, status = enqueue
        _req(ptr, "ptr", torch.int64)
        if status is not None:
            _req(status, "status", torch.int32)
        # models/contrastive.py:284 — same exception type as the reference's assert
        assert K % B == 0, "queue length %d is not a multiple of the key batch %d" % (K, B)
        check(lib.avssl_moco_infonce_fwd_bwd_enqueue(
            feat_q.data_ptr(), ctypes.addressof(key_ptrs), n_keys, queue.data_ptr(), ptr.data_ptr(),
            status.data_ptr() if status is not None else None, B, D, K, float(T),
            q.data_ptr(), loss.data_ptr(), dfeat.data_ptr(), lse.data_ptr(),
            logits.data_ptr() if logits is not None else None, ws.data_ptr(), ws.numel(), int(impl), _stream()),
            "avssl_moco_infonce_fwd_bwd_enqueue")
    return {"loss": loss, "dfeat": dfeat, "q": q, "lse": lse, "logits": logits}


# ----------------------------------------------------------------------------- K4
def queue_enqueue(queue, ptr, keys, status=None):
    """queue[ptr:ptr+n] = keys; ptr advances on the device (K4)."""
    _req(queue, "queue")
    _req(keys, "keys")
    _req(ptr, "ptr", torch.int64)
    if status is not None:
        _req(status, "status", torch.int32)
    n, D = keys.shape
    K = queue.shape[0]
    if queue.shape[1] != D:
        raise ValueError("keys dim %d != queue dim %d" % (D, queue.shape[1]))
    # models/contrastive.py:284 — same exception type as the reference's assert
    assert K % n == 0, "queue length %d is not a multiple of the key batch %d" % (K, n)
    check(lib.avssl_queue_enqueue(queue.data_ptr(), ptr.data_ptr(), keys.data_ptr(), n, K, D,
                                  status.data_ptr() if status is not None else None, _stream()),
          "avssl_queue_enqueue")


# ------------------------------------------------------------------------ K5 / K14
def _bank_dims(bank):
    if bank.dim() == 3:
        return bank.shape[0], bank.shape[1], bank.shape[2]
    if bank.dim() == 2:
        return bank.shape[0], 1, bank.shape[1]
    raise ValueError("bank must be [L, duration, D] or [L, D]")


def membank_update(bank, mem, ind, time, momentum, interp=False, status=None):
    """bank[ind, time] <- l2norm(mem*m + old*(1-m)); last duplicate wins (K5)."""
    _req(bank, "bank")
    L, dur, D = _bank_dims(bank)
    mem = _req(mem.reshape(mem.shape[0], -1), "mem")
    if mem.shape[1] != D:
        raise ValueError("mem dim %d != bank dim %d" % (mem.shape[1], D))
    n = mem.shape[0]
    ind = _req(ind.reshape(-1).long().contiguous(), "ind", torch.int64)
    if ind.numel() != n:
        raise ValueError("ind has %d entries for %d rows" % (ind.numel(), n))
    ti = tf = None
    if interp:
        tf = _req(time.reshape(-1).float().contiguous(), "time", _f32)
    elif time is not None:
        ti = _req(time.reshape(-1).long().contiguous(), "time", torch.int64)
    if status is not None:
        _req(status, "status", torch.int32)
    m = float(momentum)
    check(lib.avssl_membank_update(bank.data_ptr(), L, dur, D, mem.data_ptr(), ind.data_ptr(),
                                   ti.data_ptr() if ti is not None else None,
                                   tf.data_ptr() if tf is not None else None, n, m, 1.0 - m,
                                   1 if interp else 0, status.data_ptr() if status is not None else None,
                                   _stream()), "avssl_membank_update")


def membank_gather_dot(bank, q, ind, time, T, interp=False, status=None):
    """prod[n,k] = q_n . bank[ind[n,k], time[n,k]] / T (K14)."""
    _req(bank, "bank")
    _req(q, "q")
    L, dur, D = _bank_dims(bank)
    B = q.shape[0]
    ind = _req(ind.reshape(B, -1).long().contiguous(), "ind", torch.int64)
    Kp = ind.shape[1]
    ti = tf = None
    if interp:
        tf = _req(time.reshape(B, -1).float().contiguous(), "time", _f32)
    elif time is not None:
        ti = _req(time.reshape(B, -1).long().contiguous(), "time", torch.int64)
    prod = torch.empty(B, Kp, dtype=_f32, device=q.device)
    check(lib.avssl_membank_gather_dot(bank.data_ptr(), L, dur, D, q.data_ptr(), ind.data_ptr(),
                                       ti.data_ptr() if ti is not None else None,
                                       tf.data_ptr() if tf is not None else None, B, Kp, float(T),
                                       1 if interp else 0, prod.data_ptr(),
                                       status.data_ptr() if status is not None else None, _stream()),
          "avssl_membank_gather_dot")
    return prod


# ------------------------------------------------------------------------- kNN evaluation (eval_knn)
def topk_rows_supported(N, M, k):
    return 1 <= k <= M and N >= 1 and int(lib.avssl_topk_rows_workspace_bytes(int(N), int(M), int(k))) > 0


def topk_rows(dist, k, q_scale_rows=None):
    """Exact row-wise top-k (largest, sorted, ties towards the smaller index) of a [N, M] fp32 matrix whose rows are
    contiguous (stride(1) == 1; a column-offset view such as logits[:, 1:] is fine).  `q_scale_rows` [N, D]: the
    values are multiplied by ||q_row|| (similarities computed on normalised queries).  Returns (yd [N, k] fp32,
    yi [N, k] int64) like torch.topk (models/contrastive.py:240)."""
    if not isinstance(dist, torch.Tensor) or not dist.is_cuda:
        raise RuntimeError("dist must be a CUDA tensor: the contrastive hot path has no CPU fallback")
    if dist.dtype != _f32 or dist.dim() != 2 or dist.stride(1) != 1:
        raise ValueError("topk_rows: dist must be a 2-D fp32 tensor with contiguous rows")
    N, M = dist.shape
    ld = dist.stride(0) if N > 1 else max(M, dist.stride(0))
    k = int(k)
    if not 1 <= k <= M:
        raise RuntimeError("topk_rows: k = %d not in [1, %d]" % (k, M))  # torch.topk raises here as well
    q_ptr, D = None, 0
    if q_scale_rows is not None:
        _req(q_scale_rows, "q_scale_rows")
        if q_scale_rows.dim() != 2 or q_scale_rows.shape[0] != N:
            raise ValueError("topk_rows: q_scale_rows must be [N, D]")
        q_ptr, D = q_scale_rows.data_ptr(), q_scale_rows.shape[1]
    nbytes = int(lib.avssl_topk_rows_workspace_bytes(N, M, k))
    if nbytes == 0:
        raise _lib.AvsslError("topk_rows: k=%d over M=%d is outside the two-pass plan" % (k, M))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dist.device)
    yd = torch.empty(N, k, dtype=_f32, device=dist.device)
    yi = torch.empty(N, k, dtype=torch.int64, device=dist.device)
    check(lib.avssl_topk_rows(dist.data_ptr(), int(ld), N, M, k, q_ptr, int(D), yd.data_ptr(), yi.data_ptr(), ws.data_ptr(),
                              nbytes, _stream()), "avssl_topk_rows")
    return yd, yi


def knn_similarity_topk(q, bank, k, impl=_lib.IMPL_AUTO):
    """eval_knn (models/contrastive.py:232-241): top-k of q bank^T.  The similarity matrix is produced by the head's
    tcgen05 mainloop (avssl_moco_infonce_sweep with the bank in the queue's place and T = 1: fp32-grade 3-term split,
    logits[:, 1:]); shapes it does not take (D outside {32, 64, 96, 128}) go through a library GEMM.  The top-k is
    `topk_rows` either way."""
    _req(q, "q")
    _req(bank, "bank")
    N, D = q.shape
    M = bank.shape[0]
    if bank.dim() != 2 or bank.shape[1] != D:
        raise ValueError("knn_similarity_topk: q [N, D] and bank [M, D] expected")
    if D in (32, 64, 96, 128) and impl != _lib.IMPL_SIMT and q.data_ptr() % 16 == 0 and bank.data_ptr() % 16 == 0:
        logits = torch.empty(N, M + 1, dtype=_f32, device=q.device)
        ws = _workspace(q.device, moco_infonce_workspace_bytes(N, D, M, 1), "knn")
        moco_infonce_sweep(q, bank, 1.0, ws, n_keys=1, logits=logits, impl=impl, sweep_ctas=0)
        return topk_rows(logits[:, 1:], k, q_scale_rows=q)
    return topk_rows(q @ bank.t(), k)


# ------------------------------------------------------------------------- K2 / K7
def l2norm_fwd(x, eps=0.0):
    _req(x, "x")
    n, D = x.shape
    y = torch.empty_like(x)
    nrm = torch.empty(n, dtype=_f32, device=x.device)
    check(lib.avssl_l2norm_fwd(x.data_ptr(), n, D, float(eps), y.data_ptr(), nrm.data_ptr(), _stream()),
          "avssl_l2norm_fwd")
    return y, nrm


def l2norm_bwd(y, nrm, dy, eps=0.0):
    _req(y, "y")
    _req(nrm, "norm")
    dy = _req(dy.contiguous(), "dy")
    n, D = y.shape
    dx = torch.empty_like(y)
    check(lib.avssl_l2norm_bwd(y.data_ptr(), nrm.data_ptr(), dy.data_ptr(), n, D, float(eps), dx.data_ptr(),
                               _stream()), "avssl_l2norm_bwd")
    return dx


def linear_l2norm_supported(in_features, out_features):
    """Shapes the fused projection tail handles (avssl_linear_l2norm_fwd): Dout <= 256, Kin % 4 == 0."""
    return out_features <= lib.avssl_linear_l2norm_max_dout() and in_features % 4 == 0


def linear_l2norm_fwd(x, weight, bias=None, eps=0.0, normalize=True):
    """q = Normalize(x W^T + b) in one launch, the raw projection is never written (projection tail,
    models/head_helper.py:52-58 + models/contrastive.py:923-934).  Returns (q [B, Dout], ||y|| [B])."""
    _req(x, "x")
    _req(weight, "weight")
    if bias is not None:
        _req(bias, "bias")
    if x.dim() != 2 or weight.dim() != 2 or x.shape[1] != weight.shape[1]:
        raise ValueError("linear_l2norm: x [B, Kin] and weight [Dout, Kin] expected, got %s and %s"
                         % (tuple(x.shape), tuple(weight.shape)))
    B, Kin = x.shape
    Dout = weight.shape[0]
    if bias is not None and tuple(bias.shape) != (Dout,):
        raise ValueError("linear_l2norm: bias must be [%d]" % Dout)
    q = torch.empty(B, Dout, dtype=_f32, device=x.device)
    nrm = torch.empty(B, dtype=_f32, device=x.device)
    check(lib.avssl_linear_l2norm_fwd(x.data_ptr(), weight.data_ptr(), bias.data_ptr() if bias is not None else None,
                                      B, Kin, Dout, float(eps), 1 if normalize else 0, q.data_ptr(), nrm.data_ptr(),
                                      _stream()), "avssl_linear_l2norm_fwd")
    return q, nrm


def linear_l2norm_bwd(x, weight, q, nrm, grad_q, eps=0.0, normalize=True, need_dx=True, need_dw=True, need_db=True):
    """Backward of `linear_l2norm_fwd` in one launch: the gradient of the normalisation is applied where grad_q is
    read, then dx = dy W, dW = dy^T x, db = sum_b dy.  Returns (dx, dW, db), None where not requested."""
    _req(x, "x")
    _req(weight, "weight")
    _req(q, "q")
    _req(nrm, "norm")
    grad_q = _req(grad_q.contiguous(), "grad_q")
    B, Kin = x.shape
    Dout = weight.shape[0]
    if tuple(grad_q.shape) != (B, Dout) or tuple(q.shape) != (B, Dout) or nrm.numel() != B:
        raise ValueError("linear_l2norm_bwd: shape mismatch")
    dx = torch.empty_like(x) if need_dx else None
    dW = torch.empty_like(weight) if need_dw else None
    db = torch.empty(Dout, dtype=_f32, device=x.device) if need_db else None
    if B == 0:
        return (dx, dW.zero_() if dW is not None else None, db.zero_() if db is not None else None)
    ptr = lambda t: t.data_ptr() if t is not None else None  # noqa: E731
    check(lib.avssl_linear_l2norm_bwd(x.data_ptr(), weight.data_ptr(), q.data_ptr(), nrm.data_ptr(), grad_q.data_ptr(),
                                      B, Kin, Dout, float(eps), 1 if normalize else 0, ptr(dx), ptr(dW), ptr(db),
                                      _stream()), "avssl_linear_l2norm_bwd")
    return dx, dW, db


def byol_simloss(pred, key, T, normalize=True, want_grad=True):
    """loss = -mean(p.k)/T with p = pred/||pred|| when `normalize` (K7). Returns (loss[1], dpred|None)."""
    _req(pred, "pred")
    _req(key, "key")
    if pred.shape != key.shape or pred.dim() != 2:
        raise ValueError("pred and key must both be [B, D]")
    n, D = pred.shape
    loss = torch.empty(1, dtype=_f32, device=pred.device)
    dpred = torch.empty_like(pred) if want_grad else None
    ws = _workspace(pred.device, lib.avssl_byol_simloss_workspace_bytes(n), "byol_simloss")
    check(lib.avssl_byol_simloss_fwd_bwd(pred.data_ptr(), key.data_ptr(), n, D, float(T), 1 if normalize else 0,
                                         loss.data_ptr(), dpred.data_ptr() if want_grad else None,
                                         ws.data_ptr(), ws.numel(), _stream()), "avssl_byol_simloss_fwd_bwd")
    return loss, dpred


def ce_target0_fwd(logits):
    """Mean cross-entropy against class 0 (ContrastiveLoss). Returns (loss[1], row_lse[n])."""
    _req(logits, "logits")
    n, C = logits.shape
    loss = torch.empty(1, dtype=_f32, device=logits.device)
    lse = torch.empty(n, dtype=_f32, device=logits.device)
    ws = _workspace(logits.device, lib.avssl_ce_target0_workspace_bytes(n), "ce_target0")
    check(lib.avssl_ce_target0_fwd(logits.data_ptr(), n, C, loss.data_ptr(), lse.data_ptr(), ws.data_ptr(),
                                   ws.numel(), _stream()), "avssl_ce_target0_fwd")
    return loss, lse


def ce_target0_bwd(logits, lse, grad_out):
    _req(logits, "logits")
    _req(lse, "row_lse")
    g = _req(grad_out.reshape(1).contiguous(), "grad_out")
    n, C = logits.shape
    d = torch.empty_like(logits)
    check(lib.avssl_ce_target0_bwd(logits.data_ptr(), lse.data_ptr(), n, C, g.data_ptr(), d.data_ptr(), _stream()),
          "avssl_ce_target0_bwd")
    return d


# ----------------------------------------------------------------------------- K6
_ntx_const = {}


def _ntx_rows(rank, B, N, dev):
    """Global row ids of this rank's 2B rows of [q_all ; q2_all] (int32, cached)."""
    key = ("rows", rank, B, N, dev)
    t = _ntx_const.get(key)
    if t is None:
        t = _ntx_const[key] = torch.cat([torch.arange(rank * B, (rank + 1) * B, dtype=torch.int32, device=dev),
                                         torch.arange(N + rank * B, N + (rank + 1) * B, dtype=torch.int32, device=dev)])
    return t


def _ntx_zperm(world, dev):
    """Row gather that turns the exchanged [W][2][B] row sums into [2][W][B] (the order of [q_all ; q2_all])."""
    key = ("zperm", world, dev)
    t = _ntx_const.get(key)
    if t is None:
        t = _ntx_const[key] = torch.tensor([w * 2 + v for v in range(2) for w in range(world)], dtype=torch.int64, device=dev)
    return t


def ntxent(feat1, feat2, T, gather=None, impl=_lib.IMPL_AUTO, xchg=None, status=None):
    """SimCLR NT-Xent (K6) with the cross-rank gather (C4) and the reduce-scatter-
    equivalent gradient (C5) folded in.  feat1/feat2: this rank's raw [B, D] features.
    Returns (loss[1], dfeat1, dfeat2); the gradients carry the reference's world-size
    factor (utils/distributed.py:142-155).
    `xchg` = (PeerExchange(2B, D), PeerExchange(2, B)) over WORLD: both gathers (rows, row sums) go over NVLink
    peer stores instead of NCCL all_gather kernels (one box)."""
    import torch.distributed as dist
    _req(feat1, "feat1")
    _req(feat2, "feat2")
    if feat1.shape != feat2.shape or feat1.dim() != 2:
        raise ValueError("feat1 and feat2 must both be [B, D]")
    B, D = feat1.shape
    dev = feat1.device
    world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
    if gather is None:
        gather = world > 1
    if not gather:
        world = 1
    rank = dist.get_rank() if world > 1 else 0
    y, nrm = l2norm_fwd(torch.cat([feat1, feat2], 0), 0.0)  # [2B, D] unit rows, local
    N = world * B
    out = torch.empty(2 * N, D, dtype=_f32, device=dev)
    out_r = torch.empty(2 * N, D, dtype=torch.float16, device=dev)
    peer = world > 1 and xchg is not None
    if peer:
        rows_x, z_x = xchg
        if rows_x.world != world or z_x.world != world:
            raise ValueError("the exchanges span %d ranks, the gather %d" % (rows_x.world, world))
        rows_x.push(y)
        # wait for every rank's block, then [q_all ; q2_all] and its fp16 copy in one pass over the exchange buffer
        check(lib.avssl_ntxent_prepare_peer(ctypes.addressof(rows_x.desc), status.data_ptr() if status is not None else None,
                                            B, D, out.data_ptr(), out_r.data_ptr(), _stream()), "avssl_ntxent_prepare_peer")
    else:
        if world > 1:
            allq = torch.empty(world, 2, B, D, dtype=_f32, device=dev)
            dist.all_gather_into_tensor(allq.view(world * 2 * B, D), y)
        else:
            allq = y
        # [q_all ; q2_all] and its fp16 copy (the tensor cores' operand) in one pass over the gathered rows
        check(lib.avssl_ntxent_prepare(allq.data_ptr(), world, B, D, out.data_ptr(), out_r.data_ptr(), _stream()),
              "avssl_ntxent_prepare")
    rows = _ntx_rows(rank, B, N, dev)
    n_loc = 2 * B
    ws = _workspace(dev, lib.avssl_ntxent_workspace_bytes(2 * N, D, n_loc), "ntxent")
    z_loc = torch.empty(n_loc, dtype=_f32, device=dev)
    check(lib.avssl_ntxent_rowsum(out.data_ptr(), out_r.data_ptr(), rows.data_ptr(), rank * B, N + rank * B, 2 * N, D, n_loc, float(T),
                                  z_loc.data_ptr(), ws.data_ptr(), ws.numel(), int(impl), _stream()), "avssl_ntxent_rowsum")
    if peer:
        z_x.push(z_loc.view(2, B))
        z_all = z_x.wait_gather(_ntx_zperm(world, dev), status=status).view(-1)
    elif world > 1:
        zg = torch.empty(world, 2, B, dtype=_f32, device=dev)
        dist.all_gather_into_tensor(zg.view(-1), z_loc)
        z_all = zg.permute(1, 0, 2).reshape(-1).contiguous()
    else:
        z_all = z_loc
    loss = torch.empty(1, dtype=_f32, device=dev)
    dfeat = torch.empty(n_loc, D, dtype=_f32, device=dev)
    check(lib.avssl_ntxent_grad(out.data_ptr(), out_r.data_ptr(), rows.data_ptr(), rank * B, N + rank * B, z_all.data_ptr(), nrm.data_ptr(),
                                2 * N, D, n_loc, float(T), float(world), loss.data_ptr(), dfeat.data_ptr(), ws.data_ptr(),
                                ws.numel(), int(impl), _stream()), "avssl_ntxent_grad")
    return loss, dfeat[:B], dfeat[B:]


# ---------------------------------------------------------------------- K10 / K11
def sinkhorn(scores, eps, iters, keep_last=None, lane=0):
    """codes = sinkhorn(exp(scores/eps)) (K10).  scores [Btot, P]; returns the last
    `keep_last` rows (default all), each summing to 1.  `lane`: calls that may run concurrently
    (different streams) pass different lanes and get separate scratch."""
    _req(scores, "scores")
    Btot, P = scores.shape
    keep = Btot if keep_last is None else int(keep_last)
    out = torch.empty(keep, P, dtype=_f32, device=scores.device)
    ws = _workspace(scores.device, lib.avssl_sinkhorn_workspace_bytes(Btot, P), "sinkhorn" if lane == 0 else "sinkhorn/%d" % lane)
    check(lib.avssl_sinkhorn(scores.data_ptr(), Btot, P, float(eps), int(iters), keep, out.data_ptr(),
                             ws.data_ptr(), ws.numel(), _stream()), "avssl_sinkhorn")
    return out


def sinkhorn_distributed(Q, iters):
    """distributed_sinkhorn (models/contrastive.py:889-910), multi-node only
    (cfg.NUM_SHARDS > 1; never taken on one 8-GPU box).  Q: [P, B_local]
    = exp(scores/eps)^T on the device.  Cold path: device tensor ops with the row-sum
    all_reduce between the half-steps; the single-box hot path is `sinkhorn`."""
    import torch.distributed as dist
    if not Q.is_cuda:
        raise RuntimeError("Q must be a CUDA tensor: the contrastive hot path has no CPU fallback")
    world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1

    def allsum(t):
        if world > 1:
            dist.all_reduce(t)
        return t
    Q = Q / allsum(torch.sum(Q))
    r = 1.0 / Q.shape[0]
    c = 1.0 / (world * Q.shape[1])
    cur = allsum(torch.sum(Q, dim=1))
    for _ in range(iters):
        Q = Q * (r / cur).unsqueeze(1)
        Q = Q * (c / torch.sum(Q, dim=0)).unsqueeze(0)
        cur = allsum(torch.sum(Q, dim=1))
    return (Q / torch.sum(Q, dim=0, keepdim=True)).t().float()


def swav_pair_weights(n_crops, n_assign, bs):
    """w[a][v] of models/contrastive.py:672-679: mean over bs, /(n_crops-1), /n_assign;
    assign crop a is crop a (swav_crops_for_assign = arange(2))."""
    w = np.zeros((n_assign, n_crops), dtype=np.float32)
    for a in range(n_assign):
        for v in range(n_crops):
            if v != a:
                w[a, v] = 1.0 / (bs * (n_crops - 1) * n_assign)
    return w


def swav_ce(scores, codes, n_crops, bs, T, pair_w=None, want_grad=True):
    """SwAV swapped-prediction cross-entropy, forward + backward (K11).
    scores [n_crops*bs, P]; codes [n_assign, bs, P].  Returns (loss[1], dscores|None)."""
    _req(scores, "scores")
    codes = _req(codes.contiguous(), "codes")
    n_assign = codes.shape[0]
    P = scores.shape[1]
    if scores.shape[0] != n_crops * bs or tuple(codes.shape[1:]) != (bs, P):
        raise ValueError("scores %s / codes %s do not match n_crops=%d bs=%d" %
                         (tuple(scores.shape), tuple(codes.shape), n_crops, bs))
    if pair_w is None:
        pair_w = swav_pair_weights(n_crops, n_assign, bs) if n_crops > 1 else np.full((1, 1), 1.0 / bs, np.float32)
    pair_w = np.ascontiguousarray(pair_w, dtype=np.float32)
    loss = torch.empty(1, dtype=_f32, device=scores.device)
    d = torch.empty_like(scores) if want_grad else None
    ws = _workspace(scores.device, lib.avssl_swav_ce_workspace_bytes(n_crops * bs), "swav_ce")
    check(lib.avssl_swav_ce_fwd_bwd(scores.data_ptr(), codes.data_ptr(), n_crops, n_assign, bs, P, float(T),
                                    pair_w.ctypes.data, loss.data_ptr(), d.data_ptr() if want_grad else None,
                                    ws.data_ptr(), ws.numel(), _stream()), "avssl_swav_ce_fwd_bwd")
    return loss, d
