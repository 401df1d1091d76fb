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

    def run(self, m, iter_buf, bump_iter=False):
        """hist <- online*(1-m) + hist*m, bit-exact with the reference's fp32 ops."""
        _req(iter_buf, "iter", torch.int64)
        m = float(m)
        check(lib.avssl_ema_multi_tensor(self.table.data_ptr(), self.n_chunks, m, 1.0 - m,
                                         iter_buf.data_ptr(), 1 if bump_iter else 0,
                                         self.done.data_ptr(), _stream()), "avssl_ema_multi_tensor")

    @property
    def algorithmic_bytes(self):
        return 12 * self.n_params


# ---------------------------------------------------------------------------- K2+K3
_workspaces = {}


def _workspace(device, nbytes):
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.zeros(nbytes, dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


def moco_infonce(feat_q, keys, queue, T, want_logits=True, impl=_lib.IMPL_AUTO, out=None):
    """Fused l2-norm + logits + InfoNCE forward/backward (K2+K3).

    Returns dict(loss[1], dfeat[B,D], q[B,D], lse[n_keys*B], logits[n_keys*B,K+1] or None).
    `out` may carry preallocated tensors under the same names (CUDA-graph replay).
    """
    _req(feat_q, "feat_q")
    _req(queue, "queue")
    if feat_q.dim() != 2 or queue.dim() != 2 or queue.shape[1] != feat_q.shape[1]:
        raise ValueError("feat_q [B,D] and queue [K,D] expected, got %s and %s" % (tuple(feat_q.shape), tuple(queue.shape)))
    B, D = feat_q.shape
    K = queue.shape[0]
    n_keys = len(keys)
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
    ws = _workspace(dev, nbytes)
    key_ptrs = (ctypes.c_void_p * n_keys)(*[k.data_ptr() for k in keys])
    check(lib.avssl_moco_infonce_fwd_bwd(
        feat_q.data_ptr(), ctypes.addressof(key_ptrs), n_keys, queue.data_ptr(), B, D, K, float(T),
        q.data_ptr(), loss.data_ptr(), dfeat.data_ptr(), lse.data_ptr(),
        logits.data_ptr() if logits is not None else None, ws.data_ptr(), ws.numel(), int(impl), _stream()),
        "avssl_moco_infonce_fwd_bwd")
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
