"""ctypes binding of `libavssl_b200.so` (the C-ABI declared in include/avssl_b200.h).

The library is the product: there is no Python/torch fallback.  If the shared
object is missing the import fails with instructions to build it; if a compute
entry point is called without a CUDA device the library returns an error that is
raised here.
"""
import ctypes
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libavssl_b200.so")

c_void_p, c_int, c_int64, c_float, c_size_t, c_uint64 = (
    ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float, ctypes.c_size_t, ctypes.c_uint64)


class AvsslError(RuntimeError):
    """A C-ABI call returned a non-zero avssl_status."""


class EmaChunk(ctypes.Structure):
    """Mirror of `avssl_ema_chunk` (include/avssl_b200.h)."""
    _fields_ = [("online", c_void_p), ("hist", c_void_p), ("n", ctypes.c_uint32), ("flags", ctypes.c_uint32)]


class PeerXchg(ctypes.Structure):
    """Mirror of `avssl_peer_xchg` (include/avssl_b200.h)."""
    _fields_ = [("base", c_void_p * 16), ("world", c_int), ("rank", c_int), ("rows_per_rank", c_int), ("D", c_int),
                ("timeout_ms", ctypes.c_uint32), ("reserved_", ctypes.c_uint32)]


MAX_PEERS = 16
IPC_HANDLE_BYTES = 64

# name -> (restype, argtypes).  Every function declared in include/avssl_b200.h is
# listed here; tests/test_abi.py checks the two stay in sync.
SIGNATURES = {
    "avssl_abi_version": (c_int, []),
    "avssl_last_error": (ctypes.c_char_p, []),
    "avssl_device_sm_count": (c_int, []),
    "avssl_ema_chunk_elems": (c_int64, []),
    "avssl_ema_plan_chunks": (c_int64, [c_void_p, c_int]),
    "avssl_ema_plan_fill": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int64]),
    "avssl_ema_multi_tensor": (c_int, [c_void_p, c_int64, c_float, c_float, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "avssl_moco_infonce_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "avssl_moco_infonce_fwd_bwd": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_float,
                                           c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t,
                                           c_int, c_void_p]),
    "avssl_moco_infonce_fwd_bwd_enqueue": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int,
                                                   c_int, c_float, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                                   c_void_p, c_size_t, c_int, c_void_p]),
    "avssl_queue_enqueue": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "avssl_membank_update": (c_int, [c_void_p, c_int64, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                                     c_float, c_float, c_int, c_void_p, c_void_p]),
    "avssl_membank_gather_dot": (c_int, [c_void_p, c_int64, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                         c_int, c_int, c_float, c_int, c_void_p, c_void_p, c_void_p]),
    "avssl_l2norm_fwd": (c_int, [c_void_p, c_int, c_int, c_float, c_void_p, c_void_p, c_void_p]),
    "avssl_l2norm_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_float, c_void_p, c_void_p]),
    "avssl_byol_simloss_workspace_bytes": (c_size_t, [c_int]),
    "avssl_byol_simloss_fwd_bwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_float, c_int, c_void_p, c_void_p,
                                           c_void_p, c_size_t, c_void_p]),
    "avssl_ce_target0_workspace_bytes": (c_size_t, [c_int]),
    "avssl_ce_target0_fwd": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "avssl_ce_target0_bwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "avssl_ntxent_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "avssl_ntxent_prepare": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "avssl_ntxent_prepare_peer": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "avssl_ntxent_rowsum": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_float, c_void_p, c_void_p, c_size_t,
                                    c_int, c_void_p]),
    "avssl_ntxent_grad": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_float,
                                  c_void_p, c_void_p, c_void_p, c_size_t, c_int, c_void_p]),
    "avssl_sinkhorn_workspace_bytes": (c_size_t, [c_int, c_int]),
    "avssl_sinkhorn": (c_int, [c_void_p, c_int, c_int, c_float, c_int, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    "avssl_peer_xchg_bytes": (c_size_t, [c_int, c_int, c_int]),
    "avssl_peer_alloc": (c_int, [c_size_t, c_void_p, c_void_p]),
    "avssl_peer_open": (c_int, [c_void_p, c_void_p]),
    "avssl_peer_close": (c_int, [c_void_p]),
    "avssl_peer_free": (c_int, [c_void_p]),
    "avssl_peer_push_rows": (c_int, [c_void_p, c_void_p, c_void_p]),
    "avssl_peer_scatter_bytes": (c_size_t, [c_int, c_int64]),
    "avssl_peer_scatter_exchange": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "avssl_l2norm_push_rows": (c_int, [c_void_p, c_void_p, c_float, c_void_p, c_void_p]),
    "avssl_peer_wait_gather": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "avssl_ema_multi_tensor_push": (c_int, [c_void_p, c_int64, c_float, c_float, c_void_p, c_int, c_int, c_void_p,
                                            c_void_p, c_void_p, c_void_p]),
    "avssl_moco_infonce_fwd_bwd_enqueue_peer": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p,
                                                        c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_void_p, c_void_p, c_void_p,
                                                        c_void_p, c_void_p, c_void_p, c_size_t, c_int, c_void_p]),
    "avssl_moco_infonce_fwd_bwd_enqueue_indexed": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_int,
                                                           c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_float,
                                                           c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                                           c_size_t, c_int, c_void_p]),
    "avssl_moco_infonce_sweep": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_int, c_void_p, c_void_p, c_size_t,
                                         c_int, c_int, c_void_p]),
    "avssl_multi_l2norm_workspace_bytes": (c_size_t, [c_int64, c_int]),
    "avssl_multi_l2norm": (c_int, [c_void_p, c_int64, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "avssl_linear_l2norm_max_dout": (c_int, []),
    "avssl_linear_l2norm_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_int, c_void_p, c_void_p,
                                        c_void_p]),
    "avssl_linear_l2norm_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_int,
                                        c_void_p, c_void_p, c_void_p, c_void_p]),
    "avssl_topk_rows_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "avssl_topk_rows": (c_int, [c_void_p, c_int64, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p,
                                c_size_t, c_void_p]),
    "avssl_swav_ce_workspace_bytes": (c_size_t, [c_int]),
    "avssl_swav_ce_fwd_bwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float, c_void_p, c_void_p,
                                      c_void_p, c_void_p, c_size_t, c_void_p]),
}

IMPL_AUTO, IMPL_SIMT, IMPL_TC3X, IMPL_TC1X = 0, 1, 2, 3


def head_swept(sweep_ctas):
    """AVSSL_HEAD_SWEPT(sweep_ctas): or-ed into `impl` of a head call that follows avssl_moco_infonce_sweep."""
    return 0x100 | (int(sweep_ctas) << 16)
MAX_KEYS = 8
DEVFLAG_QUEUE_OVERRUN = 1
DEVFLAG_BAD_INDEX = 2
DEVFLAG_PEER_TIMEOUT = 4
ABI_VERSION = 2


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "advise_video_ssl_b200: %s is missing. Build it with "
            "`python -c 'import __graft_entry__ as g; g.build()'` or "
            "`python advise_video_ssl_b200/build.py` (needs nvcc, targets sm_100a). "
            "There is no CPU or PyTorch fallback for the contrastive hot path." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so is stale: fail loudly
        fn.restype = res
        fn.argtypes = args
    if lib.avssl_abi_version() != ABI_VERSION:
        raise ImportError("advise_video_ssl_b200: %s has ABI version %d, this package needs %d: rebuild it "
                          "(python advise_video_ssl_b200/build.py)" % (LIB_PATH, lib.avssl_abi_version(), ABI_VERSION))
    return lib


lib = _load()


def last_error():
    msg = lib.avssl_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(rc, what):
    if rc != 0:
        raise AvsslError("%s failed (status %d): %s" % (what, rc, last_error()))
