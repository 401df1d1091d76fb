"""CPU-only: the C-ABI library loads, exports every symbol include/avssl_b200.h
declares, the ctypes table matches the header, and compute entry points fail
loudly (no CPU fallback) when there is no CUDA device."""
import ctypes
import os
import re
import subprocess

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "avssl_b200.h")


def _declared():
    src = open(HEADER).read()
    return sorted(set(re.findall(r"AVSSL_API\s+[\w\s\*]+?\b(avssl_\w+)\s*\(", src)))


def test_header_symbols_exported_and_bound():
    from advise_video_ssl_b200 import _lib
    names = _declared()
    assert len(names) >= 20
    out = subprocess.check_output(["nm", "-D", "--defined-only", _lib.LIB_PATH], text=True)
    exported = set(re.findall(r"\bT (avssl_\w+)", out))
    assert set(names) == exported, (set(names) ^ exported)
    assert set(names) == set(_lib.SIGNATURES), (set(names) ^ set(_lib.SIGNATURES))
    # nothing but the C-ABI leaks out of the shared object
    leaked = [l for l in out.splitlines() if " T " in l and "avssl_" not in l]
    assert not leaked, leaked
    assert _lib.lib.avssl_abi_version() == _lib.ABI_VERSION == 2


def test_header_compiles_as_plain_c(tmp_path):
    c = tmp_path / "t.c"
    c.write_text('#include "avssl_b200.h"\nint main(void){ avssl_ema_chunk c; avssl_peer_xchg x; (void)c; (void)x; return AVSSL_OK; }\n')
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-c", str(c),
                           "-o", str(tmp_path / "t.o")])


def test_struct_layout_matches():
    from advise_video_ssl_b200 import _lib
    assert ctypes.sizeof(_lib.EmaChunk) == 24
    assert _lib.EmaChunk.hist.offset == 8 and _lib.EmaChunk.n.offset == 16 and _lib.EmaChunk.flags.offset == 20
    assert _lib.lib.avssl_ema_chunk_elems() == 4096
    # avssl_peer_xchg: 16 pointers + world, rank, rows_per_rank, D + timeout_ms, reserved
    assert ctypes.sizeof(_lib.PeerXchg) == 16 * 8 + 24
    assert _lib.PeerXchg.world.offset == 128 and _lib.PeerXchg.D.offset == 140 and _lib.PeerXchg.timeout_ms.offset == 144
    assert _lib.lib.avssl_peer_xchg_bytes(8, 64, 128) == 256 + 2 * 8 * 64 * 128 * 4
    assert _lib.lib.avssl_peer_xchg_bytes(17, 64, 128) == 0


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback_without_gpu():
    from advise_video_ssl_b200 import _lib, ops
    assert _lib.lib.avssl_device_sm_count() < 0
    assert "no CPU fallback" in _lib.last_error()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.moco_infonce(torch.randn(4, 8), [torch.randn(4, 8)], torch.randn(16, 8), 0.1)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.EmaPlan([torch.randn(4)], [torch.randn(4)])
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.sinkhorn(torch.randn(4, 8), 0.05, 3)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "advise_video_ssl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt, os.path.join(dirpath, f)


def test_only_test_infrastructure_imports_the_oracle():
    """oracle/ is a checker: besides tests/, only __graft_entry__.smoke() and bench.py's CPU legs may use it."""
    allowed = {os.path.join(ROOT, "bench.py"), os.path.join(ROOT, "__graft_entry__.py")}
    for dirpath, dirs, files in os.walk(ROOT):
        dirs[:] = [d for d in dirs if d not in (".git", "gpurun_out", "tests", "oracle", "__pycache__", "build", "baseline")]
        for f in files:
            if not f.endswith(".py"):
                continue
            path = os.path.join(dirpath, f)
            if path in allowed:
                continue
            txt = open(path).read()
            assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, re.M), path
