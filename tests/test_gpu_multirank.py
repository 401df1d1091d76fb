"""-m gpu, self-spawning: when the box has more than one GPU, run tests/mgpu/worker.py under torchrun
(one rank per GPU over NCCL) and require every multi-rank parity check to pass.  On a one-GPU box the
test skips; the committed record of the N = 2 and N = 8 runs is profiles/r2_multirank_parity_n*.json."""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world", [2, 4, 8])
def test_multirank_parity(world, tmp_path):
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs, box has %d" % (world, torch.cuda.device_count()))
    out = tmp_path / "report.json"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tests", "mgpu", "worker.py"), "--out", str(out)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, (r.stdout[-3000:], r.stderr[-6000:])
    rep = json.loads(out.read_text())
    assert rep["ok"] and rep["world"] == world
    assert {"shuffle", "moco", "simclr", "bank"} <= set(rep)
