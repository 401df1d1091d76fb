"""CPU-only checks of the bench.py contract: the reference arm prints ONE JSON line with the keys the
driver reads, and the GPU arm refuses to run without a CUDA device (no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + list(args), capture_output=True, text=True,
                          timeout=600, env=e, cwd=ROOT)


def test_reference_arm_prints_one_contract_line():
    r = _run("--impl", "reference", "--steps", "3", "--warmup", "3")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "contrastive_head_clips_per_sec" and d["unit"] == "clips/s"
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["steps"] == 3
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_other_ranks_exit_quietly():
    r = _run("--impl", "reference", "--gpus", "2", "--steps", "3", "--warmup", "3", env={"RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_gpu_arm_has_no_cpu_fallback():
    r = _run("--steps", "3", "--warmup", "3")
    assert r.returncode != 0
    assert "no CPU fallback" in (r.stderr + r.stdout)


def test_committed_bench_lines_carry_every_contract_key():
    """The bench lines kept under profiles/ (produced by bench.py on the B200 box) have every key the
    driver and the judge read; a key dropped from bench.py would show up here on the next refresh."""
    prof = os.path.join(ROOT, "profiles")
    top = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
           "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline"}
    for name, n in (("r1_bench_graph.json", 1), ("r1_bench_n2.json", 2), ("r1_bench_n4.json", 4), ("r1_bench_n8.json", 8)):
        d = json.loads(open(os.path.join(prof, name)).read().strip().splitlines()[-1])
        assert top <= set(d), (name, top - set(d))
        assert d["n_gpus"] == n and d["metric"] == "contrastive_head_clips_per_sec" and d["dtype"] == "f32"
        assert d["scaling"] == "weak" and d["vs_baseline"] is None and d["data"] == "synthetic"
        assert "workload" in d["config"] and "model" not in d["config"]
        assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(d["e2e"])
        assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
        assert d["e2e"]["value"] < d["value"]  # the end-to-end number is measured, not a copy of `value`
        assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
        assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
        r = d["roofline"]
        assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r) and r["bound"] == "hbm"
        assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and 0.5 < r["frac"] < 1.0
        assert d["gpu_launches"] == 2 * d["steps"]
        assert abs(d["value"] - n * 64 / (d["ms_per_step"] * 1e-3)) < 1e-6 * d["value"]
        if n == 1:
            assert {"value", "unit", "cores", "kind", "sample"} <= set(d["cpu_baseline"])
        else:
            assert d["config"]["key_exchange_verified_vs_nccl"] is True


def test_round2_bench_lines_go_through_the_module_api():
    """Round 2: `value` / `e2e` are measured through contrastive_forward + backward; the kernel-only number sits
    beside them, the strict end-to-end number is at the top level, and the N > 1 lines record the exchanges."""
    prof = os.path.join(ROOT, "profiles")
    top = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
           "vs_baseline", "dtype", "data", "config", "e2e", "e2e_strict", "ops_level", "gpu_launches", "clocks", "roofline",
           "step_roofline"}
    vals = {}
    for n in (1, 2, 4, 8):
        d = json.loads(open(os.path.join(prof, "r2_bench_n%d.json" % n)).read().strip().splitlines()[-1])
        assert top <= set(d), (n, top - set(d))
        assert d["n_gpus"] == n and d["metric"] == "contrastive_head_clips_per_sec" and d["dtype"] == "f32"
        assert "contrastive_forward" in d["config"]["api"] and d["config"]["cuda_graph"] is True
        assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
        assert d["e2e"]["value"] < d["value"] and d["e2e_strict"]["value"] <= d["e2e"]["value"]
        assert d["ops_level"]["step_us"] < d["step_us"]  # the module adds autograd and the shuffle, never removes work
        assert abs(d["value"] - n * 64 / (d["ms_per_step"] * 1e-3)) < 1e-6 * d["value"]
        assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
        r = d["roofline"]
        assert r["bound"] == "hbm" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and 0.5 < r["frac"] < 1.0
        assert d["gpu_launches"] > 0
        vals[n] = d["value"]
        if n == 1:
            assert d["step_roofline"]["frac"] >= 0.75
            assert {"value", "unit", "cores", "kind", "sample"} <= set(d["cpu_baseline"])
        else:
            assert "NVLink" in d["config"]["key_exchange"]
    assert vals[8] / (8 * vals[1]) >= 0.85  # north_star: weak-scaling efficiency at 8 GPUs
    for n in (2, 8):
        rep = json.loads(open(os.path.join(prof, "r2_multirank_parity_n%d.json" % n)).read())
        assert rep["ok"] and rep["world"] == n and {"shuffle", "moco", "simclr", "bank"} <= set(rep)
        assert all(v["queues_identical"] for k, v in rep["moco"].items() if not k.startswith("local"))
