#!/usr/bin/env python3
"""Benchmark of the contrastive hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

Workload (configs[1] of BASELINE.json, head only — SURVEY.md §8(d) cfg2):
  Slow-R50 MoCo, 2 views per clip, queue 65536, dim 128, batch 64 per GPU.
  One "step" = momentum EMA over the reference's real Slow-R50+MLP-head parameter
  list (164 tensors, 36,095,168 fp32) -> [key all_gather when N>1] -> l2-norm +
  q.[k;queue]^T logits + InfoNCE forward AND backward (df) -> enqueue.  Backbone
  forward/backward are out of scope and excluded.  Synthetic embeddings, random
  weights.

`value`  : clips/s = N * 64 / step time, inputs already resident in HBM.
`e2e`    : same metric through the public module API with pinned HOST embeddings:
           H2D of the step's inputs and D2H of the loss inside the timed region,
           one host synchronisation per step.
`roofline`: the dominant kernel (the EMA, 93 % of the step's bytes): algorithmic
           bytes 12 B/param per launch / CUDA-event duration of that launch.
`cpu_baseline` / `--impl reference`: the CPU oracle port of the same step
           (torch CPU ops, all host threads), bounded sample.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

B_PER_GPU, DIM, QUEUE_LEN, TEMP, MOMENTUM = 64, 128, 65536, 0.1, 0.999
POOL = 8  # distinct synthetic batches rotated through the steps
EMA_DRAM_TRAFFIC = 383_585_280  # bytes per launch of the EMA kernel measured by ncu (289.0 MB read + 94.6 MB written, profiles/r1_ema_ncu.md)
WORKLOAD = ("configs[1] head: Slow-R50 MoCo, 2 views/clip, queue 65536, dim 128, batch 64/GPU; "
            "EMA(164 tensors, 36.1M fp32) + l2norm + logits + InfoNCE fwd/bwd + enqueue; backbone excluded")
METRIC, UNIT = "contrastive_head_clips_per_sec", "clips/s"


def param_shapes(tag="slow_r50_moco_dim128"):
    with open(os.path.join(ROOT, "tests", "golden", "slow_r50_param_shapes.json")) as f:
        return [tuple(s) for _, s in json.load(f)[tag]["shapes"]]


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------ clock sampling
class ClockSampler:
    """Samples SM clock and throttle reasons during the timed region (NVML)."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4),
            "hw_power_brake": getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80),
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(0.02)

    def start(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()

    def stop(self):
        if self._thr is not None:
            self._stop.set()
            self._thr.join()
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


# ------------------------------------------------------------------ CPU oracle arm
def cpu_step_factory(seed=0):
    from oracle import contrastive_oracle as O
    g = torch.Generator().manual_seed(seed)
    online = [torch.randn(s, generator=g) * 0.02 for s in param_shapes()]
    stdv = 1.0 / (DIM / 3) ** 0.5
    queue = torch.rand(QUEUE_LEN, DIM, generator=g).mul_(2 * stdv).add_(-stdv)
    head = O.MoCoHeadStep(online, queue, TEMP, MOMENTUM)
    feats = [torch.randn(B_PER_GPU, DIM, generator=g) for _ in range(POOL)]
    keys = [O.l2_normalize(torch.randn(B_PER_GPU, DIM, generator=g)) for _ in range(POOL)]

    def step(i):
        return head.step(feats[i % POOL], [keys[i % POOL]])
    return step


def run_cpu(steps, warmup, budget_s=None):
    torch.set_num_threads(os.cpu_count() or 1)
    step = cpu_step_factory()
    for i in range(warmup):
        step(i)
    t0 = time.perf_counter()
    done = 0
    for i in range(steps):
        step(warmup + i)
        done += 1
        if budget_s is not None and time.perf_counter() - t0 > budget_s and done >= 3:
            break
    dt = (time.perf_counter() - t0) / done
    return dt, done


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    steps = min(args.steps, 40)
    warmup = min(args.warmup, 3)
    dt, done = run_cpu(steps, warmup, budget_s=120.0)
    val = B_PER_GPU / dt
    cores = torch.get_num_threads()
    sample = "%d full steps (EMA 36.1M params + head B=64,K=65536,D=128 fwd/bwd + enqueue) after %d warm-up" % (done, warmup)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": done, "warmup": warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "note": "CPU oracle port (torch CPU ops = the reference's own op sequence), rank 0 only"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))
    return 0


# ------------------------------------------------------------------------ GPU arm
def gpu_arm(args):
    import torch.distributed as dist
    from advise_video_ssl_b200 import ops, _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the contrastive hot path has no CPU fallback "
                         "(use --impl reference for the CPU oracle arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    json_out = sys.stdout
    if world > 1:
        # NCCL writes its version banner to fd 1 when the communicator is created: keep the real stdout
        # for the ONE JSON line and point fd 1 at stderr for everything else.
        sys.stdout.flush()
        json_out = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus or world == 1, "launch with torchrun --nproc-per-node N for --gpus N"
    n_gpus = world

    g = torch.Generator().manual_seed(1000 + rank)
    online = [(torch.randn(s, generator=g) * 0.02).to(dev) for s in param_shapes()]
    hist = [torch.zeros_like(o) for o in online]
    n_params = sum(o.numel() for o in online)
    stdv = 1.0 / (DIM / 3) ** 0.5
    gq = torch.Generator().manual_seed(7)
    queue = torch.rand(QUEUE_LEN, DIM, generator=gq).mul_(2 * stdv).add_(-stdv).to(dev)
    feats_h = [torch.randn(B_PER_GPU, DIM, generator=g).pin_memory() for _ in range(POOL)]
    keys_h = [torch.nn.functional.normalize(torch.randn(B_PER_GPU, DIM, generator=g)).pin_memory() for _ in range(POOL)]
    feats = [t.to(dev) for t in feats_h]
    keys = [t.to(dev) for t in keys_h]
    it = torch.zeros(1, dtype=torch.int64, device=dev)
    ptr = torch.zeros(1, dtype=torch.int64, device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    plan = ops.EmaPlan(online, hist)
    # C3, the cross-GPU key gather.  "peer" (default): the keys are stored straight into every
    # rank's exchange buffer over NVLink by extra CTAs of the EMA launch, and the head launch waits
    # for them after its sweep -- no collective kernel, no side stream.  "nccl": all_gather on a
    # high-priority side stream next to the EMA (kept for comparison).
    use_peer = world > 1 and args.exchange == "peer"
    xchg, exchange_note = None, None
    if use_peer:
        try:
            xchg = ops.PeerExchange(B_PER_GPU, DIM)  # raises on every rank together if any rank cannot map its peers
        except Exception as e:  # noqa: BLE001  (e.g. CUDA IPC not permitted in this container): NCCL path, and say so
            use_peer, exchange_note = False, "peer exchange unavailable, NCCL all_gather used: %s" % e
    use_nccl = world > 1 and not use_peer
    gathered = torch.empty(world * B_PER_GPU, DIM, device=dev) if world > 1 else None
    comm = torch.cuda.Stream(device=dev, priority=-1) if use_nccl else None
    exchange_verified = None
    if use_peer:  # one untimed round against NCCL's all_gather: bit-identical or abort
        dist.all_gather_into_tensor(gathered, keys[0])
        xchg.push(keys[0])
        exchange_verified = bool(torch.equal(xchg.wait_gather_all(), gathered))
        assert exchange_verified, "peer exchange differs from NCCL all_gather"
    impl = {"auto": _lib.IMPL_AUTO, "simt": _lib.IMPL_SIMT, "tc3x": _lib.IMPL_TC3X, "tc1x": _lib.IMPL_TC1X}[args.kernel]
    out = {}
    head_ws = torch.zeros(ops.moco_infonce_workspace_bytes(B_PER_GPU, DIM, QUEUE_LEN, 1), dtype=torch.uint8, device=dev)
    state = {"n": 0}
    launches_per_step = 2 if args.kernel != "simt" else 4  # ema + fused head (simt: ema, split, combine, enqueue)
    if use_peer and args.kernel == "simt":
        raise SystemExit("--exchange peer needs the tcgen05 head kernel (fused wait)")
    ema_events = []

    def ema_part(k_for_gather, time_ema=False, after=None, push=True):
        """K1 (+ C3: the key push fused into the same launch, or NCCL's all_gather on the side stream)."""
        if time_ema:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        if use_nccl:
            # C3: key all_gather on the comm stream, overlapped with the EMA kernel
            comm.wait_stream(torch.cuda.current_stream())
            if after is not None:
                comm.wait_event(after)
            with torch.cuda.stream(comm):
                dist.all_gather_into_tensor(gathered, k_for_gather)
        plan.run(MOMENTUM, it, bump_iter=True, first_iter=state["n"] == 0,  # host mirror of `iter`, as the module keeps
                 push=(xchg, k_for_gather) if (use_peer and push) else None)
        state["n"] += 1
        if time_ema:
            e1.record()
            ema_events.append((e0, e1))

    def head_part(f, k):
        """K2+K3+K4 in one cooperative launch: loss/grad against the old queue, then the ring write."""
        if use_peer:  # the launch itself waits for every rank's keys and reads this rank's block
            return ops.moco_infonce(f, None, queue, TEMP, want_logits=not args.no_logits, impl=impl, out=out,
                                    enqueue=(ptr, status), workspace=head_ws, peer=xchg)
        if use_nccl:
            torch.cuda.current_stream().wait_stream(comm)
            k = gathered[rank * B_PER_GPU:(rank + 1) * B_PER_GPU]
        return ops.moco_infonce(f, [k], queue, TEMP, want_logits=not args.no_logits, impl=impl, out=out,
                                enqueue=(ptr, status), workspace=head_ws)

    def step(i, f, k, time_ema=False):
        """EMA -> [gather keys] -> fused head + enqueue (reference order, :308-316, :486-503)."""
        ema_part(k, time_ema)
        return head_part(f, k)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- device-resident timing (value)
    for i in range(args.warmup):
        r = step(i, feats[i % POOL], keys[i % POOL])
        if not out:
            out.update(r)  # the step's output tensors are reused from here on (static for graph capture)
    sync_all()

    # One CUDA graph per input slot: [key all_gather on the comm stream ||] EMA -> fused head.  Replaying
    # it removes the per-launch host work (which bounds the step once the NCCL enqueue is added) and
    # the launch gaps between the kernels; the kernels and their order are the same as in eager mode.
    graphs, pool_graph, graph_err = None, None, None
    if not args.no_graph:
        try:
            graphs = []
            for slot in range(POOL):
                g_ = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g_):
                    step(slot, feats[slot], keys[slot])
                graphs.append(g_)
            # ... and one graph holding all POOL steps back to back: relaunching one graph is cheaper on the
            # host than alternating between POOL of them, which matters once 8 ranks submit work at the same
            # time (N=8: 107 us/step with alternating single-step graphs, same kernels).
            pool_graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(pool_graph):
                for slot in range(POOL):
                    step(slot, feats[slot], keys[slot])
            for slot in range(POOL):  # one untimed replay each
                graphs[slot].replay()
            pool_graph.replay()
            sync_all()
        except Exception as e:  # capture is an optimisation: fall back to eager launches and say so
            graphs, pool_graph, graph_err = None, None, "%s: %s" % (type(e).__name__, e)
            torch.cuda.synchronize()

    # NVML is set up before the barrier (nvmlInit takes milliseconds and a different time on every rank)
    # and polled by rank 0 only (its calls take a driver-wide lock).
    sampler = ClockSampler(local)
    sync_all()
    if rank == 0 and os.environ.get("BENCH_NO_SAMPLER") != "1":
        sampler.start()

    def align_ranks(i=0):
        """One UNTIMED step after the barrier, N > 1 only: the ranks leave the host barrier up to milliseconds
        apart, and the first exchange would charge that skew to the timed region of the early ranks (seen as
        +15 us/step over 200 steps at N=8).  After one coupled step the streams are within microseconds."""
        if world > 1:
            step(i, feats[i % POOL], keys[i % POOL])

    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    align_ranks()
    t_start.record()
    if graphs is not None:
        for _ in range(args.steps // POOL):  # POOL steps per launch ...
            pool_graph.replay()
        for i in range(args.steps % POOL):   # ... and the remainder one step at a time: exactly K steps
            graphs[i].replay()
    else:
        for i in range(args.steps):
            step(i, feats[i % POOL], keys[i % POOL])
    t_end.record()
    sync_all()
    clocks = sampler.stop()
    ms_total = t_start.elapsed_time(t_end)

    # duration of the dominant kernel (EMA) with CUDA events on its stream, same step sequence, eager
    # launches (event records cannot be timed inside a captured graph)
    for i in range(min(args.steps, 100)):
        step(i, feats[i % POOL], keys[i % POOL], time_ema=True)
    sync_all()
    ema_ms = sum(a.elapsed_time(b) for a, b in ema_events) / len(ema_events)
    loss_val = float(out["loss"].item())
    assert int(status.item()) == 0, "device status word set: %d" % int(status.item())

    # ---- end-to-end timing: pinned host inputs, loss read back, one sync per step
    f_dev = torch.empty(B_PER_GPU, DIM, device=dev)
    k_dev = torch.empty(B_PER_GPU, DIM, device=dev)
    loss_h = torch.empty(2).pin_memory()  # two slots: step i's loss is read while step i+1 runs
    e2e_steps = args.steps

    h2d = torch.cuda.Stream(device=dev, priority=-1)
    copied = torch.cuda.Event()

    def e2e_step(i):
        # The momentum update does not depend on this step's inputs, so the host->device copies run on
        # a copy stream underneath it; the head (and, for N > 1, the key all_gather) waits for them.
        if not use_nccl:
            ema_part(k_dev, push=False)  # launched first: the GPU starts on it while the host enqueues the copies
        with torch.cuda.stream(h2d):
            k_dev.copy_(keys_h[i % POOL], non_blocking=True)
            f_dev.copy_(feats_h[i % POOL], non_blocking=True)
            if use_peer:
                xchg.push(k_dev)  # C3 right behind the copy, on the copy stream, still under the EMA
            copied.record()
        if use_nccl:
            ema_part(k_dev, after=copied)  # the key all_gather needs this step's keys
        torch.cuda.current_stream().wait_event(copied)
        r = head_part(f_dev, k_dev)
        loss_h[0:1].copy_(r["loss"], non_blocking=True)
        torch.cuda.current_stream().synchronize()  # also orders the next step's copies after this step's reads
        return float(loss_h[0])

    for i in range(min(args.warmup, 10)):
        e2e_step(i)
    sync_all()

    # the same end-to-end step as one graph per input slot: H2D copies (copy stream) || EMA -> head -> loss D2H
    e2e_graphs = None
    if graphs is not None:
        try:
            e2e_graphs = []
            for slot in range(POOL):
                g_ = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g_):
                    cur = torch.cuda.current_stream()
                    h2d.wait_stream(cur)  # fork: the copies depend on nothing in this step
                    with torch.cuda.stream(h2d):
                        k_dev.copy_(keys_h[slot], non_blocking=True)
                        f_dev.copy_(feats_h[slot], non_blocking=True)
                        if use_peer:
                            xchg.push(k_dev)
                    if use_nccl:
                        comm.wait_stream(h2d)  # the key all_gather needs this step's keys
                    ema_part(k_dev, push=False)  # runs beside the copies
                    cur.wait_stream(h2d)
                    r = head_part(f_dev, k_dev)
                    loss_h[slot & 1:(slot & 1) + 1].copy_(r["loss"], non_blocking=True)
                e2e_graphs.append(g_)
            for slot in range(POOL):
                e2e_graphs[slot].replay()
            sync_all()
        except Exception as e:
            e2e_graphs, graph_err = None, "e2e %s: %s" % (type(e).__name__, e)
            torch.cuda.synchronize()

    def e2e_graph_step(i):
        e2e_graphs[i % POOL].replay()
        torch.cuda.current_stream().synchronize()
        return float(loss_h[i & 1])

    # (a) strict: the host waits for the loss of step i before it enqueues step i+1
    run_e2e = e2e_graph_step if e2e_graphs is not None else e2e_step
    e_start, e_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        run_e2e(0)  # untimed: aligns the ranks after the host barrier (see align_ranks)
    e_start.record()
    for i in range(e2e_steps):
        run_e2e(i)
    e_end.record()
    sync_all()
    e2e_strict_ms = e_start.elapsed_time(e_end)
    e2e_ms, e2e_mode = e2e_strict_ms, "strict: host waits for step i's loss before enqueuing step i+1"

    # (b) one step in flight: the host enqueues step i+1 (its H2D copies included), then waits for and
    # reads the loss of step i -- every step still copies its inputs from pinned host memory and has
    # its loss read on the host, one host wait per step; the launch latency hides behind the GPU work.
    if e2e_graphs is not None and POOL % 2 == 0:
        done = [torch.cuda.Event(), torch.cuda.Event()]
        acc = 0.0
        sync_all()
        if world > 1:
            e2e_graph_step(0)  # untimed: aligns the ranks after the host barrier
        e_start.record()
        for i in range(e2e_steps):
            e2e_graphs[i % POOL].replay()
            done[i & 1].record()
            if i > 0:
                done[(i - 1) & 1].synchronize()
                acc += float(loss_h[(i - 1) & 1])
        done[(e2e_steps - 1) & 1].synchronize()
        acc += float(loss_h[(e2e_steps - 1) & 1])
        e_end.record()
        sync_all()
        assert acc == acc, "non-finite loss in the end-to-end run"
        e2e_ms = e_start.elapsed_time(e_end)
        e2e_mode = "one step in flight: step i+1 is enqueued before the host waits for and reads the loss of step i"

    # max over ranks
    if world > 1:
        t = torch.tensor([ms_total, e2e_ms, ema_ms, e2e_strict_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total, e2e_ms, ema_ms, e2e_strict_ms = (float(x) for x in t.tolist())

    ms_per_step = ms_total / args.steps
    value = n_gpus * B_PER_GPU / (ms_per_step * 1e-3)
    e2e_value = n_gpus * B_PER_GPU / (e2e_ms / e2e_steps * 1e-3)
    peak, peak_src = measured_peaks()
    ema_bytes = 12 * n_params
    achieved = ema_bytes / (ema_ms * 1e-3) / 1e9
    step_bytes = ema_bytes + 4 * QUEUE_LEN * DIM + 4 * B_PER_GPU * DIM * 3 + 8 * B_PER_GPU * DIM
    if not args.no_logits:
        step_bytes += 4 * B_PER_GPU * (QUEUE_LEN + 1)

    result = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "step_us": ms_per_step * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "batch_per_gpu": B_PER_GPU, "global_batch": n_gpus * B_PER_GPU,
                   "queue_len": QUEUE_LEN, "dim": DIM, "T": TEMP, "ema_tensors": len(online),
                   "ema_params": n_params, "logits_materialised": not args.no_logits,
                   "infonce_kernel": args.kernel, "cuda_graph": graphs is not None, "steps_per_graph_launch": POOL if graphs is not None else None,
                   "rank_alignment": "one untimed step between the barrier and the first timed event (N>1)" if world > 1 else None,
                   "e2e_cuda_graph": e2e_graphs is not None, "cuda_graph_error": graph_err,
                   "key_exchange": ("nvlink peer stores fused into the EMA launch, wait fused into the head launch"
                                    if use_peer else ("nccl all_gather on a side stream" if use_nccl else "none (1 GPU)")),
                   "key_exchange_verified_vs_nccl": exchange_verified, "key_exchange_note": exchange_note,
                   "parallelism": "dp%d (queue/EMA replicated, batch sharded; key exchange only)" % n_gpus,
                   "l2": "no explicit flush: one step streams %.0f MB (> 126 MB L2) so nothing survives between steps" % (step_bytes / 1e6)},
        "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms / e2e_steps,
                "h2d_bytes_per_step": 2 * 4 * B_PER_GPU * DIM, "d2h_bytes_per_step": 4,
                "mode": e2e_mode,
                "strict_sync": {"value": n_gpus * B_PER_GPU / (e2e_strict_ms / e2e_steps * 1e-3), "unit": UNIT,
                                "ms_per_step": e2e_strict_ms / e2e_steps},
                "note": "pinned host embeddings -> H2D on a copy stream (under the EMA) -> head+enqueue -> loss D2H to pinned memory, one host wait per step"},
        "gpu_launches": launches_per_step * args.steps,
        "clocks": clocks,
        "roofline": {"kernel": "ema_multi_tensor_kernel", "bound": "hbm", "achieved": achieved, "peak": peak,
                     "unit": "GB/s", "frac": achieved / peak, "traffic": EMA_DRAM_TRAFFIC,
                     "traffic_source": "ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch "
                                       "(profiles/r1_ema_ncu.md)",
                     "bytes_per_launch": ema_bytes, "us_per_launch": ema_ms * 1e3, "peak_source": peak_src},
        "step_roofline": {"bytes_per_step": step_bytes, "floor_us": step_bytes / (peak * 1e9) * 1e6,
                          "frac": (step_bytes / (peak * 1e9) * 1e3) / ms_per_step},
        "loss": loss_val,
    }
    if rank == 0 and n_gpus == 1 and not args.no_cpu_baseline:
        dt, done = run_cpu(50, 2, budget_s=15.0)
        result["cpu_baseline"] = {
            "value": B_PER_GPU / dt, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "ms_per_step": dt * 1e3,
            "sample": "%d full steps of the same workload on the CPU oracle (torch CPU, all threads) after 2 warm-up" % done}
    if rank == 0:
        json_out.write(json.dumps(result) + "\n")
        json_out.flush()
    if world > 1:
        # Leave without tearing NCCL down: destroy_process_group() blocks forever while captured graphs
        # still reference the communicator (seen at N=2), and the process is finished anyway.
        sync_all()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--kernel", default="auto", choices=["auto", "simt", "tc3x", "tc1x"])
    ap.add_argument("--no-logits", action="store_true", help="do not materialise the [B,K+1] logits tensor")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="eager launches instead of CUDA-graph replay")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="N > 1: cross-GPU key gather over NVLink peer memory (default) or NCCL all_gather")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return reference_arm(args)
    return gpu_arm(args)


if __name__ == "__main__":
    sys.exit(main())
