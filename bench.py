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

Both numbers go through the reference-facing API, the call tools/train.py:63-77 makes:
    contrastive_forward(model, cfg, inputs, index, time, epoch_exact) ; loss.backward()
on a `ContrastiveModel` whose two backbones are a stub registered in `_MODEL_TYPES`: it
OWNS the 164 Slow-R50 parameter tensors (so the EMA streams the real 36.1 M parameters)
and forwards the synthetic [B, D] embedding it is given (backbone compute is excluded).
`value`  : clips/s = N * 64 / step time, inputs already resident in HBM (CUDA-graph replay
           of the captured module step).
`e2e`    : the same call with pinned HOST embeddings: H2D of the step's inputs and D2H of
           the loss inside the timed region.  `e2e.strict_sync` (also at the top level as
           `e2e_strict`) waits for step i's loss before enqueuing step i+1.
`ops_level`: the kernel-only step (EMA[+push] -> head launch) driven through ops.* as in
           round 1, kept beside the module number.
`roofline`: the dominant kernel (the EMA, 93 % of the step's bytes): algorithmic
           bytes 12 B/param per launch / CUDA-event duration of that launch.
`cpu_baseline` / `--impl reference`: the CPU oracle port of the same step
           (torch CPU ops, all host threads), bounded sample.
"""
import argparse
import json
import math
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

B_PER_GPU, DIM, QUEUE_LEN, TEMP, MOMENTUM = 64, 128, 65536, 0.1, 0.999
POOL = 8  # distinct synthetic batches rotated through the steps
EMA_DRAM_TRAFFIC = 383_306_752  # bytes per launch of the EMA kernel measured by ncu (289.0 MB read + 94.3 MB written, profiles/r2_ema_ncu.md)
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
    """CPU arm: the oracle port of the step (torch CPU ops in the reference's own order).  The unmodified
    reference module cannot travel to the GPU box (it is never copied into this repo, and its imports
    need fvcore / pytorchvideo, absent from the image), so `kind` is "port": the port is pinned bit for
    bit to the unmodified reference by tests/golden (see oracle/contrastive_oracle.py)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    steps = min(args.steps, 1000)    # ~13 ms per step on 16 cores; the time budget below bounds the run either way
    warmup = min(args.warmup, 50)    # the same --warmup the GPU arm gets
    dt, done = run_cpu(steps, warmup, budget_s=120.0)
    val = B_PER_GPU / dt
    cores = torch.get_num_threads()
    sample = "%d full steps (EMA 36.1M params + head B=64,K=65536,D=128 fwd/bwd + enqueue) after %d warm-up" % (done, warmup)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": done, "warmup": warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "batch_per_gpu": B_PER_GPU, "global_batch": B_PER_GPU, "queue_len": QUEUE_LEN, "dim": DIM,
                   "T": TEMP, "ema_tensors": len(param_shapes()), "ema_params": sum(math.prod(s) for s in param_shapes()),
                   "note": "CPU oracle port (torch CPU ops = the reference's own op sequence, pinned to the "
                                                 "unmodified reference by tests/golden; the reference itself cannot travel to "
                                                 "the GPU box), rank 0 only"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))
    return 0


# ------------------------------------------------------------------------ GPU arm
class _Node:
    """Attribute tree standing in for the host application's fvcore CfgNode."""

    def __init__(self, **kw):
        self.__dict__.update(kw)


def head_cfg(world):
    """The cfg keys the contrastive path reads (configs/defaults.py:87-158), at the BASELINE configs[1] values."""
    return _Node(
        NUM_GPUS=world, NUM_SHARDS=1, SHARD_ID=0,
        MODEL=_Node(MODEL_NAME="ContrastiveModel", ARCH="stub_slow_r50"),
        BN=_Node(NORM_TYPE="batchnorm", NUM_SYNC_DEVICES=1),  # no cross-GPU sync BN -> shuffle BN is ON (the reference default)
        DATA=_Node(TRAIN_CROP_NUM_TEMPORAL=2, TRAIN_CROP_NUM_SPATIAL=1),
        SOLVER=_Node(MAX_EPOCH=200),
        TRAIN=_Node(BATCH_SIZE=B_PER_GPU * world),
        CONTRASTIVE=_Node(T=TEMP, DIM=DIM, LENGTH=239975, QUEUE_LEN=QUEUE_LEN, MOMENTUM=MOMENTUM, MOMENTUM_ANNEALING=False,
                          TYPE="moco", INTERP_MEMORY=False, MEM_TYPE="1d", LOCAL_SHUFFLE_BN=True,
                          MOCO_MULTI_VIEW_QUEUE=False, PREDICTOR_DEPTHS=[], SEQUENTIAL=False, SIMCLR_DIST_ON=True,
                          SWAV_QEUE_LEN=0, KNN_ON=False))


def register_stub_backbone():
    import torch.nn as nn
    from advise_video_ssl_b200 import contrastive as C

    class SlowR50Stub(nn.Module):
        """Owns the reference's Slow-R50 + MLP-head parameter list (164 tensors, 36,095,168 fp32) so that the
        momentum update streams the real thing; its forward hands the synthetic [B, D] embedding through
        (backbone forward/backward are out of scope and excluded)."""

        def __init__(self, cfg):
            super().__init__()
            g = torch.Generator().manual_seed(1234)
            self.weights = nn.ParameterList([nn.Parameter(torch.randn(s, generator=g) * 0.02) for s in param_shapes()])

        def forward(self, x):
            return x[0] if isinstance(x, (list, tuple)) else x

    C._MODEL_TYPES["stub_slow_r50"] = SlowR50Stub
    return C


class Timer:
    def __init__(self):
        self.a, self.b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def ms(self):
        return self.a.elapsed_time(self.b)


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
    impl = {"auto": _lib.IMPL_AUTO, "simt": _lib.IMPL_SIMT, "tc3x": _lib.IMPL_TC3X, "tc1x": _lib.IMPL_TC1X}[args.kernel]

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- synthetic data (seeded per rank), shared by the module path and the ops-level path
    g = torch.Generator().manual_seed(1000 + rank)
    stdv = 1.0 / (DIM / 3) ** 0.5
    gq = torch.Generator().manual_seed(7)
    queue0 = torch.rand(QUEUE_LEN, DIM, generator=gq).mul_(2 * stdv).add_(-stdv)
    feats_h = [torch.randn(B_PER_GPU, DIM, generator=g).pin_memory() for _ in range(POOL)]
    kfeat_h = [torch.randn(B_PER_GPU, DIM, generator=g).pin_memory() for _ in range(POOL)]  # raw key-encoder outputs

    # =========================================================== the module path (value, e2e)
    C = register_stub_backbone()
    cfg = head_cfg(world)
    torch.manual_seed(99)
    model = C.ContrastiveModel(cfg).to(dev).train()
    model.infonce_impl = impl
    model.materialize_logits = not args.no_logits
    if args.exchange == "nccl":
        model.enable_peer_exchange(False)
    with torch.no_grad():
        model.queue_x.copy_(queue0)
    n_params = sum(p.numel() for p in model.backbone_hist.parameters())
    n_tensors = len(list(model.backbone_hist.parameters()))
    index = torch.arange(B_PER_GPU, device=dev)
    time_in = torch.zeros(B_PER_GPU, 2, 1, device=dev)
    xq = [t.to(dev).requires_grad_(True) for t in feats_h]  # static leaves: one pair per input slot
    xk = [t.to(dev) for t in kfeat_h]
    last = {}

    def module_step(slot):
        """What tools/train.py does per iteration with the head (:63-77, :202-205): forward through the
        public entry point, then backward of the returned loss."""
        xq[slot].grad = None
        _, preds, loss, do_backward = C.contrastive_forward(model, cfg, [[xq[slot]], [xk[slot]]], index, time_in, 0.0)
        if do_backward:
            loss.backward()
        last["loss"], last["grad"], last["preds"] = loss, xq[slot].grad, preds
        return loss

    for i in range(args.warmup):
        module_step(i % POOL)
    sync_all()
    deferred_path = any(ex is not None for ex in model._peer_xchgs.values())
    torch.manual_seed(4242)  # every rank captures with its own CPU generator state; rank 0's draws decide (C2)

    graphs, pool_graph, graph_err = None, None, None
    losses_static = [None] * POOL
    if not args.no_graph:
        try:
            graphs = []
            for slot in range(POOL):
                g_ = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g_):
                    losses_static[slot] = module_step(slot)
                graphs.append(g_)
            # ... and one graph holding all POOL steps back to back: relaunching one graph is cheaper on the host
            # than alternating between POOL of them, which matters once 8 ranks submit work at the same time
            pool_graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(pool_graph):
                for slot in range(POOL):
                    module_step(slot)
            for slot in range(POOL):
                graphs[slot].replay()
            pool_graph.replay()
            sync_all()
        except Exception as e:  # capture is an optimisation: fall back to eager launches and say so
            graphs, pool_graph, graph_err = None, None, "%s: %s" % (type(e).__name__, e)
            torch.cuda.synchronize()

    def run_steps(n, one_step, many=None):
        if many is not None:
            for _ in range(n // POOL):
                many()
            for i in range(n % POOL):
                one_step(i)
        else:
            for i in range(n):
                one_step(i % POOL)

    # NVML is set up before the barrier (nvmlInit takes milliseconds and a different time on every rank)
    # and polled by rank 0 only (its calls take a driver-wide lock).
    sampler = ClockSampler(local)
    sync_all()
    if rank == 0 and os.environ.get("BENCH_NO_SAMPLER") != "1":
        sampler.start()

    replay_one = (lambda i: graphs[i].replay()) if graphs is not None else module_step
    replay_pool = pool_graph.replay if pool_graph is not None else None
    t_mod = Timer()
    if world > 1:
        # one UNTIMED coupled step after the barrier: the ranks leave the host barrier up to milliseconds apart
        # and the first exchange would charge that skew to the timed region of the early ranks
        replay_one(0)
    t_mod.a.record()
    run_steps(args.steps, replay_one, replay_pool)
    t_mod.b.record()
    sync_all()
    clocks = sampler.stop()
    ms_total = t_mod.ms()
    loss_val = float(last["loss"].item()) if graphs is None else float(losses_static[(args.steps - 1) % POOL].item())
    model.check_device_status()

    # ---- end-to-end through the same API: pinned host inputs in, loss out, every step.
    # The step's two views travel as ONE pinned host tensor [2, B, D] (one copy per step, like a collated batch)
    # into one of two device landing buffers on a copy stream -- the usual input prefetcher of a training loop --
    # so the copy of step i+1 can run while step i computes; the loss of every step is copied back to pinned
    # memory and read by the host.  The module call itself is the captured graph of `value`, reading the landing
    # buffer of its parity.
    both_h = [torch.stack([feats_h[s_], kfeat_h[s_]]).pin_memory() for s_ in range(POOL)]
    both_in = [torch.empty(2, B_PER_GPU, DIM, device=dev) for _ in range(2)]
    xq_in = [b_[0].detach().requires_grad_(True) for b_ in both_in]  # views of the landing buffers; leaf = query view
    xk_in = [b_[1] for b_ in both_in]
    loss_h = torch.empty(2).pin_memory()  # two slots: step i's loss is read while step i+1 runs
    copy_stream = torch.cuda.Stream(device=dev, priority=-1)
    landed = [torch.cuda.Event(), torch.cuda.Event()]
    consumed = [torch.cuda.Event(), torch.cuda.Event()]
    e2e_loss = [None] * POOL

    def e2e_compute(par):
        xq_in[par].grad = None
        _, _, loss, do_backward = C.contrastive_forward(model, cfg, [[xq_in[par]], [xk_in[par]]], index, time_in, 0.0)
        if do_backward:
            loss.backward()
        return loss.detach().reshape(1)

    e2e_graphs = None

    def e2e_enqueue(i):
        """Everything the host submits for step i: H2D of its inputs, the module step, D2H of its loss."""
        slot, par = i % POOL, i & 1
        main = torch.cuda.current_stream()
        copy_stream.wait_event(consumed[par])  # the step that last read this landing buffer is done with it
        with torch.cuda.stream(copy_stream):
            both_in[par].copy_(both_h[slot], non_blocking=True)
            landed[par].record()
        main.wait_event(landed[par])
        if e2e_graphs is not None:
            e2e_graphs[slot].replay()
            loss_dev = e2e_loss[slot]
        else:
            loss_dev = e2e_compute(par)
        consumed[par].record()
        loss_h[par:par + 1].copy_(loss_dev, non_blocking=True)

    def e2e_strict_step(i):
        e2e_enqueue(i)
        torch.cuda.current_stream().synchronize()
        return float(loss_h[i & 1])

    for i in range(min(args.warmup, 10)):
        e2e_strict_step(i)
    sync_all()
    if graphs is not None:
        try:
            cap = []
            for slot in range(POOL):
                g_ = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g_):
                    e2e_loss[slot] = e2e_compute(slot & 1)
                cap.append(g_)
            e2e_graphs = cap
            for i in range(POOL):
                e2e_strict_step(i)
            sync_all()
        except Exception as e:
            e2e_graphs, graph_err = None, "e2e %s: %s" % (type(e).__name__, e)
            torch.cuda.synchronize()

    # (a) strict: the host waits for the loss of step i before it submits anything of step i+1.  With nothing to
    # overlap, the fewest host calls win: the whole step -- H2D copy, module call, D2H of the loss -- is ONE graph.
    strict_graphs = None
    if e2e_graphs is not None:
        try:
            cap = []
            for slot in range(POOL):
                g_ = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g_):
                    both_in[slot & 1].copy_(both_h[slot], non_blocking=True)
                    loss_h[slot & 1:(slot & 1) + 1].copy_(e2e_compute(slot & 1), non_blocking=True)
                cap.append(g_)
            strict_graphs = cap
        except Exception as e:
            graph_err = "e2e strict %s: %s" % (type(e).__name__, e)
            torch.cuda.synchronize()

    def strict_step(i):
        if strict_graphs is None:
            return e2e_strict_step(i)
        strict_graphs[i % POOL].replay()
        torch.cuda.current_stream().synchronize()
        return float(loss_h[i & 1])

    for i in range(POOL):
        strict_step(i)
    sync_all()
    t_e2e = Timer()
    if world > 1:
        strict_step(0)  # untimed: aligns the ranks after the host barrier
    t_e2e.a.record()
    for i in range(args.steps):
        strict_step(i)
    t_e2e.b.record()
    sync_all()
    e2e_strict_ms = t_e2e.ms()
    e2e_ms, e2e_mode = e2e_strict_ms, "strict: host waits for step i's loss before submitting step i+1"

    # (b) one step in flight: the host submits step i+1 (its H2D copy included), then waits for and reads the loss of
    # step i -- every step still copies its inputs from pinned host memory and has its loss read on the host, one
    # host wait per step; launch latency and the input copy hide behind the previous step's GPU work.
    done = [torch.cuda.Event(), torch.cuda.Event()]
    acc = 0.0
    sync_all()
    if world > 1:
        e2e_strict_step(0)
    t_e2e.a.record()
    for i in range(args.steps):
        e2e_enqueue(i)
        done[i & 1].record()
        if i > 0:
            done[(i - 1) & 1].synchronize()
            acc += float(loss_h[(i - 1) & 1])
    done[(args.steps - 1) & 1].synchronize()
    acc += float(loss_h[(args.steps - 1) & 1])
    t_e2e.b.record()
    sync_all()
    assert acc == acc, "non-finite loss in the end-to-end run"
    e2e_ms = t_e2e.ms()
    e2e_mode = "one step in flight: step i+1 is submitted before the host waits for and reads the loss of step i"
    model.check_device_status()

    # ---- C9 evidence in the line itself: after all those steps every rank must hold the same queue / ptr / iter
    queue_consistent = None
    if world > 1:
        probe = torch.stack([model.queue_x.view(torch.int32).to(torch.int64).sum(), model.ptr[0], model.iter[0],
                             (model.queue_x.view(torch.int32).to(torch.int64) * torch.arange(1, DIM + 1, device=dev)).sum()])
        allp = [torch.empty_like(probe) for _ in range(world)]
        dist.all_gather(allp, probe)
        queue_consistent = bool(all(torch.equal(allp[0], t_) for t_ in allp))

    # ================================================ the ops-level step (kernel-only, as in round 1)
    # also yields the duration of the dominant kernel (the EMA): CUDA events around its launch inside the eager
    # two-launch sequence, where the host keeps ahead of the GPU (event records cannot be timed inside a captured
    # graph, and the eager MODULE step is host-bound, so events there would time launch gaps, not the kernel)
    ops_ms, ema_ms, beside = ops_level(args, ops, dist, dev, world, rank, impl, queue0, feats_h, kfeat_h, model, sync_all)
    ops_note = None if ops_ms is not None else "--no-ops-level"
    two_launch = beside is not None

    # max over ranks
    if world > 1:
        t = torch.tensor([ms_total, e2e_ms, ema_ms, e2e_strict_ms, ops_ms if ops_ms is not None else -1.0], device=dev,
                         dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total, e2e_ms, ema_ms, e2e_strict_ms, om = (float(x) for x in t.tolist())
        ops_ms = om if om >= 0 else None

    ms_per_step = ms_total / args.steps
    clips = n_gpus * B_PER_GPU
    value = clips / (ms_per_step * 1e-3)
    peak, peak_src = measured_peaks()
    ema_bytes = 12 * n_params
    achieved = ema_bytes / (ema_ms * 1e-3) / 1e9
    step_bytes = ema_bytes + 4 * QUEUE_LEN * DIM + 4 * B_PER_GPU * DIM * 3 + 8 * B_PER_GPU * DIM
    if not args.no_logits:
        step_bytes += 4 * B_PER_GPU * (QUEUE_LEN + 1)
    floor_us = step_bytes / (peak * 1e9) * 1e6
    # own kernels per step: EMA, head(+Normalize of the keys, un-shuffle, enqueue) on one GPU; across GPUs one more
    # launch normalises the keys and stores them into every rank's exchange buffer (or: Normalize, then NCCL)
    own_launches = 2 if (world == 1 or deferred_path) else 3
    if two_launch:
        own_launches += 2  # the head is sweep + merge of its partials (own stream, beside the EMA) + finish
    if world > 1:
        own_launches += 1  # the NVLink scatter of the key-encoder input rows (side stream, under the EMA)

    def per_step(ms):
        return {"value": clips / (ms / args.steps * 1e-3), "unit": UNIT, "ms_per_step": ms / args.steps}

    result = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "step_us": ms_per_step * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "batch_per_gpu": B_PER_GPU, "global_batch": clips,
                   "queue_len": QUEUE_LEN, "dim": DIM, "T": TEMP, "ema_tensors": n_tensors, "ema_params": n_params,
                   "api": "contrastive_forward(model, cfg, inputs, index, time, epoch) + loss.backward() on ContrastiveModel "
                          "(moco, SEQUENTIAL off, shuffle BN on, KNN_ON off, queue_mode=%s); stub backbone forwards the "
                          "synthetic embedding" % model.queue_mode,
                   "logits_materialised": not args.no_logits, "infonce_kernel": args.kernel,
                   "cuda_graph": graphs is not None, "steps_per_graph_launch": POOL if graphs is not None else None,
                   "shuffle_perm": "drawn per captured step (frozen in its graph); %d input slots" % POOL,
                   "rank_alignment": "one untimed step between the barrier and the first timed event (N>1)" if world > 1 else None,
                   "e2e_cuda_graph": e2e_graphs is not None, "cuda_graph_error": graph_err,
                   "key_exchange": ("inside the head launch: one extra CTA normalises this rank's key rows and stores them into "
                                    "every rank's buffer over NVLink while the others sweep the queue; the merge waits, "
                                    "un-shuffles by index and enqueues rank 0's rows" if deferred_path and world > 1 else
                                    ("nccl all_gather of the normalised keys; the head launch un-shuffles by index and enqueues "
                                     "rank 0's rows" if world > 1 else
                                     "none (one GPU): the head launch normalises the raw key rows and un-shuffles by index")),
                   "clip_shuffle": ("NVLink row scatter (each row written once into its final position on the destination rank) "
                                    "on a side stream under the EMA" if deferred_path and world > 1 else
                                    ("NCCL all-to-all on a side stream under the EMA" if world > 1
                                     else "local row gather on a side stream under the EMA")),
                   "queue_ptr_iter_identical_across_ranks": queue_consistent,
                   "parallelism": "dp%d (queue/EMA replicated, batch sharded; exchanges: shuffle rows, keys)" % n_gpus,
                   "l2": "no explicit flush: one step streams %.0f MB (> 126 MB L2) so nothing survives between steps" % (step_bytes / 1e6)},
        "e2e": dict(per_step(e2e_ms), h2d_bytes_per_step=2 * 4 * B_PER_GPU * DIM, d2h_bytes_per_step=4, mode=e2e_mode,
                    strict_sync=per_step(e2e_strict_ms),
                    note="pinned host embeddings -> H2D on a copy stream into a double-buffered landing zone -> "
                         "contrastive_forward + backward -> loss D2H to pinned memory, one host wait per step"),
        "e2e_strict": per_step(e2e_strict_ms),
        "ops_level": ({"ms_per_step": ops_ms / args.steps, "step_us": ops_ms / args.steps * 1e3,
                       "roofline_frac": floor_us / (ops_ms / args.steps * 1e3),
                       "what": "EMA[+push] launch -> head launch driven through ops.* (no module, no autograd), CUDA-graph replay"}
                      if ops_ms is not None else {"skipped": ops_note or "--no-ops-level"}),
        "gpu_launches": own_launches * args.steps,
        "gpu_launches_note": ("own kernels per step: head sweep + merge of its partials (own stream, beside the EMA), EMA, "
                              "head finish(+key Normalize, [N>1: key push over NVLink + wait,] un-shuffle, enqueue)"
                              if two_launch else
                              "own kernels per step: EMA, head(+key Normalize, [N>1: key push over NVLink + wait,] un-shuffle, "
                              "enqueue)") + " [, N>1: row scatter of the shuffle, side stream]; plus torch's row gather of the "
                             "shuffle on one GPU and autograd's ones-fill and elementwise multiply in backward",
        "clocks": clocks,
        "roofline": {"kernel": "ema_multi_tensor_kernel", "bound": "hbm", "achieved": achieved, "peak": peak,
                     "unit": "GB/s", "frac": achieved / peak, "traffic": EMA_DRAM_TRAFFIC,
                     "traffic_source": "ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch "
                                       "(profiles/r2_ema_ncu.md)",
                     "bytes_per_launch": ema_bytes, "us_per_launch": ema_ms * 1e3, "peak_source": peak_src},
        "step_roofline": {"bytes_per_step": step_bytes, "floor_us": floor_us, "frac": floor_us / (ms_per_step * 1e3)},
        "roofline_in_step": ({
            "what": "the module step runs the head's sweep of the queue (tcgen05, %d CTAs, own stream) BESIDE the EMA kernel: "
                    "EMA launch duration with the sweep sharing HBM, CUDA events in an eager ops-level sequence" % beside["ctas"],
            "ema_us_beside_sweep": beside["ema_ms"] * 1e3, "ema_us_alone": ema_ms * 1e3,
            "bytes": ema_bytes + beside["sweep_bytes"], "unit": "GB/s", "peak": peak,
            "achieved": (ema_bytes + beside["sweep_bytes"]) / (beside["ema_ms"] * 1e-3) / 1e9,
            "frac": (ema_bytes + beside["sweep_bytes"]) / (beside["ema_ms"] * 1e-3) / 1e9 / peak} if two_launch else None),
        "loss": loss_val,
    }
    if rank == 0 and n_gpus == 1 and not args.no_cpu_baseline:
        dt, done_n = run_cpu(50, 2, budget_s=15.0)
        result["cpu_baseline"] = {
            "value": B_PER_GPU / dt, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "ms_per_step": dt * 1e3,
            "sample": "%d full steps of the same workload on the CPU oracle (torch CPU, all threads) after 2 warm-up" % done_n}
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


def _setup_simclr(C, cfg, dev, rank, world):
    """BASELINE configs[2]: SimCLR, global all_gather negatives, 512 clips per GPU (total batch 4096 at 8 GPUs), dim 256."""
    import torch.nn as nn
    B, D, T = 512, 256, 0.1

    class Identity(nn.Module):
        def __init__(self, cfg):
            super().__init__()
            self.dummy = nn.Parameter(torch.zeros(4))

        def forward(self, x):
            return x[0] if isinstance(x, (list, tuple)) else x

    C._MODEL_TYPES["identity_embed"] = Identity
    cfg.MODEL.ARCH = "identity_embed"
    cfg.CONTRASTIVE.TYPE, cfg.CONTRASTIVE.DIM, cfg.CONTRASTIVE.T = "simclr", D, T
    cfg.CONTRASTIVE.QUEUE_LEN = 64
    cfg.TRAIN.BATCH_SIZE = B * world
    model = C.ContrastiveModel(cfg).to(dev).train()
    g = torch.Generator().manual_seed(2000 + rank)
    f1 = [torch.randn(B, D, generator=g).to(dev).requires_grad_(True) for _ in range(POOL)]
    f2 = [torch.randn(B, D, generator=g).to(dev).requires_grad_(True) for _ in range(POOL)]
    index = torch.arange(B, device=dev)

    def step(slot):
        f1[slot].grad, f2[slot].grad = None, None
        _, loss = model([[f1[slot]], [f2[slot]]], index, None, 0.0)
        loss.backward()
        return loss

    N = B * world
    flops = 3 * 2 * (2 * B) * (2 * N) * D  # this rank's rows: S for the row sums, S again and P.V for the gradient
    return dict(
        B=B, step=step, dtype="f32 (kind::f16 operands on unit rows, fp32 accumulate)",
        workload="configs[2] head: SimCLR NT-Xent, global all_gather negatives, 512 clips/GPU (total batch %d), dim 256, "
                 "T 0.1; l2norm + all_gather + row sums + all_gather + gradient + finalise; backbone excluded" % N,
        api="ContrastiveModel(simclr).forward + loss.backward()",
        collectives="2 gathers per step (rows, row sums) inside the timed region, NVLink peer stores on one box" if world > 1 else "none",
        parallelism="dp%d, rows sharded: each rank computes its 2B rows against all 2N columns" % world,
        launches=6, launches_note="l2norm, prepare, rowsum, sum_z, grad, finish per step (+ torch cat / mul)",
        roofline=lambda per_ms, peaks: {
            "kernel": "ntxent step (per-rank useful flops 3*2*(2B)(2N)D)", "bound": "tensor",
            "achieved": flops / (per_ms * 1e-3) / 1e12, "peak": float(peaks.get("bf16_tflops", 1590.0)), "unit": "TFLOP/s",
            "frac": flops / (per_ms * 1e-3) / 1e12 / float(peaks.get("bf16_tflops", 1590.0)), "traffic": None,
            "peak_source": "MEASURED_PEAKS.json bf16_tflops (kind::f16 runs at the bf16 rate)"})


def _setup_byol(C, cfg, dev, rank, world):
    """BASELINE configs[3]: BYOL, momentum EMA over the Slow-R50 + projection + predictor list (169 tensors, 37.4 M fp32,
    dim 256, PREDICTOR_DEPTHS [2]) + the symmetric predictor loss over 2 views, batch 64 per GPU."""
    import torch.nn as nn
    B, D, T = 64, 256, 0.1
    shapes = param_shapes("slow_r50_byol_dim256_pred2")

    class ByolStub(nn.Module):
        """Owns the reference's parameter list so that the momentum update streams the real thing; hands the synthetic
        projection / prediction pair through as [feat, pred] (models/head_helper.py:232-235)."""

        def __init__(self, cfg):
            super().__init__()
            g = torch.Generator().manual_seed(1234)
            self.weights = nn.ParameterList([nn.Parameter(torch.randn(s, generator=g) * 0.02) for s in shapes])

        def forward(self, x):
            return [x[0], x[1]]

    C._MODEL_TYPES["stub_byol"] = ByolStub
    cfg.MODEL.ARCH = "stub_byol"
    cfg.CONTRASTIVE.TYPE, cfg.CONTRASTIVE.DIM, cfg.CONTRASTIVE.T = "byol", D, T
    cfg.CONTRASTIVE.PREDICTOR_DEPTHS, cfg.CONTRASTIVE.MOMENTUM = [2], 0.996
    cfg.TRAIN.BATCH_SIZE = B * world
    model = C.ContrastiveModel(cfg).to(dev).train()
    g = torch.Generator().manual_seed(3000 + rank)
    views = [[[torch.randn(B, D, generator=g).to(dev), torch.randn(B, D, generator=g).to(dev).requires_grad_(True)]
              for _ in range(2)] for _ in range(POOL)]
    index = torch.arange(B, device=dev)

    def step(slot):
        for v in views[slot]:
            v[1].grad = None
        _, loss = model(views[slot], index, None, 0.0)
        loss.backward()
        return loss

    n_params = sum(math.prod(s) for s in shapes)
    step_bytes = 12 * n_params + 2 * 4 * B * D * 3  # EMA + the two sim_loss pairs
    return dict(
        B=B, step=step, dtype="f32",
        workload="configs[3] head: BYOL, EMA(%d tensors, %.1fM fp32) + Normalize of the keys + symmetric predictor loss over "
                 "2 views (2 x sim_loss fwd/bwd), batch %d/GPU, dim %d; backbone excluded" % (len(shapes), n_params / 1e6, B, D),
        api="ContrastiveModel(byol).forward + loss.backward()", collectives="none (BYOL has no exchange in the head)",
        parallelism="dp%d (EMA replicated, batch sharded)" % world,
        launches=5, launches_note="EMA, 2 x Normalize(keys), 2 x sim_loss per step (+ torch cat / split of the batched key pass, autograd adds)",
        roofline=lambda per_ms, peaks: {
            "kernel": "step (EMA-dominated: 12 B per parameter)", "bound": "hbm", "achieved": step_bytes / (per_ms * 1e-3) / 1e9,
            "peak": float(peaks.get("hbm_gbs", 6650.0)), "unit": "GB/s",
            "frac": step_bytes / (per_ms * 1e-3) / 1e9 / float(peaks.get("hbm_gbs", 6650.0)), "traffic": None,
            "bytes_per_step": step_bytes, "peak_source": "MEASURED_PEAKS.json hbm_gbs"})


def _setup_swav(C, cfg, dev, rank, world):
    """BASELINE configs[4]: SwAV multi-crop (2 + 4 crops), 3000 prototypes, 3 Sinkhorn iterations, batch 256 per GPU."""
    import torch.nn as nn
    B, D, T, P, crops = 256, 128, 0.1, 3000, 6

    class Identity(nn.Module):
        def __init__(self, cfg):
            super().__init__()
            self.dummy = nn.Parameter(torch.zeros(4))

        def forward(self, x):
            return x[0] if isinstance(x, (list, tuple)) else x

    C._MODEL_TYPES["identity_embed"] = Identity
    cfg.MODEL.ARCH = "identity_embed"
    cfg.CONTRASTIVE.TYPE, cfg.CONTRASTIVE.DIM, cfg.CONTRASTIVE.T = "swav", D, T
    cfg.CONTRASTIVE.SWAV_NUM_PROTOTYPES = P
    cfg.CONTRASTIVE.QUEUE_LEN = 64
    cfg.TRAIN.BATCH_SIZE = B * world
    model = C.ContrastiveModel(cfg).to(dev).train()
    g = torch.Generator().manual_seed(4000 + rank)
    embs = [[torch.randn(B, D, generator=g).to(dev).requires_grad_(True) for _ in range(crops)] for _ in range(POOL)]
    index = torch.arange(B, device=dev)

    def step(slot):
        for e in embs[slot]:
            e.grad = None
        model.swav_prototypes.weight.grad = None
        _, loss = model([[e] for e in embs[slot]], index, None, 0.0)
        loss.backward()
        return loss

    # the score rows cross HBM for the loss kernel (read + gradient write), the codes for Sinkhorn
    step_bytes = 4 * P * (crops * B * 2 + 2 * B) + 2 * 8 * P * B
    return dict(
        B=B, step=step, dtype="f32",
        workload="configs[4] head: SwAV, %d crops x %d clips/GPU, %d prototypes, eps 0.05, 3 Sinkhorn iterations, T %.1f: prototype "
                 "renorm + per-crop Normalize + scores (library GEMM) + 2 x Sinkhorn + swapped-prediction CE fwd/bwd + the "
                 "gradients back to embeddings and prototypes; backbone excluded" % (crops, B, P, T),
        api="ContrastiveModel(swav).forward + loss.backward()", collectives="none on one box (NUM_SHARDS = 1)",
        parallelism="dp%d (prototypes replicated, batch sharded)" % world,
        launches=3 + crops * 2, launches_note="prototype renorm, 6 x Normalize fwd (+ 6 x bwd in backward), 2 x Sinkhorn, swapped CE "
                                               "per step (+ cuBLAS score GEMMs and their backward, torch cat)",
        roofline=lambda per_ms, peaks: {
            "kernel": "step (K10 + K11 algorithmic bytes; the step is launch- and latency-bound: ~40 kernels)", "bound": "hbm",
            "achieved": step_bytes / (per_ms * 1e-3) / 1e9, "peak": float(peaks.get("hbm_gbs", 6650.0)), "unit": "GB/s",
            "frac": step_bytes / (per_ms * 1e-3) / 1e9 / float(peaks.get("hbm_gbs", 6650.0)), "traffic": None,
            "bytes_per_step": step_bytes, "peak_source": "MEASURED_PEAKS.json hbm_gbs"})


def extra_arm(args):
    """--workload simclr | byol | swav: BASELINE configs[2..4] through the same module API as the headline
    (ContrastiveModel.forward + loss.backward(), exchanges inside the timed region, CUDA-graph replay).  Not the headline
    (that is the MoCo step); one JSON line of the same shape."""
    import torch.distributed as dist
    from advise_video_ssl_b200 import contrastive as C

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the contrastive hot path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    json_out = sys.stdout
    if world > 1:
        sys.stdout.flush()
        json_out = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=dev)
    w = {"simclr": _setup_simclr, "byol": _setup_byol, "swav": _setup_swav}[args.workload](C, head_cfg(world), dev, rank, world)
    step, B = w["step"], w["B"]

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for i in range(args.warmup):
        step(i % POOL)
    sync_all()
    graph, graph_err, loss_t = None, None, None
    if not args.no_graph:
        try:
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                for slot in range(POOL):
                    loss_t = step(slot)
            graph.replay()
            sync_all()
        except Exception as e:  # noqa: BLE001
            graph, graph_err = None, "%s: %s" % (type(e).__name__, e)
            torch.cuda.synchronize()
    sampler = ClockSampler(local)
    sync_all()
    if rank == 0:
        sampler.start()
    n_launch = (args.steps + POOL - 1) // POOL  # whole pools: steps rounded up to a multiple of POOL
    steps = n_launch * POOL
    t = Timer()
    if world > 1 and graph is not None:
        graph.replay()
    t.a.record()
    for i in range(n_launch):
        if graph is not None:
            graph.replay()
        else:
            for slot in range(POOL):
                loss_t = step(slot)
    t.b.record()
    sync_all()
    clocks = sampler.stop()
    ms = t.ms()
    if world > 1:
        tt = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt.item())
    per = ms / steps
    peaks = {}
    pth = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pth):
        with open(pth) as f:
            peaks = json.load(f)
    if rank == 0:
        json_out.write(json.dumps({
            "metric": METRIC, "value": world * B / (per * 1e-3), "unit": UNIT, "n_gpus": world, "steps": steps,
            "warmup": args.warmup, "ms_per_step": per, "step_us": per * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": w["dtype"], "data": "synthetic",
            "config": {"workload": w["workload"], "api": w["api"], "cuda_graph": graph is not None,
                       "cuda_graph_error": graph_err, "collectives": w["collectives"], "parallelism": w["parallelism"]},
            "gpu_launches": w["launches"] * steps, "gpu_launches_note": w["launches_note"],
            "clocks": clocks, "roofline": w["roofline"](per, peaks),
            "loss": float(loss_t.item()) if loss_t is not None else None}) + "\n")
        json_out.flush()
    if world > 1:
        sync_all()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)
    return 0


def ops_level(args, ops, dist, dev, world, rank, impl, queue0, feats_h, kfeat_h, model, sync_all):
    """The two-launch kernel-only step of round 1, for comparison: EMA (+ the key push riding in the same
    launch) -> head (+ wait + enqueue).  Keys are pre-normalised device tensors, no autograd, no shuffle."""
    online, hist = model._ema_lists()
    plan = ops.EmaPlan(online, hist)
    queue = queue0.to(dev)
    feats = [t.to(dev) for t in feats_h]
    keys = [torch.nn.functional.normalize(t).to(dev) for t in kfeat_h]
    it = torch.ones(1, dtype=torch.int64, device=dev)
    ptr = torch.zeros(1, dtype=torch.int64, device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    xchg = None
    if world > 1 and args.exchange == "peer" and args.kernel != "simt":
        xchg = ops.PeerExchange(B_PER_GPU, DIM)
    gathered = torch.empty(world * B_PER_GPU, DIM, device=dev) if (world > 1 and xchg is None) else None
    comm = torch.cuda.Stream(device=dev, priority=-1) if gathered is not None else None
    gperm = torch.Generator().manual_seed(31337)  # the same un-shuffle table on every rank
    restore = torch.argsort(torch.randperm(world * B_PER_GPU, generator=gperm)).view(world, B_PER_GPU).to(dev)
    out = {}
    ws = torch.zeros(ops.moco_infonce_workspace_bytes(B_PER_GPU, DIM, QUEUE_LEN, 1), dtype=torch.uint8, device=dev)

    def step(slot, ema_timers=None):
        f, k = feats[slot], keys[slot]
        if gathered is not None:
            comm.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(comm):
                dist.all_gather_into_tensor(gathered, k)
        if ema_timers is not None:
            tm = Timer()
            tm.a.record()
        plan.run(MOMENTUM, it, bump_iter=True, first_iter=False, push=(xchg, k) if xchg is not None else None)
        if ema_timers is not None:
            tm.b.record()
            ema_timers.append(tm)
        if xchg is not None:
            r = ops.moco_infonce(f, None, queue, TEMP, want_logits=not args.no_logits, impl=impl, out=out,
                                 enqueue=(ptr, status), workspace=ws, peer=xchg, peer_row_idx=restore[rank].contiguous(),
                                 enq_row_idx=restore[0].contiguous())
        else:
            if gathered is not None:
                torch.cuda.current_stream().wait_stream(comm)
                k = gathered[rank * B_PER_GPU:(rank + 1) * B_PER_GPU]
            r = ops.moco_infonce(f, [k], queue, TEMP, want_logits=not args.no_logits, impl=impl, out=out,
                                 enqueue=(ptr, status), workspace=ws)
        if not out:
            out.update(r)

    for i in range(max(3, min(args.warmup, 10))):
        step(i % POOL)
    sync_all()
    ema_t = []
    for i in range(min(args.steps, 100)):
        step(i % POOL, ema_t)
    sync_all()
    ema_ms = sum(t.ms() for t in ema_t) / len(ema_t)
    # ---- the EMA kernel as the MODULE step runs it: beside the head's sweep (two-launch head), same eager timing
    beside = None
    ctas = model.sweep_ctas if model.sweep_ctas is not None else model._sweep_ctas_beside_ema(B_PER_GPU, dev)
    if ctas is not None and model.overlap_sweep and args.kernel != "simt":
        side = torch.cuda.Stream(device=dev, priority=-1)
        lg = None if args.no_logits else torch.empty(B_PER_GPU, QUEUE_LEN + 1, device=dev)

        def step2(slot, timers):
            f, k = feats[slot], keys[slot]
            main = torch.cuda.current_stream()
            side.wait_stream(main)
            with torch.cuda.stream(side):
                ops.moco_infonce_sweep(f, queue, TEMP, ws, logits=lg, impl=impl, sweep_ctas=ctas)
            tm = Timer()
            tm.a.record()
            plan.run(MOMENTUM, it, bump_iter=True, first_iter=False)
            tm.b.record()
            timers.append(tm)
            main.wait_stream(side)
            ops.moco_infonce(f, [k], queue, TEMP, want_logits=lg is not None, impl=impl,
                             out={"logits": lg} if lg is not None else None, enqueue=(ptr, status), workspace=ws, swept=ctas)

        for i in range(5):
            step2(i % POOL, [])
        sync_all()
        t2 = []
        for i in range(min(args.steps, 100)):
            step2(i % POOL, t2)
        sync_all()
        ptr.zero_()
        beside = {"ctas": ctas, "ema_ms": sum(t.ms() for t in t2) / len(t2),
                  "sweep_bytes": 4 * QUEUE_LEN * DIM + (0 if lg is None else 4 * B_PER_GPU * QUEUE_LEN)}
    if args.no_ops_level:
        return None, ema_ms, beside
    one, many = step, None
    if not args.no_graph:
        pool_graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(pool_graph):
            for slot in range(POOL):
                step(slot)
        singles = []
        for slot in range(POOL):
            g_ = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g_):
                step(slot)
            singles.append(g_)
        pool_graph.replay()
        sync_all()
        one, many = (lambda i: singles[i].replay()), pool_graph.replay
    t = Timer()
    if world > 1:
        one(0)
    t.a.record()
    if many is not None:
        for _ in range(args.steps // POOL):
            many()
        for i in range(args.steps % POOL):
            one(i)
    else:
        for i in range(args.steps):
            one(i % POOL)
    t.b.record()
    sync_all()
    assert int(status.item()) == 0, "device status word set in the ops-level run: %d" % int(status.item())
    return t.ms(), ema_ms, beside


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--kernel", default="auto", choices=["auto", "simt", "tc3x", "tc1x"])
    ap.add_argument("--no-logits", action="store_true", help="do not materialise the [B,K+1] logits tensor")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ops-level", action="store_true", help="skip the secondary kernel-only (ops.*) measurement")
    ap.add_argument("--no-graph", action="store_true", help="eager launches instead of CUDA-graph replay")
    ap.add_argument("--workload", default="moco", choices=["moco", "simclr", "byol", "swav"],
                    help="moco = BASELINE configs[1] (the headline); simclr / byol / swav = configs[2] / [3] / [4], extra lines")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="N > 1: cross-GPU key gather over NVLink peer memory (default) or NCCL all_gather")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return reference_arm(args)
    if args.workload != "moco":
        return extra_arm(args)
    return gpu_arm(args)


if __name__ == "__main__":
    sys.exit(main())
