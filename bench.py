#!/usr/bin/env python
"""Benchmark of the DSTAGNN block-stack training step (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--mode train|infer]
                    [--workload pems07|pems04|pems08|pems03|gambia|syn4096] [--precision tf32x3|fp32|tf32]
                    [--batch B] [--sweep B1,B2,...]

ours:      N ranks (torchrun for N>1), each a full replica with B_local samples; a step is
           zero-grad -> forward -> SmoothL1 -> backward (-> gradient all-reduce) -> Adam.
           `value`  = samples/s with the batch resident in HBM (CUDA events, max over ranks);
           `e2e`    = the same step through the public module API with HOST (pinned) inputs,
                      H2D copies and a D2H read of the loss inside the timed region.
reference: the reference's own CPU path on the box's host cores, rank 0 only: the UNMODIFIED
           model/DSTAGNN_my.py staged in oracle/_ref by oracle/build_ref.py (`kind: "reference"`), or -- when it is
           not staged -- the CPU oracle port with the reference's op order (`kind: "port"`).  Batch = the batch of
           the workload's own configurations/*.conf (12 for PEMS07), reduced only if the host RAM cannot hold the
           reference's autograd saves (stated in `sample`).
--mode infer: eval() / no_grad() forward only (BASELINE config 2: PEMS08 batch sweep with --sweep).
One JSON line is printed by rank 0 (one per batch size with --sweep).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

WORKLOADS = {
    # name: N, T, K, H, D, nb_block, per-GPU batch, P   
    # conf_B: batch_size of the reference's configurations/<name>_dstagnn.conf (the CPU arms run it)
    "pems07": dict(N=883, T=12, K=3, H=4, D=512, nb=4, B=64, P=12, conf_B=12),
    "pems04": dict(N=307, T=12, K=3, H=3, D=512, nb=4, B=32, P=12, conf_B=32),
    "pems08": dict(N=170, T=12, K=3, H=3, D=512, nb=4, B=64, P=12, conf_B=64),
    "pems03": dict(N=358, T=12, K=3, H=3, D=512, nb=4, B=32, P=12, conf_B=32),
    "gambia": dict(N=2139, T=144, K=2, H=2, D=64, nb=2, B=4, P=12, conf_B=4),
    "syn4096": dict(N=4096, T=12, K=3, H=4, D=512, nb=4, B=16, P=12, conf_B=2),
    "tiny": dict(N=64, T=12, K=3, H=2, D=32, nb=2, B=4, P=12, conf_B=4),
}


def metric_name(args):
    what = "train samples/sec (fwd+bwd)" if args.mode == "train" else "inference samples/sec (eval, no_grad)"
    return f"{what} {'PEMS07' if args.workload == 'pems07' else args.workload} shape"


def synthetic_graph(N, seed=0, density=0.01):
    """STAG/STRG-like {0,1} matrices: top max(1, floor(0.01 N)) ones per row."""
    import numpy as np
    rng = np.random.RandomState(seed)
    m = max(1, int(density * N))
    tmd = np.zeros((N, N), np.float32)
    pa = np.zeros((N, N), np.float32)
    for i in range(N):
        cols = rng.permutation(N)[: m + 1]
        tmd[i, cols[:m]] = 1.0
        pa[i, cols] = 1.0
    return tmd, pa


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return dict(hbm=d["hbm_gbs"], tensor=d.get("bf16_tflops_sustained", d["bf16_tflops"]), src="measured")
    return dict(hbm=6650.0, tensor=1400.0, src="fallback")


class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.FIELDS}",
                                       "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            out, _ = self.p.communicate(timeout=5)
        except Exception:
            self.p.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the reference's own model (oracle/_ref) or the oracle port, on host cores
# ----------------------------------------------------------------------------------------------
def reference_batch(w, mode):
    """The conf's own batch, reduced only when the host cannot hold the reference's autograd saves: its
    Python loop keeps ~3 (B, N, N) fp32 tensors per (time step, order) and block."""
    batch = w["conf_B"]
    if mode != "train":
        return batch, ""
    try:
        import psutil
        avail = psutil.virtual_memory().available
    except Exception:
        avail = 32 << 30
    per_sample = w["nb"] * w["T"] * w["K"] * 3.2 * w["N"] ** 2 * 4
    fit = int(0.6 * avail / per_sample)
    if fit >= batch:
        return batch, ""
    fit = max(1, fit)
    return fit, f" (conf batch {batch} needs ~{batch * per_sample / 2**30:.0f} GiB of autograd saves; host has {avail / 2**30:.0f} GiB free)"


def cpu_step_factory(w, batch, mode="train"):
    """One step of the workload on the host cores.  Returns (step, kind, description)."""
    torch.set_num_threads(os.cpu_count())
    from oracle.build_ref import import_reference_make_model
    g = torch.Generator().manual_seed(0)
    x = torch.randn(batch, w["N"], 1, w["T"], generator=g)
    y = torch.randn(batch, w["N"], w["P"], generator=g)
    make_ref = None if os.environ.get("DSTAG_BENCH_PORT") == "1" else import_reference_make_model()
    if make_ref is not None:
        tmd, pa = synthetic_graph(w["N"], 0)
        torch.manual_seed(1)
        # argument order of train_DSTAGNN_my.py:85-104; graph = AG => adj_merge = adj_TMD
        net = make_ref("cpu", 1, w["nb"], 1, w["K"], 32, 32, 1, torch.FloatTensor(tmd), torch.FloatTensor(pa),
                       torch.FloatTensor(tmd), w["P"], w["T"], w["N"], w["D"], 32, 32, w["H"])
        kind, desc = "reference", "unmodified reference model/DSTAGNN_my.py make_model('cpu') (oracle/_ref)"
    else:
        from oracle import dstagnn_oracle as O
        net = O.build_oracle(w["N"], w["T"], w["K"], w["H"], w["D"], w["nb"], P=w["P"], faithful=True)
        kind, desc = "port", "oracle port of the reference CPU path (model/DSTAGNN_my.py op order)"
    if mode == "train":
        net.train()
        opt = torch.optim.Adam(net.parameters(), lr=1e-4)
        crit = torch.nn.SmoothL1Loss()

        def step():
            opt.zero_grad()
            loss = crit(net(x), y)
            loss.backward()
            opt.step()
            return float(loss.detach())
    else:
        net.eval()

        def step():
            with torch.no_grad():
                return float(net(x).sum())
    return step, kind, desc


def cpu_baseline(w, mode="train"):
    """Bounded sample of the same workload: the conf's batch, one warm-up + best of 2 timed steps."""
    batch, note = reference_batch(w, mode)
    step, kind, desc = cpu_step_factory(w, batch, mode)
    step()
    ts = []
    for _ in range(2):
        t0 = time.time(); step(); ts.append(time.time() - t0)
    best = min(ts)
    return {"value": batch / best, "unit": "samples/s", "cores": os.cpu_count(), "kind": kind,
            "sample": f"{desc}, fp32, torch CPU {torch.get_num_threads()} threads, {mode} step at batch {batch}{note}, "
                      f"1 warm-up + best of 2"}


def run_reference(args, w):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    batch, note = reference_batch(w, args.mode)
    if args.batch:
        batch, note = args.batch, " (--batch override)"
    step, kind, desc = cpu_step_factory(w, batch, args.mode)
    # bounded run: the whole --steps / --warmup schedule must end within a few minutes
    t0 = time.time(); step(); t1 = time.time() - t0
    budget = 240.0
    steps, warmup = args.steps, max(0, args.warmup - 1)
    if t1 * (steps + warmup) > budget:
        warmup = 0
        steps = max(2, min(args.steps, int(budget / t1)))
    for _ in range(warmup):
        step()
    t0 = time.time()
    for _ in range(steps):
        step()
    dt = time.time() - t0
    val = batch * steps / dt
    sample = (f"{desc}, fp32, {torch.get_num_threads()} threads, {args.mode} step at batch {batch}{note}; "
              f"{steps} timed steps" + ("" if steps == args.steps else f" (of the {args.steps} requested: time bound)"))
    line = {"impl": "reference", "metric": metric_name(args), "value": val, "unit": "samples/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": warmup + 1, "ms_per_step": 1e3 * dt / steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": workload_config(args, w, batch),
            "cpu_baseline": {"value": val, "unit": "samples/s", "cores": os.cpu_count(), "kind": kind,
                             "sample": sample},
            "e2e": {"value": val, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def needs_l2_flush(args, w):
    """True when one step's working set may fit the 126 MB L2 (small graphs, inference): the timed steps are
    then separated by an L2 flush.  PEMS07 / GAMBIA / SYN4096 training moves > 1 GB per step."""
    return args.mode == "infer" or w["N"] < 800


def workload_config(args, w, batch):
    graphed = args.impl == "ours" and not args.no_graph and (args.gpus <= 1 or os.environ.get("DSTAG_GRAPH_MULTI", "1") != "0")
    what = (f"train step (zero_grad+fwd+SmoothL1+bwd{'+allreduce' if args.gpus > 1 else ''}+Adam)"
            if args.mode == "train" else "inference (eval, no_grad forward)") + (", CUDA graph" if graphed else "")
    return {"workload": f"{args.workload}: N={w['N']} T={w['T']} K={w['K']} H={w['H']} d_model={w['D']} "
                        f"nb_block={w['nb']} in_channels=1, {what}",
            "batch_per_gpu": batch, "global_batch": batch * max(1, args.gpus),
            "parallelism": f"dp{args.gpus}", "graph": "synthetic 1%-dense STAG/STRG, seed 0",
            "l2_policy": ("L2 flushed (256 MB fill) between timed steps, outside the per-step CUDA-event pairs"
                          if needs_l2_flush(args, w) and args.impl == "ours" else
                          "per-step working set (>1 GB of N x N maps and activations) exceeds the 126 MB L2")}


# ----------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------
def kernel_family(tag):
    """Kernel FUNCTION behind a profiler tag: every contraction runs gemm_tc_kernel<> (gemm_simt_kernel<> below 2^19
    MACs / K < 8); the other entry points are one kernel each."""
    if tag.startswith("gemm:") or tag.startswith("cheb_") or tag.startswith("sat_P") or (tag.startswith("gtu") and "_" in tag) \
            or tag in ("fcmy_dw",):
        return "contraction engine: gemm_tc_kernel<> (tcgen05 / TMEM / TMA), all contractions of the step"
    return {"dstag_gtu_conv_gate_fwd": "gtu_fwd_tc_kernel<> (fused GTU forward, tcgen05)",
            "dstag_gtu_gate_conv_bwd": "gtu_bwd_tc_kernel<> (fused GTU backward, tcgen05)"}.get(tag, tag)


def ncu_traffic(w, args, B, fam):
    """DRAM bytes (read + write) per launch of the dominant kernel function, from the committed ncu capture of one
    step (profiles/r02_ncu_traffic.json: {family: mean bytes per launch}); only valid for the configuration it was
    taken on."""
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r02_ncu_traffic.json")
    if (args.workload != "pems07" or args.precision != "tf32x3" or B != 64 or args.mode != "train"
            or not os.path.exists(path)):
        return None
    with open(path) as f:
        d = json.load(f)
    for key, v in d.items():
        if key != "_meta" and key.split(":")[0] in fam:
            return v
    return None


def run_ours(args, w):
    import torch.distributed as dist
    from dstagnn_drought_b200 import _lib, make_model, set_precision
    from dstagnn_drought_b200.step import TrainStep

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != max(1, args.gpus) and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if not args.no_graph:
            from dstagnn_drought_b200.step import multi_gpu_graph_env
            multi_gpu_graph_env()
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()
    set_precision(args.precision)
    train = args.mode == "train"

    tmd, pa = synthetic_graph(w["N"], 0)
    torch.manual_seed(1)
    net = make_model(dev, 1, w["nb"], 1, w["K"], 32, 32, 1, torch.from_numpy(tmd), torch.from_numpy(pa),
                     torch.from_numpy(tmd), w["P"], w["T"], w["N"], w["D"], 32, 32, w["H"])
    if train:
        net.train()
    else:
        net.eval()
    # (multi-rank steps are captured too, NCCL all-reduces included, unless DSTAG_GRAPH_MULTI=0)
    use_graph = not args.no_graph and (world == 1 or os.environ.get("DSTAG_GRAPH_MULTI", "1") != "0")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    flush = needs_l2_flush(args, w)
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev) if flush else None

    def timed(fn, steps):
        barrier()
        if flush:
            # small working sets: evict L2 between steps; every step gets its own event pair so that the
            # flush itself is outside the timed region
            evs = []
            for _ in range(steps):
                flush_buf.fill_(1)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                fn()
                b.record()
                evs.append((a, b))
            barrier()
            ms = sum(a.elapsed_time(b) for a, b in evs)
        else:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(steps):
                fn()
            b.record()
            barrier()
            ms = a.elapsed_time(b)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    train_steps = {}

    def measure(B, with_cpu):
        g = torch.Generator().manual_seed(100 + rank)
        x_host = torch.randn(B, w["N"], 1, w["T"], generator=g).pin_memory()
        y_host = torch.randn(B, w["N"], w["P"], generator=g).pin_memory()
        x_dev, y_dev = x_host.to(dev), y_host.to(dev)
        # the step's result read back by the host: the loss (train) or the predictions (inference)
        res_host = (torch.zeros((), dtype=torch.float32) if train else torch.zeros(B, w["N"], w["P"])).pin_memory()

        if train:
            # the product's own step (zero-grad, forward, SmoothL1, backward, [all-reduce], Adam), captured in a
            # CUDA graph on one GPU: dstagnn_drought_b200.step.TrainStep
            ts = train_steps.get(B)
            if ts is None:
                ts = train_steps[B] = TrainStep(net, 1e-4, graph=use_graph, warmup=2)

            def step(x, y):
                return ts.step(x, y)

            def eager_step(x, y):            # host-launched kernels: what the per-kernel event profile needs
                return ts._eager(x, y)

            def replays():
                return ts.replays

            def kernels_per_replay():
                return ts.kernels_per_replay
        else:
            # inference: eval() / no_grad() forward, captured in a CUDA graph per batch size
            x_static = torch.empty_like(x_dev)
            state = {"graph": None, "y": None, "calls": 0, "replays": 0, "kpr": 0}

            def replays():
                return state["replays"]

            def kernels_per_replay():
                return state["kpr"]

            def eager_step(x, y):
                with torch.no_grad():
                    return net(x)

            def step(x, y):
                if not use_graph:
                    return eager_step(x, y)
                x_static.copy_(x, non_blocking=True)
                state["calls"] += 1
                if state["graph"] is None:
                    with torch.no_grad():
                        if state["calls"] <= 2:
                            return net(x_static)
                        g = torch.cuda.CUDAGraph()
                        torch.cuda.synchronize()
                        n0 = _lib.launch_count()
                        with torch.cuda.graph(g):
                            state["y"] = net(x_static)
                        state["kpr"] = _lib.launch_count() - n0
                        state["graph"] = g
                state["graph"].replay()
                state["replays"] += 1
                return state["y"]

        def step_e2e():
            x = x_host.to(dev, non_blocking=True)
            y = y_host.to(dev, non_blocking=True) if train else None
            res_host.copy_(step(x, y), non_blocking=True)

        for _ in range(max(args.warmup, 3)):
            step(x_dev, y_dev)
        barrier()

        # ---- headline: device-resident inputs ----
        sampler = ClockSampler(local) if rank == 0 else None
        _lib.reset_launch_count()
        r0 = replays()
        ms = timed(lambda: step(x_dev, y_dev), args.steps)
        # kernels of this library that ran in the timed region: launched from the host, or replayed from the graph
        launches = _lib.launch_count() + (replays() - r0) * kernels_per_replay()
        # ---- end to end through the module API with host buffers ----
        ms_e2e = timed(step_e2e, args.steps)
        torch.cuda.synchronize()
        clocks = sampler.stop() if sampler else None
        final = float(res_host) if train else float(res_host.abs().mean())

        # ---- per-kernel CUDA-event profile of the same step (separate pass; not the headline) ----
        _lib.prof_enable(True)
        prof_steps = min(args.steps, 3)
        for _ in range(prof_steps):
            eager_step(x_dev, y_dev)
        torch.cuda.synchronize()
        prof = _lib.prof_results()
        _lib.prof_enable(False)
        if rank != 0:
            return

        pk = peaks()
        gemm_ms = sum(v[1] for v in prof.values()) / prof_steps
        # ---- roofline of the dominant kernel FUNCTION of this step (the one with the most device time) ----
        fams = {}
        for tag, v in prof.items():
            fams.setdefault(kernel_family(tag), []).append(v)
        fam, rows = max(fams.items(), key=lambda kv: sum(v[1] for v in kv[1]))
        f_calls = sum(v[0] for v in rows)
        f_ms = sum(v[1] for v in rows)
        f_flops = sum(v[2] for v in rows)
        f_bytes = sum(v[3] for v in rows)
        # which roof binds: per launch group the larger of flops / tensor peak and bytes / HBM peak
        t_tensor = sum(v[2] / (pk["tensor"] * 1e12) for v in rows)
        t_hbm = sum(v[3] / (pk["hbm"] * 1e9) for v in rows)
        floor_ms = 1e3 * sum(max(v[2] / (pk["tensor"] * 1e12), v[3] / (pk["hbm"] * 1e9)) for v in rows)
        bound = "tensor" if t_tensor >= t_hbm else "hbm"
        if bound == "tensor":
            achieved, peak, unit = f_flops / (f_ms * 1e-3) / 1e12, pk["tensor"], "TFLOP/s"
            src = (f"{pk['src']} dense bf16 (sustained); the path computes fp32-grade 3xTF32 products: 3 TF32 MMAs (each at half "
                   "the bf16 rate) per algorithmic flop, so its ceiling is peak/6 -- cuBLAS TF32 measured at "
                   "740 TFLOP/s on this pool (profiles/r02_tf32_peak.json) puts that ceiling at 247 TFLOP/s")
        else:
            achieved, peak, unit = f_bytes / (f_ms * 1e-3) / 1e9, pk["hbm"], "GB/s"
            src = f"{pk['src']} HBM copy bandwidth (MEASURED_PEAKS.json)"
        roofline = {"bound": bound, "kernel": fam, "achieved": achieved, "peak": peak, "unit": unit,
                    "frac": achieved / peak, "peak_source": src, "traffic": ncu_traffic(w, args, B, fam),
                    "launches_per_step": f_calls / prof_steps, "avg_launch_ms": f_ms / max(f_calls, 1),
                    "share_of_step": (f_ms / prof_steps) / (ms / args.steps), "precision": args.precision,
                    "algorithmic_gflop_per_step": f_flops / prof_steps / 1e9,
                    "algorithmic_gb_per_step": f_bytes / prof_steps / 1e9,
                    # both roofs together: sum over the family's launches of max(flops / tensor peak, bytes / HBM peak)
                    "mixed_floor_ms_per_step": floor_ms / prof_steps,
                    "mixed_frac": floor_ms / f_ms if f_ms > 0 else 0.0}
        # step roofline = sum over ALL kernels of max(flops_k / tensor peak, bytes_k / HBM peak) / step time
        step_floor = 1e3 * sum(max(v[2] / (pk["tensor"] * 1e12), v[3] / (pk["hbm"] * 1e9)) for v in prof.values()) / prof_steps
        step_roofline = {"floor_ms": step_floor, "frac": step_floor / (ms / args.steps),
                         "tflops_per_step": sum(v[2] for v in prof.values()) / prof_steps / 1e12,
                         "gb_per_step": sum(v[3] for v in prof.values()) / prof_steps / 1e9}
        breakdown = {t: {"calls_per_step": v[0] / prof_steps, "ms_per_step": round(v[1] / prof_steps, 4),
                         "tflops": round(v[2] / max(v[1], 1e-9) / 1e9, 2),
                         "gbs": round(v[3] / max(v[1], 1e-9) / 1e6, 1)} for t, v in
                     sorted(prof.items(), key=lambda kv: -kv[1][1])}

        cpu = None
        if world == 1 and with_cpu and not args.no_cpu_baseline:
            cpu = cpu_baseline(w, args.mode)

        n = max(1, args.gpus)
        h2d = x_host.numel() * 4 + (y_host.numel() * 4 if train else 0)
        d2h = res_host.numel() * 4
        line = {"metric": metric_name(args),
                "value": B * n * args.steps / (ms * 1e-3), "unit": "samples/s", "n_gpus": n, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": workload_config(args, w, B),
                "e2e": {"value": B * n * args.steps / (ms_e2e * 1e-3), "unit": "samples/s",
                        "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e / args.steps},
                "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "step_roofline": step_roofline,
                "cpu_baseline": cpu,
                "precision": args.precision, "final_loss" if train else "mean_abs_prediction": final,
                "kernel_ms_per_step_profiled": round(gemm_ms, 3)}
        print(json.dumps(line), flush=True)
        out_dir = os.path.join(ROOT, "gpurun_out")
        if os.path.isdir(out_dir):
            with open(os.path.join(out_dir, f"breakdown_{args.workload}_{args.mode}_b{B}_{args.precision}_n{n}.json"), "w") as f:
                json.dump({"ms_per_step": ms / args.steps, "launches_per_step": launches / args.steps,
                           "breakdown": breakdown}, f, indent=1)
        if args.verbose:
            for t, v in breakdown.items():
                print(f"  {t:45s} {v['calls_per_step']:6.1f} calls {v['ms_per_step']:9.3f} ms  {v['tflops']:8.2f} TF/s "
                      f"{v['gbs']:8.1f} GB/s", file=sys.stderr)

    batches = [int(b) for b in args.sweep.split(",")] if args.sweep else [args.batch or w["B"]]
    for i, B in enumerate(batches):
        measure(B, with_cpu=(i == 0))
    if world > 1:
        from dstagnn_drought_b200.step import shutdown_process_group
        shutdown_process_group(graphs_captured=use_graph)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="pems07", choices=sorted(WORKLOADS))
    ap.add_argument("--precision", default="tf32x3", choices=["fp32", "tf32x3", "tf32"])
    ap.add_argument("--batch", type=int, default=0, help="override the per-GPU batch")
    ap.add_argument("--mode", default="train", choices=["train", "infer"])
    ap.add_argument("--sweep", default="", help="comma-separated per-GPU batch sizes: one JSON line each")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel from the host (no CUDA graph)")
    ap.add_argument("--verbose", action="store_true")
    args = ap.parse_args()
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, w)
    else:
        run_ours(args, w)


if __name__ == "__main__":
    main()
