#!/usr/bin/env python
"""Benchmark of the GS/GD hologram path (BASELINE.json metric: iterations/s at 1024^2).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--alg gd|gs] [--precision fp32|fp64]
                    [--size 1024] [--batch 32] [--loops 100] [--impl reference]

A *step* is one pass of the hot path over one batch of synthetic targets: BATCH holograms of
SIZE x SIZE, LOOPS iterations each (default: config 2 of BASELINE.json -- gradient descent, 100
iterations, 1024^2 -- followed by the wavefront-mask add + 8-bit quantisation), entirely on the
device with inputs resident in HBM.  `value` = BATCH*LOOPS*K / time [iterations/s], summed over
ranks (weak scaling: every rank runs its own batch; the path has no collective).

`e2e` is the same work through the reference-facing API with HOST buffers
(algorithms.gradient_descent(target, args) -> numpy, then display_holograms.hologram_to_grey),
host<->device copies inside the timed region.

`--impl reference` times the CPU implementation (the oracle's numpy restatement of the reference,
scipy.fft with all host threads) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import contextlib
import io
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=10)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="b200", choices=["b200", "reference"])
    p.add_argument("--alg", default="gd", choices=["gd", "gs"])
    p.add_argument("--precision", default="fp32", choices=["fp32", "fp64"])
    p.add_argument("--size", type=int, default=1024)
    p.add_argument("--height", type=int, default=0, help="plane height (default: --size)")
    p.add_argument("--batch", type=int, default=32)
    p.add_argument("--loops", type=int, default=100)
    p.add_argument("--e2e-steps", type=int, default=5)
    p.add_argument("--cpu-loops", type=int, default=40, help="iterations of the CPU baseline sample")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-movie", action="store_true")
    p.add_argument("--workload", default="batch", choices=["batch", "slab"],
                   help="batch: configs[1] (default).  slab: ONE size^2 plane row-split over the ranks (configs[4]), GS")
    return p.parse_args()


def env_rank():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def workload_name(a, shape):
    alg = "gradient_descent" if a.alg == "gd" else "gerchberg_saxton"
    return (f"{alg} {shape[0]}x{shape[1]} uint8 noise targets, {a.loops} iterations, batch {a.batch}/GPU, "
            f"+ wavefront-mask add + 8-bit quantise (BASELINE.json configs[1])")


def bytes_per_px(alg, precision):
    """GS (8c+4), GD (9c+4) algorithmic bytes per pixel per iteration."""
    c = 8 if precision == "fp32" else 16
    return (8 * c + 4) if alg == "gs" else (9 * c + 4)


def kernel_bytes_per_px(kind, alg, precision):
    """Algorithmic bytes per pixel of ONE launch of a pass kernel (same accounting as SURVEY 8d:
    every plane the pass must read or write once, target accounted at 4 B/px)."""
    c = 8 if precision == "fp32" else 16
    if kind == "col_pass":
        return 2 * c + 4                       # read X, write Y, read target
    if kind == "row_pass":
        return 2 * c if alg == "gs" else 4 * c   # GS: read Y, write X; GD: + read x, write x
    if kind == "col_stats":
        return c                               # read X
    raise KeyError(kind)


def measured_traffic(kind, alg, precision, shape, batch):
    """DRAM bytes (read + write) per launch of the dominant kernel from the committed ncu capture
    (profiles/traffic.json), if one exists for exactly this configuration."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(path):
        return None
    key = f"{kind}|{alg}|{precision}|{shape[0]}x{shape[1]}|batch{batch}"
    return json.load(open(path)).get(key)


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock / throttle-reason samples DURING the timed region, taken with NVML from a background
    thread (an `nvidia-smi -lms` child process perturbed the timed region on this pool: bimodal step times)."""
    THROTTLE = {"hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40, "sw_power_cap": 0x4}

    def __init__(self, index, period_s=None):
        period_s = float(os.environ.get("SLM_BENCH_SAMPLER_PERIOD", "0.2")) if period_s is None else period_s
        self.index, self.period, self.samples, self._stop, self.thread, self.err = index, period_s, [], False, None, None

    def start(self):
        import threading
        try:
            import pynvml
            pynvml.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(visible.split(",")[self.index]) if visible and visible.split(",")[self.index].isdigit() else self.index
            self.nv, self.h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(phys)
        except Exception as exc:
            self.err = f"nvml unavailable: {exc}"
            return

        def loop():
            nv, h = self.nv, self.h
            while not self._stop:
                try:
                    t_s = time.time()
                    row = (t_s, nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM),
                           nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM),
                           nv.nvmlDeviceGetCurrentClocksEventReasons(h), nv.nvmlDeviceGetPowerUsage(h) / 1000.0)
                    self.samples.append(row + (time.time() - t_s,))
                except Exception as exc:
                    self.err = str(exc)
                    return
                time.sleep(self.period)
        self.thread = threading.Thread(target=loop, daemon=True)
        self.thread.start()

    def stop(self, t_begin=None, t_end=None):
        self._stop = True
        if self.thread:
            self.thread.join(timeout=2)
        rows = [r for r in self.samples if t_begin is None or t_begin - 0.05 <= r[0] <= t_end + 0.05] or self.samples[-1:]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [self.err or "no samples"]}
        reasons = sorted(k for k, bit in self.THROTTLE.items() if any(r[3] & bit for r in rows))
        return {"sm_mhz": float(np.median([r[1] for r in rows])), "sm_max_mhz": float(max(r[2] for r in rows)),
                "reasons": reasons, "samples": len(rows), "power_w_max": max(r[4] for r in rows), "source": "nvml",
                "sample_ms_max": round(1e3 * max(r[5] for r in rows), 2)}


# ------------------------------------------------------------------------------------------------
def cpu_port_iterations_per_s(alg, shape, loops, workers):
    """Time the oracle's restatement of the reference loop on the host."""
    import scipy.fft
    from oracle import numpy_port as P
    from spatial_light_modulator_module_b200 import synthetic
    t = synthetic.noise_target(shape, seed=0)
    with scipy.fft.set_workers(workers), np.errstate(all="ignore"):
        t0 = time.perf_counter()
        if alg == "gd":
            P.gd_run(t, loops)
        else:
            P.gs_run(t, loops)
        dt = time.perf_counter() - t0
    return loops / dt, dt


def run_reference(a):
    rank, _, world = env_rank()
    if rank != 0:
        return
    shape = (a.height or a.size, a.size)
    cores = os.cpu_count() or 1
    sample_loops = max(2, min(a.loops, 10))
    times = []
    for i in range(a.warmup + a.steps):
        _, dt = cpu_port_iterations_per_s(a.alg, shape, sample_loops, cores)
        if i >= a.warmup:
            times.append(dt)
    total = sum(times)
    value = sample_loops * a.steps / total
    sample = (f"each step = 1 hologram x {sample_loops} iterations of the numpy restatement (oracle/numpy_port.py) "
              f"incl. setup, scipy.fft workers={cores}")
    emit(({
        "impl": "reference", "metric": "GS/GD iterations/sec at 1024^2", "value": value, "unit": "iterations/s",
        "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * total / a.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(a, shape)},
        "cpu_baseline": {"value": value, "unit": "iterations/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------------------------------------
def run_b200(a):
    import torch
    import torch.distributed as dist

    rank, local_rank, world = env_rank()
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from spatial_light_modulator_module_b200 import _ffi, algorithms, display_holograms, host_logic as hl, synthetic
    from spatial_light_modulator_module_b200.engine import Engine

    shape = (a.height or a.size, a.size)
    npx = shape[0] * shape[1]
    eng = Engine(shape, a.precision, a.batch)
    dev = torch.device("cuda", local_rank)

    # synthetic inputs, resident in HBM before the timed region
    targets = torch.from_numpy(np.stack([synthetic.noise_target(shape, seed=1000 * rank + i) for i in range(a.batch)])).to(dev)
    mask = torch.from_numpy(synthetic.random_mask(shape, seed=1)).to(dev)
    rng = np.random.default_rng(rank)
    x0 = torch.from_numpy(np.exp(2j * np.pi * rng.random((a.batch,) + shape)).astype(eng.complex_dtype)).to(dev)
    x = torch.empty_like(x0)
    during, _ = hl.learning_rate_schedule(0.005, 0, a.loops)
    norms = targets.reshape(a.batch, -1).amax(dim=1).double().cpu().numpy()

    def step():
        if a.alg == "gd":
            x.copy_(x0)
            res, _ = eng.gd(targets, x, during, a.loops, want_expected=False, norms=norms)
        else:
            res = eng.gs(targets, a.loops, want_expected=False, norms=norms)
        return eng.quantize(res.hologram, mask, 256, _ffi.QUANT_FLOOR)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0 and not os.environ.get("SLM_BENCH_NO_SAMPLER"):
        sampler.start()
    # Warm-up steps keep their result alive exactly like the timed ones (`q = step()`): the previous frame is still
    # referenced while the next one is allocated, so the caching allocator needs TWO output blocks -- left to the
    # timed region, the second one cost a cudaMalloc (2-70 ms) in its second step.
    q = None
    for _ in range(max(a.warmup, 3)):
        q = step()
    barrier()
    n0 = eng.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    wall0 = time.time()
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(a.steps)]
    e0.record()
    for i in range(a.steps):
        q = step()
        marks[i].record()
    e1.record()
    barrier()
    wall1 = time.time()
    ms = e0.elapsed_time(e1)
    step_ms = [round(([e0] + marks)[i].elapsed_time(marks[i]), 3) for i in range(a.steps)]
    launches = eng.launch_count() - n0
    clocks = sampler.stop(wall0, wall1) if rank == 0 else None
    if world > 1:
        tmax = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        ms = float(tmax.item())
    iters_total = a.batch * a.loops * a.steps * world
    value = iters_total / (ms * 1e-3)

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- per-kernel device timing (instrumented re-run of the same steps) ------------------------
    eng.profile(True)
    eng.profile_read()
    for _ in range(a.steps):
        step()
    prof = eng.profile_read()
    eng.profile(False)
    peak, peak_src = measured_peak()
    kernels = {}
    for kind in ("col_pass", "row_pass", "col_stats"):
        tot, cnt = prof[kind]
        if cnt:
            avg_ms = tot / cnt
            bts = kernel_bytes_per_px(kind, a.alg, a.precision) * npx * a.batch
            kernels[kind] = {"avg_ms": avg_ms, "launches": cnt, "share": None, "alg_bytes": bts,
                             "achieved_gbs": bts / (avg_ms * 1e-3) / 1e9}
    tsum = sum(v[0] for v in prof.values())
    for kind in kernels:
        kernels[kind]["share"] = prof[kind][0] / tsum
    dom = max(kernels, key=lambda k: prof[k][0])
    names = {"col_pass": f"col_group_kernel<{'GS' if a.alg == 'gs' else 'GD_POST'}> (Fourier-plane pass)",
             "row_pass": "row_pass_kernel (SLM-plane pass)", "col_stats": "col_group_kernel<STATS> (max pre-pass)"}
    roof = {"bound": "hbm", "kernel": names[dom],
            "achieved": kernels[dom]["achieved_gbs"], "peak": peak, "unit": "GB/s",
            "frac": kernels[dom]["achieved_gbs"] / peak, "traffic": measured_traffic(dom, a.alg, a.precision, shape, a.batch),
            "peak_source": peak_src,
            "alg_bytes_per_launch": kernels[dom]["alg_bytes"], "avg_launch_ms": kernels[dom]["avg_ms"],
            "share_of_step": kernels[dom]["share"]}
    iter_bytes = bytes_per_px(a.alg, a.precision) * npx
    per_iter_s = (ms * 1e-3) / (a.batch * a.loops * a.steps)
    iteration_roofline = {"alg_bytes_per_iteration": iter_bytes, "achieved": iter_bytes / per_iter_s / 1e9, "peak": peak,
                          "unit": "GB/s", "frac": iter_bytes / per_iter_s / 1e9 / peak,
                          "note": "SURVEY 8(d) figure ((8c+4) GS / (9c+4) GD bytes per pixel per iteration) / measured time per iteration"}

    # ---- latency of ONE hologram (batch 1), device resident -----------------------------------------
    eng1 = Engine(shape, a.precision, 1)
    t1, x1, x01 = targets[:1].contiguous(), x[:1].contiguous(), x0[:1].contiguous()

    def step1():
        if a.alg == "gd":
            x1.copy_(x01)
            r, _ = eng1.gd(t1, x1, during, a.loops, want_expected=False, norms=norms[:1])
        else:
            r = eng1.gs(t1, a.loops, want_expected=False, norms=norms[:1])
        return eng1.quantize(r.hologram, mask, 256, _ffi.QUANT_FLOOR)
    for _ in range(3):
        step1()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(a.steps):
        step1()
    e1.record()
    torch.cuda.synchronize()
    lat_ms = e0.elapsed_time(e1) / a.steps
    single = {"ms_per_hologram": lat_ms, "iterations_per_s": a.loops / (lat_ms * 1e-3),
              "frac_of_hbm_roofline": iter_bytes * a.loops / (lat_ms * 1e-3) / 1e9 / peak,
              "note": "batch 1: working set fits in L2, launch/latency bound"}
    eng1.close()

    # ---- config 3 sample: optical-trap movie frames, GS 50 iterations at the SLM shape -----------------------
    movie = None
    if not a.no_movie:
        mshape, mframes, mloops = (768, 1024), 64, 50
        engm = Engine(mshape, a.precision, mframes)
        frames_dev = torch.from_numpy(synthetic.movie_frames(mframes)).to(dev)
        mnorms = np.full(mframes, 255.0)
        for _ in range(2):
            engm.gs(frames_dev, mloops, want_expected=False, norms=mnorms)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(3):
            engm.gs(frames_dev, mloops, want_expected=False, norms=mnorms)
        e1.record()
        torch.cuda.synchronize()
        mt = e0.elapsed_time(e1) / 3 * 1e-3
        movie = {"workload": f"{mframes} trap frames {mshape[0]}x{mshape[1]}, GS {mloops} iterations, device resident",
                 "holograms_per_s": mframes / mt, "iterations_per_s": mframes * mloops / mt}
        engm.close()
        # the same movie through the drop-in driver: host uint8 frames in, host float64 holograms out
        from spatial_light_modulator_module_b200 import generate_hologram_sequence as ghs
        host_frames = synthetic.movie_frames(2 * mframes)
        for _ in range(2):      # warm-up (the engine page-locks host arrays it is handed a second time; result arrays are pooled)
            ghs.sequence_holograms(host_frames, mloops, precision=a.precision, batch=mframes // 2, gather=False)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ghs.sequence_holograms(host_frames, mloops, precision=a.precision, batch=mframes // 2, gather=False)
        torch.cuda.synchronize()
        et = time.perf_counter() - t0
        movie["e2e"] = {"holograms_per_s": 2 * mframes / et / world, "frames": 2 * mframes // world, "batch": mframes // 2,
                        "h2d_bytes": int(host_frames.nbytes) // world, "d2h_bytes": int(2 * mframes * mshape[0] * mshape[1] * 8) // world,
                        "api": "generate_hologram_sequence.sequence_holograms(frames_uint8, 50) -> float64 holograms"}

    # ---- end to end through the drop-in API with host buffers ----------------------------------------
    ns = argparse.Namespace(incomming_intensity="uniform", tolerance=0, max_loops=a.loops, gif=False, print_info=False,
                            plot_error=False, initial_guess="random", random_seed=42, white_attention=1,
                            learning_rate=0.005, unsettle=0, precision=a.precision, device=local_rank)
    host_targets = [synthetic.noise_target(shape, seed=77 + i) for i in range(a.e2e_steps + 1)]
    host_mask = synthetic.random_mask(shape, seed=1)
    fn = algorithms.gradient_descent if a.alg == "gd" else algorithms.gerchberg_saxton

    def e2e_once(t):
        with contextlib.redirect_stdout(io.StringIO()):
            holo, exp, errs = fn(t, ns)
        ns.learning_rate = 0.005
        return display_holograms.hologram_to_grey(holo, host_mask, 256), errs
    e2e_once(host_targets[-1])
    e2e_once(host_targets[-1])             # (second warm-up: the engine page-locks a host array it is handed twice -- the mask)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(a.e2e_steps):
        grey, errs = e2e_once(host_targets[i])
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / a.e2e_steps
    csz = 8 if a.precision == "fp32" else 16
    h2d = npx * 1 + npx * 8 + npx * 8     # target; hologram + mask for the quantiser (the GD initial guess is drawn on the device)
    d2h = npx * 8 * 2 + a.loops * 8 + npx                                    # hologram, expected, error curve, grey frame
    e2e = {"value": a.loops / e2e_s, "unit": "iterations/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
           "ms_per_hologram": 1e3 * e2e_s, "holograms_per_s": 1 / e2e_s,
           "api": f"algorithms.{fn.__name__}(target_uint8, args) -> numpy; display_holograms.hologram_to_grey(h, mask, 256)"}

    # ---- CPU baseline (oracle port, single core as the reference runs it) ------------------------------
    cpu = None
    if not a.no_cpu_baseline and world == 1:
        v, dt = cpu_port_iterations_per_s(a.alg, shape, min(a.cpu_loops, a.loops), 1)
        cpu = {"value": v, "unit": "iterations/s", "cores": 1, "kind": "port",
               "sample": f"1 hologram x {min(a.cpu_loops, a.loops)} iterations of oracle/numpy_port.py ({dt:.1f} s), "
                         f"scipy.fft workers=1 as in the reference; host has {os.cpu_count()} cores"}

    line = {
        "metric": "GS/GD iterations/sec at 1024^2", "value": value, "unit": "iterations/s", "n_gpus": world,
        "steps": a.steps, "warmup": max(a.warmup, 3), "ms_per_step": ms / a.steps, "step_ms": step_ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32" if a.precision == "fp32" else "f64", "data": "synthetic",
        "config": {"workload": workload_name(a, shape), "algorithm": a.alg, "shape": list(shape), "batch_per_gpu": a.batch,
                   "iterations": a.loops, "l2": f"inputs larger than L2: {3 * a.batch * npx * csz / 2**20:.0f} MiB of field planes per GPU",
                   "seeds": "targets default_rng(1000*rank+i), mask default_rng(1)"},
        "holograms_per_s": a.batch * a.steps * world / (ms * 1e-3),
        "gpu_launches": int(launches), "clocks": clocks, "e2e": e2e, "roofline": roof,
        "iteration_roofline": iteration_roofline, "kernels": kernels, "single_hologram": single, "movie_config3_sample": movie, "cpu_baseline": cpu,
    }
    emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_slab(a):
    """BASELINE configs[4]: one size^2 GS hologram, slab-decomposed over the ranks (strong scaling)."""
    import torch
    import torch.distributed as dist
    rank, local_rank, world = env_rank()
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from spatial_light_modulator_module_b200.slab import SlabEngine
    n, loops = a.size, a.loops
    rows = n // world
    eng = SlabEngine(n, world, rank, a.precision)
    slab = eng._mem_upload((np.random.default_rng(100 + rank).random((rows, n)) * 255).astype(np.uint8))   # resident in HBM

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    if rank == 0 and not os.environ.get("SLM_BENCH_NO_SAMPLER"):
        sampler.start()
    for _ in range(max(a.warmup, 1)):
        eng.gs(slab, 2, want_expected=False, on_device=True)
    barrier()
    n0 = eng.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    wall0 = time.time()
    e0.record()
    for _ in range(a.steps):
        holo, _, errs = eng.gs(slab, loops, want_expected=False, on_device=True)
    e1.record()
    barrier()
    wall1 = time.time()
    clocks = sampler.stop(wall0, wall1) if rank == 0 else None
    ms = e0.elapsed_time(e1)
    if world > 1:
        tmax = torch.tensor([ms], device=torch.device("cuda", local_rank), dtype=torch.float64)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        ms = float(tmax.item())
    launches = eng.launch_count() - n0
    if rank == 0:
        peak, peak_src = measured_peak()
        c = 8 if a.precision == "fp32" else 16
        it_bytes = (8 * c + 4) * n * n
        per_it = ms * 1e-3 / (a.steps * loops)
        emit(({
            "metric": f"GS iterations/sec on one {n}^2 plane (slab-decomposed)", "value": 1.0 / per_it, "unit": "iterations/s",
            "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 1), "ms_per_step": ms / a.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32" if a.precision == "fp32" else "f64", "data": "synthetic",
            "config": {"workload": f"gerchberg_saxton, one {n}x{n} uint8 noise target, {loops} iterations incl. setup and the final "
                                   f"hologram read-back, rows split over {world} GPU(s), 2 all-to-alls + 1 all-reduce per iteration "
                                   f"(BASELINE.json configs[4])"},
            "gpu_launches": int(launches), "final_error": float(errs[-1]), "clocks": clocks,
            "iteration_roofline": {"alg_bytes_per_iteration": it_bytes, "achieved": it_bytes / per_it / 1e9 / world, "peak": peak,
                                   "unit": "GB/s per GPU", "frac": it_bytes / per_it / 1e9 / world / peak, "peak_source": peak_src},
        }))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def emit(line: dict) -> None:
    """The ONE JSON line of this run, on the process's real stdout."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = 1

if __name__ == "__main__":
    args = parse()
    # watchdog: a run that takes absurdly long (default 20 min; SLM_BENCH_WATCHDOG=seconds) dumps every thread's stack
    # to stderr and exits instead of hanging its launcher (one 2-GPU run of this round stalled after NCCL's init and
    # could not be reproduced in five repeats)
    import faulthandler
    faulthandler.dump_traceback_later(float(os.environ.get("SLM_BENCH_WATCHDOG", "1200")), exit=True)
    # libraries (NCCL's version banner, ...) write to file descriptor 1: keep it for the JSON line alone
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "slab":
        run_slab(args)
    else:
        run_b200(args)
