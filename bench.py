#!/usr/bin/env python
"""bench.py -- throughput of the grid fluid step (advect + pressure projection).

Metric (BASELINE.json): Mcell-steps/s, whole job, plus the HBM roofline of the dominant kernel.
A "step" is one pass of examples/dec_fluid.rs:46-141 on the synthetic smoke plume of
SURVEY.md 8(d) (the shipped example's rectangles scaled by N/128; f64; dt 0.05; threshold 0.1;
100 CG iterations max; identity preconditioner).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--n GRID]

N = 1  -> BASELINE configs[1]: 1024^2 on one B200.
N > 1  -> BASELINE configs[2]: 8192^2 slab-decomposed over N GPUs (strong scaling), launched by
          torchrun with one rank per GPU.
--impl reference times the CPU oracle (a C port of the Rust reference, which cannot be built in
this image) on the host cores, rank 0 only.

Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "Mcell-steps/s (advect+project)"
UNIT = "Mcell-steps/s"
# algorithmic bytes per cell (f64), SURVEY.md 8(d): one step with I CG iterations = 144 + 88*I
BYTES_ADVECT, BYTES_NEGDIV, BYTES_CG_INIT, BYTES_CG_ITER, BYTES_PROJECT = 48, 24, 32, 88, 40


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy kernel)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


def load_traffic(workload_key):
    """ncu dram bytes per launch of the dominant kernel, if a profile summary has been committed."""
    path = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    try:
        with open(path) as f:
            return json.load(f).get(workload_key)
    except Exception:
        return None


class ClockSampler:
    """Samples SM clock and throttle reasons during the timed region (NVML, 50 ms period)."""
    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self._nv = None

    def _run(self):
        nv = self._nv
        while not self._stop.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
                mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self._h)) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h))
                for bit, name in self.REASONS.items():
                    if mask & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.05)

    def __enter__(self):
        if self._nv is not None:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thr is not None:
            self._thr.join(timeout=2)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(s)}


# ----------------------------------------------------------------------------------------------- CPU arm
def cpu_step_rate(n, mode, budget_s, full_steps_max=6):
    """Mcell-steps/s of the CPU oracle on the n^2 smoke-plume workload.

    Small grids: real steps from the zero state (at n >= 512 every step runs the full 100 CG
    iterations, so early steps cost what later ones do).  Large grids: one step is too long, so
    two steps with the CG capped at 4 and 12 iterations are timed and the 100-iteration step is
    extrapolated linearly (every CG iteration does identical work)."""
    from oracle import pano_oracle as O
    threads = O.set_threading(mode)
    prm = O.smoke_params(n)
    cells = n * n
    est_step_s = cells * 2.4e-6 if mode != O.ALL_PARALLEL else cells * 0.8e-6     # crude, only picks the strategy
    if est_step_s * 2 <= budget_s:
        S = O.FluidState(**prm)
        S.step()                                   # warm-up (page faults, first touch)
        t0 = time.perf_counter()
        k = 0
        while k < full_steps_max and (k == 0 or (time.perf_counter() - t0) * (k + 1) / k < budget_s):
            S.step()
            k += 1
        dt = (time.perf_counter() - t0) / k
        sample = f"{k} full steps of the {n}^2 smoke plume after 1 warm-up step (100 CG iterations each)"
        S.close()
    else:
        times = {}
        for iters in (4, 12):
            p2 = dict(prm, max_iterations=iters)
            S = O.FluidState(**p2)
            S.step()
            t0 = time.perf_counter()
            S.step()
            times[iters] = time.perf_counter() - t0
            S.close()
        per_iter = (times[12] - times[4]) / 8.0
        dt = times[4] + 96.0 * per_iter
        sample = (f"{n}^2: one step timed with the CG capped at 4 and at 12 iterations, "
                  f"extrapolated linearly to the 100-iteration step ({per_iter * 1e3:.1f} ms per iteration)")
    O.set_threading(O.SERIAL)
    used = 1 if mode == O.SERIAL else threads
    return cells / dt / 1e6, dt, used, sample


def run_reference(args, rank, world):
    if rank != 0:
        return 0
    from oracle import pano_oracle as O
    n = args.n or (1024 if args.gpus == 1 else 8192)
    mode = O.ALL_PARALLEL if args.cpu_variant == "parallel" else O.REFERENCE_FAITHFUL
    threads = O.set_threading(mode)
    prm = O.smoke_params(n)
    cells = n * n
    kind_note = ("C port of the Rust reference (rustc/cargo absent); threading as in the reference: rayon only in the "
                 "three derivative passes, everything else serial") if mode == O.REFERENCE_FAITHFUL else \
        "C port of the Rust reference; every pass OpenMP-parallel (upper bound for the CPU)"
    if n <= 2048:
        steps, warm = max(1, min(args.steps, 8)), max(1, min(args.warmup, 2))
        S = O.FluidState(**prm)
        for _ in range(warm):
            S.step()
        t0 = time.perf_counter()
        for _ in range(steps):
            S.step()
        total = time.perf_counter() - t0
        ms = total / steps * 1e3
        sample = f"{steps} full steps of the {n}^2 smoke plume after {warm} warm-up steps"
    else:
        rate, dt, _, sample = cpu_step_rate(n, mode, 60.0)
        steps, warm, ms = 1, 1, dt * 1e3
    O.set_threading(O.SERIAL)
    value = cells / (ms * 1e-3) / 1e6
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": warm, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong" if args.gpus > 1 else "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"2D smoke plume {n}x{n} MAC grid, advect + pressure projection (dec_fluid.rs loop body)",
                       "grid": [n, n], "cg_max_iterations": 100, "device": "host CPU"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                             "note": kind_note},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


# ----------------------------------------------------------------------------------------------- GPU arm
def run_ours(args, rank, world, local_rank):
    import numpy as np
    import torch
    import torch.distributed as dist

    import panopaea_b200 as P
    from panopaea_b200 import _lib, fluid

    multi = world > 1
    if multi:
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if args.gpus != world:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {args.gpus}")
    if multi:
        raise SystemExit("multi-GPU slab decomposition is not wired into bench.py yet")

    n = args.n or 1024
    ctx = P.Context(local_rank)
    for kv in args.opt:
        k, v = kv.split("=")
        ctx.set_option(k, int(v))
    sim = fluid.DecFluid(**fluid.smoke_params(n), ctx=ctx)
    cells = n * n
    K, W = args.steps, max(args.warmup, 3)
    flush = P.Grid2d((6144, 6144), ctx).new_simplex_2()       # 302 MB > 126 MB L2

    def barrier():
        ctx.sync()
        if multi:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up
    for _ in range(W):
        sim.step(want_info=False)
    info = sim.step(want_info=True)        # part of warm-up; tells us the iteration count of this regime
    barrier()

    # ---- timed region: K steps, L2 flushed before each, device time from CUDA events on the library's stream
    ctx.set_option("step_timing", 1)
    ctx.step_times()
    launches0 = ctx.launch_count()
    iters_total = 0
    total_ms = 0.0
    with ClockSampler(local_rank) as clk:
        barrier()
        for _ in range(K):
            flush.fill(0.0)                # cudaMemsetAsync of 302 MB: evicts the fields from L2
            ctx.timer_start()
            sim.step(want_info=False)
            total_ms += ctx.timer_stop_ms()
        barrier()
    launches = ctx.launch_count() - launches0
    phase_ms, phase_steps = ctx.step_times()
    ctx.set_option("step_timing", 0)
    last = sim.step(want_info=True)
    iters = last["applies"]
    if multi:
        t = torch.tensor([total_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    ms_per_step = total_ms / K
    value = cells * K / (total_ms * 1e-3) / 1e6

    # ---- e2e: the same step through the host-buffer entry point (pinned host fields, H2D + D2H every step)
    L = _lib.load()
    n1 = (n + 1) * n + n * (n + 1)
    bufs = []
    for count in (cells, n1, cells):
        p = C.c_void_p()
        _lib.check(L.pano_host_alloc(count * 8, C.byref(p)))
        bufs.append(p)
    h_density = np.ctypeslib.as_array(C.cast(bufs[0], C.POINTER(C.c_double)), shape=(cells,))
    h_vel = np.ctypeslib.as_array(C.cast(bufs[1], C.POINTER(C.c_double)), shape=(n1,))
    h_density[:] = sim.density.view_linear()
    h_vel[:] = sim.vel.view_linear()
    e2e_steps = max(3, min(K, 10))
    for _ in range(2):
        _lib.check(L.pano_fluid_step_host(ctx.handle, C.byref(sim.params), n, n, bufs[0], bufs[1], bufs[2], None))
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        _lib.check(L.pano_fluid_step_host(ctx.handle, C.byref(sim.params), n, n, bufs[0], bufs[1], bufs[2], None))
    barrier()
    e2e_s = time.perf_counter() - t0
    if multi:
        t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = cells * e2e_steps / e2e_s / 1e6
    for p in bufs:
        L.pano_host_free(p)

    # ---- roofline of the dominant kernel (the persistent CG kernel: phase 3)
    peak, peak_src = load_peaks()
    cg_ms = phase_ms[3] / max(1, phase_steps)
    cg_bytes = cells * (BYTES_CG_INIT + BYTES_CG_ITER * iters)
    achieved = cg_bytes / (cg_ms * 1e-3) / 1e9 if cg_ms > 0 else 0.0
    step_bytes = cells * (BYTES_ADVECT + BYTES_NEGDIV + BYTES_PROJECT + BYTES_CG_INIT + BYTES_CG_ITER * iters)
    roofline = {"bound": "hbm", "kernel": "k_cg_generic (persistent CG: init + %d iterations in one launch)" % iters,
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "peak_source": peak_src, "traffic": load_traffic(f"cg_{n}"),
                "algorithmic_bytes_per_launch": cg_bytes, "kernel_ms": cg_ms,
                "kernel_share_of_step": cg_ms / ms_per_step if ms_per_step > 0 else None,
                "step_achieved_gbs": step_bytes / (ms_per_step * 1e-3) / 1e9,
                "phase_ms": dict(zip(["inflow", "advect_all", "neg_divergence", "cg", "project"],
                                     [m / max(1, phase_steps) for m in phase_ms])),
                "note": ("working set (10 fields x %.1f MB) %s the 126 MB L2; algorithmic GB/s above the HBM peak means "
                         "cache residency, not an error" % (cells * 8 / 1e6, "fits in" if cells * 80 < 126e6 else "exceeds"))}

    line = None
    if rank == 0:
        # ---- CPU baseline on this box's cores: bounded sample of the same workload
        cpu = None
        if not args.no_cpu:
            from oracle import pano_oracle as O
            v, dt, cores, sample = cpu_step_rate(n, O.REFERENCE_FAITHFUL, 12.0, full_steps_max=4)
            v2, dt2, cores2, sample2 = cpu_step_rate(n, O.ALL_PARALLEL, 8.0, full_steps_max=4)
            cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                   "threading": "reference-faithful: threads only in the 3 derivative passes the reference runs under rayon",
                   "all_parallel": {"value": v2, "cores": cores2, "sample": sample2}}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W + 1,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if multi else "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": f"2D smoke plume {n}x{n} MAC grid, advect + pressure projection (dec_fluid.rs loop body)",
                           "grid": [n, n], "cg_iterations_per_step": iters, "cg_max_iterations": 100,
                           "threshold": 0.1, "timestep": 0.05, "options": args.opt, "parallelism": f"slab{world}" if multi else "single",
                           "l2": "flushed before every timed step (302 MB memset); within a step the CG re-reads its working set"},
                "e2e": {"value": e2e_value, "unit": UNIT, "steps": e2e_steps,
                        "h2d_bytes_per_step": (cells + n1) * 8, "d2h_bytes_per_step": (2 * cells + n1) * 8,
                        "call": "pano_fluid_step_host (pinned host fields in, fields + pressure out, every step)"},
                "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu, "clocks": clk.summary(),
                "cg_info_last_step": last, "cg_info_warm_step": info}
        print(json.dumps(line), flush=True)
    if multi:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=0, help="grid size override (multiple of 128)")
    ap.add_argument("--cpu-variant", default="faithful", choices=["faithful", "parallel"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--opt", action="append", default=[], help="library option key=value (e.g. cg_kernel=1)")
    args = ap.parse_args()
    rank, world, local_rank = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if args.impl == "reference":
        return run_reference(args, rank, world)
    return run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    sys.exit(main())
