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

# stdout carries exactly ONE line (the JSON).  Libraries write to file descriptor 1 behind Python's back (NCCL prints its
# version banner there), so the real stdout is kept aside for the JSON line and fd 1 is pointed at stderr for everything else.
_JSON_OUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)

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
    """Samples SM clock and throttle reasons during the timed region (NVML, 5 ms period)."""
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
            self._stop.wait(0.005)

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
    print(json.dumps(line), file=_JSON_OUT, flush=True)
    return 0


# ----------------------------------------------------------------------------------------------- GPU arm
class _Single:
    """One GPU: device-resident fields stepped by pano_fluid_step; e2e through pano_fluid_step_host."""

    def __init__(self, ctx, n, P, fluid, _lib):
        self.ctx, self.n, self._lib = ctx, n, _lib
        self.sim = fluid.DecFluid(**fluid.smoke_params(n), ctx=ctx)
        self.cells = n * n
        self.n1 = (n + 1) * n + n * (n + 1)
        self.workload_rows = n

    def step(self):
        self.sim.step(want_info=False)

    def info(self):
        self.ctx.sync()
        return self.sim.step(want_info=True)

    def e2e_setup(self, np):
        L, _lib = self._lib.load(), self._lib
        self.bufs = []
        for count in (self.cells, self.n1, self.cells):
            p = C.c_void_p()
            _lib.check(L.pano_host_alloc(count * 8, C.byref(p)))
            self.bufs.append(p)
        d = np.ctypeslib.as_array(C.cast(self.bufs[0], C.POINTER(C.c_double)), shape=(self.cells,))
        v = np.ctypeslib.as_array(C.cast(self.bufs[1], C.POINTER(C.c_double)), shape=(self.n1,))
        d[:] = self.sim.density.view_linear()
        v[:] = self.sim.vel.view_linear()
        self.h2d, self.d2h = (self.cells + self.n1) * 8, (2 * self.cells + self.n1) * 8
        self.call = "pano_fluid_step_host (pinned host fields in, fields + pressure out, every step)"

    def e2e_step(self):
        L, _lib = self._lib.load(), self._lib
        _lib.check(L.pano_fluid_step_host(self.ctx.handle, C.byref(self.sim.params), self.n, self.n, self.bufs[0], self.bufs[1],
                                          self.bufs[2], None))

    def e2e_teardown(self):
        L = self._lib.load()
        for p in self.bufs:
            L.pano_host_free(p)


class _Slab:
    """N GPUs: this rank's slab of the grid (pano_dist_*); e2e = upload my rows, step, download my rows."""

    def __init__(self, ctx, n, rank, world, dist_t, P, fluid, _lib):
        from panopaea_b200 import dist
        self.ctx, self.n, self._lib, self.dist = ctx, n, _lib, dist
        prm = fluid.smoke_params(n)
        self.D = dist.DistFluid(ctx, n, n, rank, world, {k: v for k, v in prm.items() if k not in ("h", "w")})
        handles = [None] * world
        dist_t.all_gather_object(handles, self.D.ipc_handle())
        self.D.connect_ipc(handles)
        dist_t.barrier()
        self.cells = n * n
        self.rows = self.D.y1 - self.D.y0

    def step(self):
        self.D.step()

    def info(self):
        self.D.step()
        return self.D.sync()

    def e2e_setup(self, np):
        L, _lib, D, dist = self._lib.load(), self._lib, self.D, self.dist
        self.bufs, self.counts = [], []
        for which in (dist.DENSITY, dist.VY, dist.VX, dist.PRESSURE):
            count = D._rows(which) * D._pitch(which)
            p = C.c_void_p()
            _lib.check(L.pano_host_alloc(count * 8, C.byref(p)))
            self.bufs.append(p)
            self.counts.append(count)
        for which in (dist.DENSITY, dist.VY, dist.VX):
            _lib.check(L.pano_dist_download(D._h, which, self.bufs[which], None))
        self.h2d, self.d2h = sum(self.counts[:3]) * 8, sum(self.counts) * 8       # bytes of THIS rank
        self.call = "pano_dist_upload x3 + pano_dist_step + pano_dist_sync + pano_dist_download x4 (pinned host rows of this rank)"

    def e2e_step(self):
        L, _lib, D, dist = self._lib.load(), self._lib, self.D, self.dist
        for which in (dist.DENSITY, dist.VY, dist.VX):
            _lib.check(L.pano_dist_upload(D._h, which, self.bufs[which]))
        D.step()
        D.sync()
        for which in (dist.DENSITY, dist.VY, dist.VX, dist.PRESSURE):
            _lib.check(L.pano_dist_download(D._h, which, self.bufs[which], None))

    def e2e_teardown(self):
        L = self._lib.load()
        for p in self.bufs:
            L.pano_host_free(p)


def run_ours(args, rank, world, local_rank):
    import numpy as np
    import torch
    import torch.distributed as dist_t

    import panopaea_b200 as P
    from panopaea_b200 import _lib, fluid

    multi = world > 1
    if args.gpus != world:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {args.gpus}")
    torch.cuda.set_device(local_rank)
    if multi:
        dist_t.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    n = args.n or (8192 if multi else 1024)
    ctx = P.Context(local_rank)
    for kv in args.opt:
        k, v = kv.split("=")
        ctx.set_option(k, int(v))
    job = _Slab(ctx, n, rank, world, dist_t, P, fluid, _lib) if multi else _Single(ctx, n, P, fluid, _lib)
    cells = n * n
    K, W = args.steps, max(args.warmup, 3)
    fields_mb = cells * 8 / 1e6 / world
    l2_resident = 10 * fields_mb < 126.0
    flush = P.Grid2d((6144, 6144), ctx).new_simplex_2() if l2_resident else None     # 302 MB > 126 MB L2

    def barrier():
        ctx.sync()
        if multi:
            dist_t.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if not multi:
            return x
        t = torch.tensor([x], device="cuda", dtype=torch.float64)
        dist_t.all_reduce(t, op=dist_t.ReduceOp.MAX)
        return float(t.item())

    # ---- warm-up
    for _ in range(W):
        job.step()
    info = job.info()                      # part of the warm-up; also tells the iteration count of this regime
    barrier()

    # ---- timed region: K steps; device time from CUDA events on the library's stream
    ctx.set_option("step_timing", 1)
    ctx.step_times()
    launches0 = ctx.launch_count()
    total_ms = 0.0
    with ClockSampler(local_rank) as clk:
        barrier()
        if flush is not None:              # small grid: evict the fields from L2 before every step, time each step
            for _ in range(K):
                flush.fill(0.0)
                ctx.timer_start()
                job.step()
                total_ms += ctx.timer_stop_ms()
        else:                              # inputs larger than L2: one event pair around the K steps
            ctx.timer_start()
            for _ in range(K):
                job.step()
            total_ms = ctx.timer_stop_ms()
        barrier()
    launches = ctx.launch_count() - launches0
    phase_ms, phase_steps = ctx.step_times()
    ctx.set_option("step_timing", 0)
    total_ms = max_over_ranks(total_ms)
    last = job.info()
    iters = last["applies"]
    ms_per_step = total_ms / K
    value = cells * K / (total_ms * 1e-3) / 1e6

    # ---- e2e: the same step with the fields in (pinned) host memory, H2D + D2H inside every step
    job.e2e_setup(np)
    e2e_steps = max(3, min(K, 10))
    for _ in range(2):
        job.e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        job.e2e_step()
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_value = cells * e2e_steps / e2e_s / 1e6
    h2d, d2h = job.h2d, job.d2h
    if multi:
        t = torch.tensor([float(h2d), float(d2h)], device="cuda", dtype=torch.float64)
        dist_t.all_reduce(t)
        h2d, d2h = int(t[0].item()), int(t[1].item())
    job.e2e_teardown()

    # ---- roofline of the dominant kernel (the persistent CG kernel: phase 3), for this rank's share of the grid
    peak, peak_src = load_peaks()
    cg_ms = max_over_ranks(phase_ms[3] / max(1, phase_steps))
    my_cells = cells / world
    cg_bytes = my_cells * (BYTES_CG_INIT + BYTES_CG_ITER * iters)
    achieved = cg_bytes / (cg_ms * 1e-3) / 1e9 if cg_ms > 0 else 0.0
    step_bytes = my_cells * (BYTES_ADVECT + BYTES_NEGDIV + BYTES_PROJECT + BYTES_CG_INIT + BYTES_CG_ITER * iters)
    if not multi and n <= 1024 and -(-n // 8) * n <= 5120:
        kernel = "k_cg_cluster"                    # one thread-block cluster (csrc/pano_cg_cluster.cu): grids up to ~40 k cells
    elif not multi and cells <= 1_200_000:
        kernel = "k_cg_resident2"
    else:
        kernel = "k_cg_stream"
    roofline = {"bound": "hbm", "kernel": "%s (persistent CG: init + %d iterations in one launch%s)" % (kernel, iters, ", per GPU" if multi else ""),
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "peak_source": peak_src, "traffic": load_traffic(f"cg_{n}_x{world}"),   # ncu dram bytes per launch (profiles/roofline_traffic.json), null if not captured
                "algorithmic_bytes_per_launch": cg_bytes, "kernel_ms": cg_ms,
                "kernel_share_of_step": cg_ms / ms_per_step if ms_per_step > 0 else None,
                "step_achieved_gbs": step_bytes / (ms_per_step * 1e-3) / 1e9,
                "phase_ms": dict(zip(["inflow", "advect_all", "neg_divergence", "cg", "project"],
                                     [m / max(1, phase_steps) for m in phase_ms])),
                "note": ("algorithmic bytes = SURVEY.md 8(d): 32 + 88 B per cell and CG iteration; the kernel's real traffic is 64 B "
                         "(fused search update, z not stored); per-GPU working set 10 fields x %.1f MB %s the 126 MB L2, and at "
                         "<= 1.2 Mcell the CG state stays in registers/shared memory for the whole solve, so algorithmic GB/s "
                         "above the HBM peak means on-chip residency, not an error"
                         % (fields_mb, "fits in" if l2_resident else "exceeds"))}

    # ---- beyond the reference: the same step with the multigrid preconditioner behind the pcg.rs trait seam (every solve
    # reaches the threshold; the reference-faithful identity solve above stops at its 100-iteration cap).  Reported aside,
    # never mixed into `value`.
    mg_line = None
    if not multi and not args.no_mg:
        sim = job.sim
        sim.params.precond = _lib.PRECOND_MULTIGRID
        for _ in range(3):
            sim.step(want_info=False)
        ctx.sync()
        km = max(3, min(K, 10))
        ctx.timer_start()
        for _ in range(km):
            sim.step(want_info=False)
        mg_ms = ctx.timer_stop_ms() / km
        mg_info = sim.step(want_info=True)
        sim.params.precond = _lib.PRECOND_IDENTITY
        mg_line = {"value": cells / (mg_ms * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": mg_ms, "steps": km,
                   "cg_iterations_per_step": mg_info["applies"], "final_residual": mg_info["final_residual"],
                   "note": "pano_step_params.precond = PANO_PRECOND_MULTIGRID (DESIGN.md 5b): converged solves, host-driven PCG loop"}

    # ---- N > 1 is a STRONG-scaling run of configs[2] (8192^2), while the N = 1 default is configs[1] (1024^2): so that the
    # speed-up can be read off one line, rank 0 also steps the same grid alone on its GPU (the other ranks wait at the barrier)
    same_1gpu = None
    if multi and not args.no_single:
        if rank == 0:
            solo = _Single(ctx, n, P, fluid, _lib)
            for _ in range(3):
                solo.step()
            ctx.sync()
            ks = max(2, min(K, 5))
            ctx.timer_start()
            for _ in range(ks):
                solo.step()
            solo_ms = ctx.timer_stop_ms() / ks
            same_1gpu = {"value": cells / (solo_ms * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": solo_ms, "steps": ks, "warmup": 3,
                         "note": "the same grid stepped by rank 0 alone (pano_fluid_step, k_cg_stream) right after the timed region"}
            del solo
        barrier()

    if rank == 0:
        cpu = None
        if not args.no_cpu and not multi:
            # CPU baseline on this box's cores: bounded sample of the same workload (N = 1 only)
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
                           "threshold": 0.1, "timestep": 0.05, "options": args.opt,
                           "parallelism": f"slab{world} (rows split over {world} GPUs, halos + reductions over NVLink peer memory)" if multi else "single",
                           "l2": ("flushed before every timed step (302 MB memset); within a step the CG re-reads its working set"
                                  if l2_resident else "not flushed: every rank's fields are larger than the 126 MB L2")},
                "e2e": {"value": e2e_value, "unit": UNIT, "steps": e2e_steps, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "call": job.call},
                "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu, "clocks": clk.summary(),
                "cg_info_last_step": last, "cg_info_warm_step": info}
        if mg_line is not None:
            line["multigrid_pcg_step"] = mg_line
        if same_1gpu is not None:
            line["one_gpu_same_workload"] = same_1gpu
        print(json.dumps(line), file=_JSON_OUT, flush=True)
    if multi:
        dist_t.barrier()
        dist_t.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", "--grid", dest="n", type=int, default=0, help="grid size override (multiple of 128); use --grid under torchrun, whose own parser claims --n")
    ap.add_argument("--cpu-variant", default="faithful", choices=["faithful", "parallel"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-mg", action="store_true", help="skip the multigrid-preconditioned step timing")
    ap.add_argument("--no-single", action="store_true", help="N > 1: skip the one-GPU run of the same grid")
    ap.add_argument("--opt", action="append", default=[], help="library option key=value (e.g. cg_kernel=1)")
    args = ap.parse_args()
    rank, world, local_rank = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if args.impl == "reference":
        return run_reference(args, rank, world)
    return run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    sys.exit(main())
