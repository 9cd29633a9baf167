#!/usr/bin/env python
"""bench.py -- throughput of the grid fluid step (advect + pressure projection).

Metric (BASELINE.json): Mcell-steps/s, whole job, plus the HBM roofline of the kernels.
A "step" is one pass of examples/dec_fluid.rs:46-141 on the synthetic smoke plume of
SURVEY.md 8(d) (the shipped example's rectangles scaled by N/128; f64; dt 0.05; threshold 0.1;
100 CG iterations max; identity preconditioner).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--grid G] [--poisson-only]

Every N runs the SAME workload, 8192^2 (BASELINE configs[2]; it fits one GPU, so the N = 1 line is the
strong-scaling denominator of the N = 2/4/8 lines): one GPU steps the whole grid, N GPUs step row slabs
of it (launched by torchrun with one rank per GPU).  The N = 1 line also carries `kernels_4096` (per-kernel
rooflines at 4096^2 = configs[3], the grid the >= 70 % target is quoted on) and short 1024^2 / 128^2 legs
(configs[1], configs[0]); the N > 1 lines carry `parity_vs_one_gpu` (fields against rank 0 stepping the
same grid alone) and `poisson_16384` (configs[4]: the pressure solve alone at 16384^2).
Protocol: >= 20 warm-up steps, K timed steps, every step its own CUDA-event lap on the
library's stream with no host synchronisation in between; `value` uses the total, the median lap is
reported next to it.
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
# The streaming CG kernels fuse the three passes SURVEY.md models (K5 + K6 + K7 = 88 B): per cell and iteration each of the four
# vectors they keep is read once and written once = 64 B, the opening pass reads b = 8 B.  The roofline fraction is
# quoted on these bytes (it cannot exceed 1 unless the L2 serves part of them); the 88-byte model is reported next to it.
BYTES_CG_ITER_FUSED, BYTES_CG_INIT_FUSED = 64, 8
NOMINAL_GBS = 8000.0
DEFAULT_GRID = 8192            # BASELINE configs[2]: the grid every N steps
POISSON_GRID = 16384           # BASELINE configs[4] (2-D branch)
PARITY_TOL = 1e-8              # N > 1 fields against the one-GPU run of the same steps (relative to max|field|)


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy kernel)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


def load_traffic():
    """ncu dram bytes per launch of the shipped kernels, from the committed profile summaries."""
    path = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    try:
        with open(path) as f:
            return json.load(f)
    except Exception:
        return {}


def workload_config(n, world, opts, iters=100, poisson=False):
    """The `config` object; the reference arm prints the same one (it times the same workload on the host)."""
    cells_mb = n * n * 8 / 1e6 / world
    what = ("pressure Poisson solve alone (pcg.rs:14-82, 100 CG iterations on the right-hand side of step 20)" if poisson
            else "advect + pressure projection (dec_fluid.rs loop body)")
    return {"workload": f"2D smoke plume {n}x{n} MAC grid, {what}",
            "grid": [n, n], "cg_iterations_per_step": iters, "cg_max_iterations": 100, "threshold": 0.1, "timestep": 0.05,
            "options": list(opts),
            "parallelism": f"slab{world} (rows split over {world} GPUs, halos + reductions over NVLink peer memory)" if world > 1 else "single",
            "l2": ("flushed before every timed step (302 MB memset); within a step the CG re-reads its working set"
                   if 10 * cells_mb < 126.0 else "not flushed: every rank's fields are larger than the 126 MB L2")}


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


def median(xs):
    s = sorted(xs)
    m = len(s) // 2
    return s[m] if len(s) % 2 else 0.5 * (s[m - 1] + s[m])


# ----------------------------------------------------------------------------------------------- CPU arm
def host_threads():
    """All the host threads the port may use; set explicitly because torchrun exports OMP_NUM_THREADS=1."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def cpu_steps(n, mode, budget_s, warm=1, max_steps=4, iters=None):
    """Real steps of the CPU oracle on the n^2 smoke plume from the zero state (at n >= 512 every step runs the
    full 100 CG iterations, so the first steps cost what later ones do).  Returns (seconds per step, steps, warm-ups, threads)."""
    from oracle import pano_oracle as O
    threads = O.set_threading(mode, host_threads())
    prm = O.smoke_params(n)
    if iters is not None:
        prm = dict(prm, max_iterations=iters)
    S = O.FluidState(**prm)
    for _ in range(warm):
        S.step()                                   # page faults, first touch
    t0 = time.perf_counter()
    k = 0
    while k < max_steps and (k == 0 or (time.perf_counter() - t0) * (k + 1) / k < budget_s):
        S.step()
        k += 1
    dt = (time.perf_counter() - t0) / k
    S.close()
    O.set_threading(O.SERIAL)
    return dt, k, warm, (1 if mode == O.SERIAL else threads)


def cpu_baseline_sample(budget_faithful=12.0, budget_parallel=8.0):
    """cpu_baseline of the GPU arm: a bounded sample of the same workload.  A full 8192^2 step takes the port with the
    reference's threading about a minute, so the sample is the same plume at 2048^2 (every step runs the full 100 CG
    iterations there as well; the per-cell cost of these streaming passes does not depend on the grid once the fields
    -- 12 x 33 MB -- have left the caches)."""
    from oracle import pano_oracle as O
    n = 2048
    dt, k, w, cores = cpu_steps(n, O.REFERENCE_FAITHFUL, budget_faithful)
    dt2, k2, w2, cores2 = cpu_steps(n, O.ALL_PARALLEL, budget_parallel)
    sample = f"{k} full steps of the {n}^2 smoke plume after {w} warm-up step (100 CG iterations each); bounded sample of the benched workload"
    return {"value": n * n / dt / 1e6, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
            "threading": "reference-faithful: threads only in the 3 derivative passes the reference runs under rayon",
            "all_parallel": {"value": n * n / dt2 / 1e6, "cores": cores2,
                             "sample": f"{k2} full steps of the {n}^2 smoke plume after {w2} warm-up step, every pass OpenMP-parallel"}}


def run_reference(args, rank, world):
    if rank != 0:
        return 0
    from oracle import pano_oracle as O
    poisson = args.poisson_only
    n = args.n or (POISSON_GRID if poisson else DEFAULT_GRID)
    mode = O.ALL_PARALLEL if args.cpu_variant == "parallel" else O.REFERENCE_FAITHFUL
    cells = n * n
    kind_note = ("C port of the Rust reference (rustc/cargo absent); threading as in the reference: rayon only in the "
                 "three derivative passes, everything else serial") if mode == O.REFERENCE_FAITHFUL else \
        "C port of the Rust reference; every pass OpenMP-parallel (upper bound for the CPU)"
    budget = float(os.environ.get("PANO_REF_BUDGET_S", "150"))
    est = cells * (0.8e-6 if mode == O.REFERENCE_FAITHFUL else 0.07e-6)        # crude, only picks the strategy
    extrapolated = False
    if est <= budget:
        # REAL full steps of the benched grid: one warm-up (first touch) and as many timed steps as the budget holds
        want_w = max(1, min(args.warmup, 1 if est > 5.0 else 3))
        dt, steps, warm, threads = cpu_steps(n, mode, budget - est * want_w, warm=want_w, max_steps=max(1, args.steps))
        ms = dt * 1e3
        sample = (f"{steps} real full steps of the {n}^2 smoke plume after {warm} warm-up step(s), 100 CG iterations each "
                  f"(requested --steps {args.steps} --warmup {args.warmup}; bounded by PANO_REF_BUDGET_S={budget:.0f} s)")
    else:
        # even one step exceeds the budget: time the step with the CG capped at 4 and at 12 iterations and extrapolate
        # linearly (every CG iteration does identical work); the line says so
        t4, _, _, threads = cpu_steps(n, mode, 1e9, warm=0, max_steps=1, iters=4)
        t12, _, _, threads = cpu_steps(n, mode, 1e9, warm=0, max_steps=1, iters=12)
        per_iter = (t12 - t4) / 8.0
        ms = (t4 + 96.0 * per_iter) * 1e3
        steps, warm, extrapolated = 1, 0, True
        sample = (f"{n}^2: one step timed with the CG capped at 4 and at 12 iterations, extrapolated linearly to the "
                  f"100-iteration step ({per_iter * 1e3:.1f} ms per iteration)")
    value = cells / (ms * 1e-3) / 1e6
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": warm, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(n, args.gpus, args.opt, poisson=poisson),
            "device": "host CPU", "extrapolated": extrapolated,
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                             "note": kind_note},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), file=_JSON_OUT, flush=True)
    return 0


# ----------------------------------------------------------------------------------------------- GPU arm
class _Single:
    """One GPU: device-resident fields stepped by pano_fluid_step; e2e through pano_fluid_step_host."""

    def __init__(self, ctx, n, fluid, _lib):
        self.ctx, self.n, self._lib = ctx, n, _lib
        self.sim = fluid.DecFluid(**fluid.smoke_params(n), ctx=ctx)
        self.cells = n * n
        self.n1 = (n + 1) * n + n * (n + 1)
        self.nsteps = 0

    def step(self):
        self.sim.step(want_info=False)
        self.nsteps += 1

    def info(self):
        self.ctx.sync()
        self.nsteps += 1
        return self.sim.step(want_info=True)

    def solve(self):
        """The pressure solve alone on the right-hand side of the last step (it still sits in `temp`)."""
        s = self.sim
        L = self._lib.load()
        self._lib.check(L.pano_pcg_solve(self._lib.PRECOND_IDENTITY, s.pressure.handle, s.temp.handle, s.params.max_iterations,
                                         s.params.threshold, s.residual.handle, s.auxiliary.handle, s.search.handle,
                                         s.params.timestep, s.params.obstacle, None))

    def e2e_setup(self, np):
        L, _lib = self._lib.load(), self._lib
        self.bufs = []
        for count in (self.cells, self.n1):
            p = C.c_void_p()
            _lib.check(L.pano_host_alloc(count * 8, C.byref(p)))
            self.bufs.append(p)
        d = np.ctypeslib.as_array(C.cast(self.bufs[0], C.POINTER(C.c_double)), shape=(self.cells,))
        v = np.ctypeslib.as_array(C.cast(self.bufs[1], C.POINTER(C.c_double)), shape=(self.n1,))
        d[:] = self.sim.density.view_linear()
        v[:] = self.sim.vel.view_linear()
        self.h2d, self.d2h = (self.cells + self.n1) * 8, (self.cells + self.n1) * 8
        self.call = ("pano_fluid_step_host (pinned host density + velocity in, density + velocity out, every step; the pressure "
                     "is scratch the example never reads between steps and stays on the device)")

    def e2e_step(self):
        L, _lib = self._lib.load(), self._lib
        _lib.check(L.pano_fluid_step_host(self.ctx.handle, C.byref(self.sim.params), self.n, self.n, self.bufs[0], self.bufs[1],
                                          None, None))

    def e2e_teardown(self):
        L = self._lib.load()
        for p in self.bufs:
            L.pano_host_free(p)

    def close(self):
        self.sim = None


class _Slab:
    """N GPUs: this rank's slab of the grid (pano_dist_*); e2e = upload my rows, step, download my rows."""

    def __init__(self, ctx, n, rank, world, dist_t, fluid, _lib):
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
        self.nsteps = 0

    def step(self):
        self.D.step()
        self.nsteps += 1

    def info(self):
        self.D.step()
        self.nsteps += 1
        return self.D.sync()

    def solve(self):
        self.D.solve()

    def e2e_setup(self, np):
        L, _lib, D, dist = self._lib.load(), self._lib, self.D, self.dist
        self.bufs, self.counts = [], []
        for which in (dist.DENSITY, dist.VY, dist.VX):
            count = D._rows(which) * D._pitch(which)
            p = C.c_void_p()
            _lib.check(L.pano_host_alloc(count * 8, C.byref(p)))
            self.bufs.append(p)
            self.counts.append(count)
            _lib.check(L.pano_dist_download(D._h, which, self.bufs[which], None))
        self.h2d, self.d2h = sum(self.counts) * 8, sum(self.counts) * 8       # bytes of THIS rank
        self.call = ("pano_dist_step_host (pinned host rows of this rank in and out, every step: density, vy, vx; the density comes "
                     "back while the solver runs)")

    def e2e_step(self):
        self.D.step_host(self.bufs[self.dist.DENSITY], self.bufs[self.dist.VY], self.bufs[self.dist.VX])

    def e2e_teardown(self):
        L = self._lib.load()
        for p in self.bufs:
            L.pano_host_free(p)

    def close(self):
        self.D.close()


def cg_kernel_name(n, world, cells_per_gpu, opts):
    """The kernel the library's auto choice lands on (csrc/pano_cg.cu, pano_dist.cu), given the options of this run."""
    o = dict(kv.split("=") for kv in opts)
    sr = int(o.get("cg_single_reduction", -1))       # -1 auto: single-reduction kernels on chip and on slabs, two-reduction streaming on one GPU
    if world == 1 and n <= 1024 and -(-n // 8) * n <= 5120:
        return "k_cg_cluster"                      # one thread-block cluster (csrc/pano_cg_cluster.cu): grids up to ~40 k cells
    if world == 1 and cells_per_gpu <= 1_200_000:
        return "k_cg_resident_sr" if sr != 0 else "k_cg_resident2"
    if world > 1:
        return "k_cg_sr" if sr != 0 else "k_cg_stream"
    return "k_cg_sr" if sr > 0 else "k_cg_stream"


def timed_laps(ctx, work, K, flush=None):
    """K calls of work(), each one lap between two CUDA events on the library's stream, no host synchronisation in
    between.  flush (small grids): evict the fields from L2 before every lap; the flush gets its own, discarded, lap."""
    ctx.timer_mark()
    for _ in range(K):
        if flush is not None:
            flush.fill(0.0)
            ctx.timer_mark()
        work()
        ctx.timer_mark()
    laps = ctx.timer_marks_ms()
    return laps[1::2] if flush is not None else laps


def kernel_table(phase_ms, phase_steps, cells, iters, peak, traffic, grid_key):
    """Per-kernel rooflines from the per-phase CUDA events of pano_fluid_step (option step_timing)."""
    names = ["inflow", "advect_all", "neg_divergence", "cg", "project"]
    alg = {"advect_all": BYTES_ADVECT, "neg_divergence": BYTES_NEGDIV, "cg": BYTES_CG_INIT_FUSED + BYTES_CG_ITER_FUSED * iters, "project": BYTES_PROJECT}
    out = {}
    for nm, ms in zip(names, phase_ms):
        ms = ms / max(1, phase_steps)
        row = {"ms": ms}
        if nm in alg and ms > 0:
            b = cells * alg[nm]
            gbs = b / (ms * 1e-3) / 1e9
            row.update({"algorithmic_bytes": b, "algorithmic_gbs": gbs, "frac_of_measured_peak": gbs / peak, "frac_of_8000": gbs / NOMINAL_GBS})
            if nm == "cg":
                row["survey_88_byte_model_gbs"] = cells * (BYTES_CG_INIT + BYTES_CG_ITER * iters) / (ms * 1e-3) / 1e9
            us = traffic.get(f"{nm}_{grid_key}_ncu_us")
            if us:                                 # the kernel's own duration under ncu (no launch gap), from the committed capture
                row.update({"ncu_us": us, "algorithmic_gbs_on_ncu_duration": b / (us * 1e-6) / 1e9,
                            "frac_of_8000_on_ncu_duration": b / (us * 1e-6) / 1e9 / NOMINAL_GBS})
            t = traffic.get(f"{nm}_{grid_key}")
            row["ncu_dram_bytes"] = t
            if t:
                row.update({"dram_gbs": t / (ms * 1e-3) / 1e9, "dram_frac_of_measured_peak": t / (ms * 1e-3) / 1e9 / peak,
                            "dram_frac_of_8000": t / (ms * 1e-3) / 1e9 / NOMINAL_GBS})
        out[nm] = row
    return out


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

    poisson = args.poisson_only
    n = args.n or (POISSON_GRID if poisson else DEFAULT_GRID)
    ctx = P.Context(local_rank)
    for kv in args.opt:
        k, v = kv.split("=")
        ctx.set_option(k, int(v))
    peak, peak_src = load_peaks()
    traffic = load_traffic()

    def make_job(grid):
        return _Slab(ctx, grid, rank, world, dist_t, fluid, _lib) if multi else _Single(ctx, grid, fluid, _lib)

    def barrier():
        ctx.sync()
        if multi:
            dist_t.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if not multi:
            return x
        t = torch.tensor(x if isinstance(x, list) else [x], device="cuda", dtype=torch.float64)
        dist_t.all_reduce(t, op=dist_t.ReduceOp.MAX)
        return t.tolist() if isinstance(x, list) else float(t.item())

    def sum_over_ranks(x):
        if not multi:
            return x
        t = torch.tensor([float(x)], device="cuda", dtype=torch.float64)
        dist_t.all_reduce(t)
        return float(t.item())

    job = make_job(n)
    cells = n * n
    # 20 warm-up steps (PANO_BENCH_MIN_WARMUP lowers the floor for profiler runs, where every launch is serialised)
    K, W = max(1, args.steps), max(args.warmup, env_int("PANO_BENCH_MIN_WARMUP", 20) if not poisson else 3)
    fields_mb = cells * 8 / 1e6 / world
    l2_resident = 10 * fields_mb < 126.0
    flush = P.Grid2d((6144, 6144), ctx).new_simplex_2() if l2_resident else None     # 302 MB > 126 MB L2

    # ---- warm-up (20 steps); the last one reads the solver info, which also tells the iteration count
    for _ in range(W - 1):
        job.step()
    info = job.info()
    work = job.step
    if poisson:                                    # configs[4]: the solve alone, again and again on the rhs of step W
        work = job.solve
        for _ in range(2):
            work()
    barrier()

    # ---- timed region: K laps on the library's stream (CUDA events), max over ranks lap by lap
    ctx.set_option("step_timing", 1)
    ctx.step_times()
    launches0 = ctx.launch_count()
    with ClockSampler(local_rank) as clk:
        barrier()
        laps = timed_laps(ctx, work, K, flush)
        barrier()
    launches = int(sum_over_ranks(ctx.launch_count() - launches0))
    phase_ms, phase_steps = ctx.step_times()
    ctx.set_option("step_timing", 0)
    laps = max_over_ranks(list(laps))
    total_ms = sum(laps)
    ms_per_step = total_ms / K
    value = cells * K / (total_ms * 1e-3) / 1e6
    last = job.info() if not poisson else (job.D.sync() if multi else None)
    if poisson and not multi:
        ctx.sync()
        last = info
    iters = last["applies"] if last else 100

    # ---- N > 1: the same steps by rank 0 ALONE on its GPU -> the like-for-like one-GPU number AND a field-by-field check
    # of the multi-process path (CUDA-IPC windows, in-kernel halo stores, cross-rank reductions) on this very box
    same_1gpu, parity = None, None
    if multi and not args.no_single and not poisson:
        from panopaea_b200 import dist as pdist
        mine = [job.D.download(w_) for w_ in (pdist.DENSITY, pdist.VY, pdist.VX, pdist.PRESSURE)]
        solo_fields, solo_info = None, None
        if rank == 0:
            solo = _Single(ctx, n, fluid, _lib)
            ks = max(2, min(K, 5))
            for _ in range(job.nsteps - 1 - ks):
                solo.step()
            ctx.sync()
            solo_laps = timed_laps(ctx, solo.step, ks)
            solo_info = solo.info()
            assert solo.nsteps == job.nsteps
            vy, vx = solo.sim.vel.split()
            solo_fields = [solo.sim.density.to_host(), vy, vx, solo.sim.pressure.to_host()]
            solo_ms = sum(solo_laps) / ks
            same_1gpu = {"value": cells / (solo_ms * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": solo_ms, "median_ms_per_step": median(solo_laps),
                         "steps": ks, "warmup": job.nsteps - 1 - ks,
                         "note": "the same grid stepped by rank 0 alone (pano_fluid_step) for the same number of steps from the zero state"}
            solo.close()
            del solo
        # every rank's rows travel to rank 0 over NCCL; rank 0 compares them with its own full fields
        worst = {}
        names = ["density", "vy", "vx", "pressure"]
        if rank == 0:
            for r in range(world):
                y0, y1 = pdist.slab_range(n, r, world)
                for fi, nm in enumerate(names):
                    rows = (y1 - y0) + (1 if nm == "vy" and r == world - 1 else 0)
                    if r == 0:
                        got = mine[fi]
                    else:
                        buf = torch.empty((rows, n + 1 if nm == "vx" else n), device="cuda", dtype=torch.float64)
                        dist_t.recv(buf, src=r)
                        got = buf.cpu().numpy()
                        del buf
                    want = solo_fields[fi][y0:y0 + rows]
                    scale = max(float(np.abs(solo_fields[fi]).max()), 1e-300)
                    worst[nm] = max(worst.get(nm, 0.0), float(np.abs(got - want).max()) / scale)
            parity = {"max_rel_density": worst["density"], "max_rel_vy": worst["vy"], "max_rel_vx": worst["vx"], "max_rel_p": worst["pressure"],
                      "iterations_equal": bool(solo_info["iterations"] == last["iterations"] and solo_info["applies"] == last["applies"]),
                      "iterations": [last["iterations"], solo_info["iterations"]], "steps_compared": job.nsteps, "tolerance": PARITY_TOL,
                      "ok": bool(max(worst.values()) <= PARITY_TOL and solo_info["iterations"] == last["iterations"]),
                      "note": "all ranks' rows of density / vy / vx / pressure after the same number of steps from the zero state, relative to max|field|"}
            del solo_fields
        else:
            for a in mine:
                dist_t.send(torch.from_numpy(np.ascontiguousarray(a)).cuda(), dst=0)
        del mine
        barrier()

    # ---- e2e: the same step with the fields in (pinned) host memory, H2D + D2H inside every step
    e2e = None
    if not poisson:
        job.e2e_setup(np)
        e2e_steps = max(3, min(K, 5 if n >= 8192 else 10))
        for _ in range(2):
            job.e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            job.e2e_step()
        barrier()
        e2e_s = max_over_ranks(time.perf_counter() - t0)
        e2e = {"value": cells * e2e_steps / e2e_s / 1e6, "unit": UNIT, "steps": e2e_steps, "h2d_bytes_per_step": int(sum_over_ranks(job.h2d)),
               "d2h_bytes_per_step": int(sum_over_ranks(job.d2h)), "call": job.call}
        job.e2e_teardown()

    # ---- roofline of the dominant kernel (the persistent CG kernel: phase 3), for this rank's share of the grid
    my_cells = cells / world
    if poisson:
        cg_ms = ms_per_step
    else:
        cg_ms = max_over_ranks(phase_ms[3] / max(1, phase_steps))
    cg_bytes88 = my_cells * (BYTES_CG_INIT + BYTES_CG_ITER * iters)
    cg_bytes = my_cells * (BYTES_CG_INIT_FUSED + BYTES_CG_ITER_FUSED * iters)
    achieved = cg_bytes / (cg_ms * 1e-3) / 1e9 if cg_ms > 0 else 0.0
    achieved88 = cg_bytes88 / (cg_ms * 1e-3) / 1e9 if cg_ms > 0 else 0.0
    step_bytes = my_cells * (BYTES_ADVECT + BYTES_NEGDIV + BYTES_PROJECT + BYTES_CG_INIT + BYTES_CG_ITER * iters)
    kernel = cg_kernel_name(n, world, my_cells, args.opt)
    dram = traffic.get(f"cg_{n}_x{world}")
    roofline = {"bound": "hbm", "kernel": "%s (persistent CG: init + %d iterations in one launch%s)" % (kernel, iters, ", per GPU" if multi else ""),
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "peak_source": peak_src, "traffic": dram,   # ncu dram bytes per launch (profiles/roofline_traffic.json), null if not captured
                "frac_of_8000": achieved / NOMINAL_GBS,
                "bytes_per_cell_iteration": BYTES_CG_ITER_FUSED,
                "survey_88_byte_model": {"bytes_per_cell_iteration": BYTES_CG_ITER, "algorithmic_bytes_per_launch": cg_bytes88,
                                         "achieved": achieved88, "frac": achieved88 / peak, "frac_of_8000": achieved88 / NOMINAL_GBS},
                "dram_achieved": (dram / (cg_ms * 1e-3) / 1e9) if dram and cg_ms > 0 else None,
                "dram_frac": (dram / (cg_ms * 1e-3) / 1e9 / peak) if dram and cg_ms > 0 else None,
                "dram_frac_of_8000": (dram / (cg_ms * 1e-3) / 1e9 / NOMINAL_GBS) if dram and cg_ms > 0 else None,
                "algorithmic_bytes_per_launch": cg_bytes, "kernel_ms": cg_ms,
                "kernel_share_of_step": cg_ms / ms_per_step if ms_per_step > 0 else None,
                "step_achieved_gbs": None if poisson else step_bytes / (ms_per_step * 1e-3) / 1e9,
                "phase_ms": None if poisson else dict(zip(["inflow", "advect_all", "neg_divergence", "cg", "project"],
                                                          [m / max(1, phase_steps) for m in phase_ms])),
                "note": ("`achieved`/`frac`: algorithmic bytes of the fused kernel, 8 + 64 B per cell and CG iteration (r, s, p, x each read once "
                         "and written once; DESIGN.md 4) / kernel time / peak.  SURVEY.md 8(d) models the iteration as three passes with z "
                         "stored (32 + 88 B): `survey_88_byte_model` -- above 1 by construction, because the fusion removes a quarter of that "
                         "model's traffic.  `traffic` = ncu dram bytes of the same kernel on the same grid (profiles/roofline_traffic.json), "
                         "`dram_frac` = traffic / time / peak.  Per-GPU working set 10 fields x %.1f MB %s the 126 MB L2%s"
                         % (fields_mb, "fits in" if l2_resident else "exceeds",
                            "; the whole CG state stays in shared memory / registers for the solve, so this is NOT an HBM measurement" if my_cells <= 1_200_000 and not multi else ""))}

    extra = {}
    if not multi and not args.no_extra and not poisson:
        job.close()
        del job
        job = None
        # ---- configs[3]: per-kernel rooflines at 4096^2, the grid BASELINE quotes the >= 70 % target on
        j4 = _Single(ctx, 4096, fluid, _lib)
        for _ in range(5):
            j4.step()
        i4 = j4.info()
        ctx.set_option("step_timing", 1)
        ctx.step_times()
        laps4 = timed_laps(ctx, j4.step, 10)
        pm4, ps4 = ctx.step_times()
        ctx.set_option("step_timing", 0)
        k4 = kernel_table(pm4, ps4, 4096 * 4096, i4["applies"], peak, traffic, "4096")
        extra["kernels_4096"] = {"value": 4096 * 4096 / (median(laps4) * 1e-3) / 1e6, "unit": UNIT, "median_ms_per_step": median(laps4),
                                 "steps": 10, "warmup": 6, "cg_iterations_per_step": i4["applies"], "kernels": k4,
                                 "peak_measured_gbs": peak, "peak_nominal_gbs": NOMINAL_GBS,
                                 "note": "per-kernel device time from CUDA events between the kernels of pano_fluid_step (option step_timing); "
                                         "algorithmic bytes per SURVEY.md 8(d); ncu_dram_bytes from profiles/roofline_traffic.json (same kernel, same grid)"}
        # the multigrid-preconditioned step behind the pcg.rs trait seam (beyond the reference; reported aside)
        if not args.no_mg:
            sim = j4.sim
            sim.params.precond = _lib.PRECOND_MULTIGRID
            for _ in range(3):
                sim.step(want_info=False)
            ctx.sync()
            mg_laps = timed_laps(ctx, lambda: sim.step(want_info=False), 5)
            mg_info = sim.step(want_info=True)
            sim.params.precond = _lib.PRECOND_IDENTITY
            extra["multigrid_pcg_step_4096"] = {"value": 4096 * 4096 / (median(mg_laps) * 1e-3) / 1e6, "unit": UNIT, "median_ms_per_step": median(mg_laps),
                                                "cg_iterations_per_step": mg_info["applies"], "final_residual": mg_info["final_residual"],
                                                "note": "pano_step_params.precond = PANO_PRECOND_MULTIGRID (DESIGN.md 5b): converged solves, host-driven PCG loop"}
        j4.close()
        del j4
        # ---- configs[1] and configs[0]: the on-chip regimes, L2 flushed before every lap
        fl = P.Grid2d((6144, 6144), ctx).new_simplex_2()
        for g_ in (1024, 128):
            js = _Single(ctx, g_, fluid, _lib)
            for _ in range(20):
                js.step()
            isml = js.info()
            lp = timed_laps(ctx, js.step, 20, fl)
            extra[f"config_{g_}"] = {"value": g_ * g_ / (median(lp) * 1e-3) / 1e6, "unit": UNIT, "median_ms_per_step": median(lp),
                                     "steps": 20, "warmup": 21, "cg_iterations_per_step": isml["applies"],
                                     "kernel": cg_kernel_name(g_, 1, g_ * g_, args.opt), "l2": "flushed before every timed step"}
            js.close()
            del js
        del fl
        # ---- Grid3d (the reference has no 3-D fluid path, so configs[3] is quoted on its 2-D branch above):
        # the 256^3 smoke plume through pano_fluid3_step, per-kernel algorithmic rooflines
        if not args.no_3d:
            from panopaea_b200 import grid3
            n3 = 256
            s3 = grid3.DecFluid3(**grid3.smoke_params(n3), ctx=ctx)
            for _ in range(20):
                s3.step(want_info=False)
            i3 = s3.step()
            ctx.set_option("step_timing", 1)
            ctx.step_times()
            ap3 = []
            lp3 = timed_laps(ctx, lambda: ap3.append(s3.step()["applies"]), 10)   # the info of every timed step: the count drifts
            pm3, ps3 = ctx.step_times()
            ctx.set_option("step_timing", 0)
            i3b = {"applies": sum(ap3) / len(ap3)}
            c3 = n3 ** 3
            b3 = {"advect_all": 64, "neg_divergence": 32, "cg": 8 + 64 * i3b["applies"], "project": 56}
            k3 = {}
            for nm, ms in zip(["inflow", "advect_all", "neg_divergence", "cg", "project"], pm3):
                ms = ms / max(1, ps3)
                row = {"ms": ms}
                if nm in b3 and ms > 0:
                    gbs = c3 * b3[nm] / (ms * 1e-3) / 1e9
                    row.update({"algorithmic_bytes": c3 * b3[nm], "algorithmic_gbs": gbs, "frac_of_measured_peak": gbs / peak, "frac_of_8000": gbs / NOMINAL_GBS})
                k3[nm] = row
            extra["grid3d_256"] = {"value": c3 / (median(lp3) * 1e-3) / 1e6, "unit": UNIT, "median_ms_per_step": median(lp3), "steps": 10, "warmup": 21,
                                   "grid": [n3, n3, n3], "cg_iterations_per_step": i3b["applies"], "cg_applies_per_timed_step": ap3,
                                   "cg_info_before_timing": i3, "kernels": k3,
                                   "kernel": "k3_cg_tile (persistent 7-point CG: ring of 3-D TMA boxes, dynamic tiles)",
                                   "note": "3-D smoke plume, 7-point Laplacian, identity preconditioner, converging solves (threshold 0.1); parity "
                                           "unpinned: the reference holds only the struct Grid3d and the unused trilinear (DESIGN.md 5c)"}
            del s3

    # ---- configs[4]: the pressure Poisson solve alone at 16384^2 on these N GPUs and on rank 0 alone
    if multi and not args.no_poisson and not poisson:
        barrier()
        if job is not None:
            job.close()
            del job
            job = None
        barrier()
        extra["poisson_16384"] = poisson_leg(ctx, POISSON_GRID, rank, world, dist_t, fluid, _lib, barrier, max_over_ranks, args)

    if rank == 0:
        cpu = None
        if not args.no_cpu and not multi and not poisson:
            cpu = cpu_baseline_sample()
        line = {"metric": METRIC if not poisson else "Mcell-solves/s (pressure Poisson solve only, 100 CG iterations)",
                "value": value, "unit": UNIT if not poisson else "Mcell-solves/s", "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms_per_step, "median_ms_per_step": median(laps), "min_ms_per_step": min(laps), "max_ms_per_step": max(laps),
                "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": workload_config(n, world, args.opt, iters, poisson),
                "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu, "clocks": clk.summary(),
                "cg_info_last_step": last, "cg_info_warm_step": info,
                "timing": "K laps between CUDA events on the library's stream, no host synchronisation in between; per lap the max over ranks; "
                          "value = cells*K / sum(laps)"}
        line.update(extra)
        if same_1gpu is not None:
            line["one_gpu_same_workload"] = same_1gpu
        if parity is not None:
            line["parity_vs_one_gpu"] = parity
        print(json.dumps(line), file=_JSON_OUT, flush=True)
    rc = 0
    if multi:
        ok = torch.tensor([1.0 if (parity is None or parity["ok"]) else 0.0], device="cuda", dtype=torch.float64)
        dist_t.broadcast(ok, src=0)
        rc = 0 if ok.item() > 0.5 else 3
        dist_t.barrier()
        dist_t.destroy_process_group()
    if rc:
        log("bench.py: parity_vs_one_gpu FAILED (see the JSON line); exit code 3")
    return rc


def poisson_leg(ctx, n, rank, world, dist_t, fluid, _lib, barrier, max_over_ranks, args):
    """BASELINE configs[4] / SURVEY.md 8(d) config 5: the pressure solve alone (fixed 100 CG iterations on the right-hand side
    of step 20) at 16384^2, on the N GPUs and -- same job, same box -- on rank 0 alone."""
    cells = n * n
    job = _Slab(ctx, n, rank, world, dist_t, fluid, _lib)
    for _ in range(19):
        job.step()
    info = job.info()                              # step 20
    for _ in range(2):
        job.solve()
    barrier()
    ks = 5
    laps = max_over_ranks(list(timed_laps(ctx, job.solve, ks)))
    pinfo = job.D.sync()
    barrier()
    job.close()
    del job
    out = {"grid": [n, n], "n_gpus": world, "solves": ks, "cg_iterations": pinfo["applies"], "rhs": "step 20 of the smoke plume from the zero state",
           "ms_per_solve": sum(laps) / ks, "median_ms_per_solve": median(laps),
           "value": cells / (median(laps) * 1e-3) / 1e6, "unit": "Mcell-solves/s",
           "final_residual": pinfo["final_residual"], "step20_info": info}
    if not args.no_single:
        solo_ms = None
        if rank == 0:
            solo = _Single(ctx, n, fluid, _lib)
            for _ in range(19):
                solo.step()
            sinfo = solo.info()
            for _ in range(2):
                solo.solve()
            ctx.sync()
            sl = timed_laps(ctx, solo.solve, 3)
            solo_ms = median(sl)
            out["one_gpu"] = {"median_ms_per_solve": solo_ms, "value": cells / (solo_ms * 1e-3) / 1e6, "unit": "Mcell-solves/s",
                              "solves": 3, "step20_info": sinfo}
            out["speedup_vs_one_gpu"] = solo_ms / median(laps)
            solo.close()
            del solo
        barrier()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", "--grid", dest="n", type=int, default=0, help="grid size override (multiple of 128); use --grid under torchrun, whose own parser claims --n")
    ap.add_argument("--poisson-only", action="store_true", help="BASELINE configs[4]: time the pressure solve alone (default grid 16384)")
    ap.add_argument("--cpu-variant", default="faithful", choices=["faithful", "parallel"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-mg", action="store_true", help="skip the multigrid-preconditioned step timing")
    ap.add_argument("--no-single", action="store_true", help="N > 1: skip the one-GPU run of the same grid (and the parity check against it)")
    ap.add_argument("--no-extra", action="store_true", help="N = 1: skip the 4096^2 / 1024^2 / 128^2 legs")
    ap.add_argument("--no-3d", dest="no_3d", action="store_true", help="N = 1: skip the 256^3 Grid3d leg")
    ap.add_argument("--no-poisson", action="store_true", help="N > 1: skip the 16384^2 Poisson-only leg")
    ap.add_argument("--opt", action="append", default=[], help="library option key=value (e.g. cg_kernel=1)")
    args = ap.parse_args()
    rank, world, local_rank = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if args.impl == "reference":
        return run_reference(args, rank, world)
    return run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    sys.exit(main())
