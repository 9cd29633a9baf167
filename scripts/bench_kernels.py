#!/usr/bin/env python
"""Times the non-solver kernels alone on a developed smoke plume: scripts/bench_kernels.py N [key=value,key=value ...]
Each argument after N is one option set applied before timing advect_all (20 launches, CUDA events)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import panopaea_b200 as P
from panopaea_b200 import fluid

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
sets = sys.argv[2:] or [""]
ctx = P.Context(0)
sim = fluid.DecFluid(**fluid.smoke_params(n), ctx=ctx)
for _ in range(12):
    sim.step(want_info=False)
ctx.sync()
cells = n * n
reps = 20
for opts in sets:
    kv = dict(o.split("=") for o in opts.split(",") if o)
    for k, v in kv.items():
        ctx.set_option(k, int(v))
    for _ in range(3):
        fluid.advect_all(sim.temp, sim.vel_temp, sim.density, sim.vel, 0.05)
    ctx.timer_start()
    for _ in range(reps):
        fluid.advect_all(sim.temp, sim.vel_temp, sim.density, sim.vel, 0.05)
    ms = ctx.timer_stop_ms() / reps
    print(f"n={n} advect_all [{opts}] {ms*1e3:.1f} us  {cells*48/ms/1e6:.0f} GB/s (48 B/cell)  = {cells*48/ms/1e6/6521.4:.3f} of the copy peak")
ctx.set_option("advect_kernel", 0)
b = sim.temp
for name, fn, by in (("neg_divergence", lambda: fluid.neg_divergence(b, sim.vel, (70 * n // 128, 80 * n // 128, 50 * n // 128, 70 * n // 128), want_max=False), 24),
                     ("project", lambda: fluid.project(sim.vel, sim.pressure, 0.0), 40)):
    for _ in range(3):
        fn()
    ctx.timer_start()
    for _ in range(reps):
        fn()
    ms = ctx.timer_stop_ms() / reps
    print(f"n={n} {name} {ms*1e3:.1f} us  {cells*by/ms/1e6:.0f} GB/s ({by} B/cell) = {cells*by/ms/1e6/6521.4:.3f} of the copy peak")
