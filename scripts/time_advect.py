#!/usr/bin/env python
"""Time pano_advect_all (auto kernel and the marching kernel) on a list of h x w grids with a smooth flow.
usage: time_advect.py H1xW1 H2xW2 ... [key=value ...]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import panopaea_b200 as P
from panopaea_b200 import fluid
ctx = P.Context(0)
shapes = [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:] if "x" in a and "=" not in a]
for kv in sys.argv[1:]:
    if "=" in kv:
        k, v = kv.split("="); ctx.set_option(k, int(v))
for h, w in shapes:
    g = P.Grid2d((h, w), ctx)
    rng = np.random.default_rng(0)
    q, v, dq, dv = g.new_simplex_2(), g.new_simplex_1(), g.new_simplex_2(), g.new_simplex_1()
    q.upload(rng.uniform(-1, 1, (h, w)))
    v.upload(rng.uniform(-25, 25, g.num_elem_1()))
    for kern in (0, 3):
        ctx.set_option("advect_kernel", kern)
        for _ in range(3):
            fluid.advect_all(dq, dv, q, v, 0.05)
        ctx.sync()
        ctx.timer_mark()
        for _ in range(10):
            fluid.advect_all(dq, dv, q, v, 0.05)
            ctx.timer_mark()
        laps = sorted(ctx.timer_marks_ms())
        med = laps[len(laps) // 2]
        print(f"{h}x{w} advect_kernel={kern}: median {med*1e3:.1f} us  {h*w*48/med/1e6:.0f} GB/s  ({h*w*48/med/1e6/8000:.3f} of 8 TB/s)", flush=True)
    del q, v, dq, dv
