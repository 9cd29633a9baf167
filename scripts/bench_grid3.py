#!/usr/bin/env python
"""Times the Grid3d step on the 3-D smoke plume: scripts/bench_grid3.py [N=256] [steps=10] [key=value ...]
Per-phase device time (CUDA events between the kernels), algorithmic bytes and fractions of both peaks."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import panopaea_b200 as P
from panopaea_b200 import grid3

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
ctx = P.Context(0)
for kv in sys.argv[3:]:
    k, v = kv.split("=")
    ctx.set_option(k, int(v))
sim = grid3.DecFluid3(**grid3.smoke_params(n), ctx=ctx)
for _ in range(8):
    info = sim.step()
ctx.set_option("step_timing", 1)
ctx.step_times()
applies = []
for _ in range(steps):
    info = sim.step()                 # the solver info of EVERY timed step: the iteration count drifts as the plume develops
    applies.append(info["applies"])
ms, cnt = ctx.step_times()
ctx.set_option("step_timing", 0)
cells = n ** 3
it = sum(applies) / len(applies)
per = [m / cnt for m in ms]
bytes_per_cell = {"advect_all": 64, "neg_divergence": 32, "cg": 8 + 64 * it, "project": 56}
names = ["inflow", "advect_all", "neg_divergence", "cg", "project"]
out = {"grid": [n, n, n], "cells": cells, "steps": cnt, "cg_info": info, "cg_applies_per_timed_step": applies, "ms_per_step": sum(per),
       "mcell_steps_per_s": cells / sum(per) / 1e3, "phases": {}}
for nm, m in zip(names, per):
    d = {"ms": m}
    if nm in bytes_per_cell:
        gbs = cells * bytes_per_cell[nm] / m / 1e6
        d.update(algorithmic_bytes=cells * bytes_per_cell[nm], algorithmic_gbs=gbs, frac_of_6541=gbs / 6541.5, frac_of_8000=gbs / 8000.0)
    out["phases"][nm] = d
print(json.dumps(out))
