#!/usr/bin/env python
"""Per-iteration cost of the CG kernels on small grids: fixed iteration counts (threshold never reached), CUDA events.
usage: scripts/bench_small_cg.py N [N ...]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import panopaea_b200 as P
from panopaea_b200 import fluid, pcg

ctx = P.Context(0)
for n in [int(a) for a in sys.argv[1:]] or [128]:
    g = P.Grid2d((n, n), ctx)
    k = max(1, n // 128)
    ob = (70 * n // 128, 80 * n // 128, 50 * n // 128, 70 * n // 128)
    rng = np.random.default_rng(0)
    p = g.new_simplex_2(); p.upload(rng.normal(size=(n, n)))
    b = g.new_simplex_2(); fluid.laplacian_apply(b, p, 0.05, ob)
    x, r, aux, s = (g.new_simplex_2() for _ in range(4))
    for kern, name in ((5, "cluster"), (3, "resident"), (1, "generic")):
        ctx.set_option("cg_kernel", kern)
        res = {}
        try:
            for iters in (10, 110):
                for _ in range(3):
                    pcg.solve_grid_laplacian(x, b, iters, 1e-300, r, aux, s, 0.05, ob, want_info=False)
                ctx.sync(); ctx.timer_start()
                for _ in range(20):
                    pcg.solve_grid_laplacian(x, b, iters, 1e-300, r, aux, s, 0.05, ob, want_info=False)
                res[iters] = ctx.timer_stop_ms() / 20
            per = (res[110] - res[10]) / 100 * 1e3
            print(f"n={n} {name:9s}: {per:6.2f} us/iteration, fixed {res[10]*1e3 - 10*per:6.1f} us  (10 it: {res[10]*1e3:.1f} us, 110 it: {res[110]*1e3:.1f} us)")
        except Exception as e:
            print(f"n={n} {name}: {e}")
    ctx.set_option("cg_kernel", 0)
