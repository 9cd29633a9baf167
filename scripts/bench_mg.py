#!/usr/bin/env python
"""Multigrid-preconditioned solve vs the reference-faithful 100-iteration CG on the same smoke-plume rhs.
usage: scripts/bench_mg.py N [N ...]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import panopaea_b200 as P
from panopaea_b200 import fluid, pcg

ctx = P.Context(0)
for a in [a for a in sys.argv[1:] if "=" in a]:      # library options, e.g. mg_down_blocks=6
    k, v = a.split("=")
    ctx.set_option(k, int(v))
for n in [int(a) for a in sys.argv[1:] if "=" not in a] or [1024]:
    prm = fluid.smoke_params(n)
    sim = fluid.DecFluid(**prm, ctx=ctx)
    for _ in range(10):
        sim.step(want_info=False)
    ob, g = prm["obstacle"], sim.grid
    b = g.new_simplex_2(); b.assign(sim.temp)
    fluid.neg_divergence(b, sim.vel, ob, want_max=False)      # rhs of the NEXT solve would need the advection; use div of the current field + inflow
    sim.vel.fill_rect(prm["inflow"], 20.0, 1)
    fluid.neg_divergence(b, sim.vel, ob, want_max=False)
    x, r, aux, s = (g.new_simplex_2() for _ in range(4))
    M = pcg.Multigrid(g, 0.05, ob)
    z = g.new_simplex_2()
    for _ in range(3):
        M.apply(z, b)
    ctx.timer_start()
    for _ in range(10):
        M.apply(z, b)
    v_ms = ctx.timer_stop_ms() / 10
    for kind, pre, thr in (("identity, 100-iteration cap", None, 0.1), ("multigrid", M, 0.1), ("multigrid", M, 1e-6)):
        pcg.solve_grid_laplacian(x, b, 100, thr, r, aux, s, 0.05, ob, preconditioner=pre)
        ctx.sync(); t0 = time.perf_counter()
        info = pcg.solve_grid_laplacian(x, b, 100, thr, r, aux, s, 0.05, ob, preconditioner=pre)
        ctx.sync(); ms = (time.perf_counter() - t0) * 1e3
        print(f"n={n} {kind:28s} thr={thr:g}: {ms:8.3f} ms  iterations={info['iterations']} final max|r|={info['final_residual']:.3g} (max|b|={info['rhs_max']:.3g})")
    print(f"n={n} one V-cycle: {v_ms*1e3:.1f} us = {n*n*92/v_ms/1e6:.0f} GB/s on the 92 B/cell of level 0; levels={M.levels()}")
    M.close()
