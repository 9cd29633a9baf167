"""compute-sanitizer run of the kernels added in round 2: the TMA-staged advection (interior, border and fallback paths), the
single-reduction CG kernels (streaming with both tile functions, static and dynamic lists; SM-resident), and the fused-halo
multi-rank step with early halo flags and local-done loads (loop-back: two ranks on one GPU)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import panopaea_b200 as P
from panopaea_b200 import fluid, dist
ctx = P.Context(0)
# ---- advection: 160 x 320 has interior tiles, border tiles and ragged ones; three velocity regimes (in box / fallback / far)
g = P.Grid2d((160, 320), ctx)
rng = np.random.default_rng(0)
ctx.set_option("advect_kernel", 4)
for vmax in (30.0, 90.0, 1e300):
    q, v = g.new_simplex_2(), g.new_simplex_1()
    q.upload(rng.uniform(-1, 1, (160, 320)))
    v.upload(rng.uniform(-vmax, vmax, g.num_elem_1()))
    dq, dv = g.new_simplex_2(), g.new_simplex_1()
    fluid.advect_all(dq, dv, q, v, 0.05)
    print("advect_tma vmax", vmax, float(np.abs(dq.to_host()).max()), flush=True)
ctx.set_option("advect_kernel", 0)
# ---- single-reduction CG kernels inside the step
for n, opts in ((256, dict(cg_kernel=6, cg_dynamic=0, cg_sr_exchange=1)), (256, dict(cg_kernel=6, cg_dynamic=1, cg_sr_exchange=1)),
                (256, dict(cg_kernel=6, cg_dynamic=1, cg_sr_exchange=0)), (128, dict(cg_kernel=7))):
    for k, v in opts.items():
        ctx.set_option(k, v)
    prm = {k: v for k, v in fluid.smoke_params(n).items() if k not in ("h", "w")}
    prm["max_iterations"] = 20
    sim = fluid.DecFluid(h=n, w=n, ctx=ctx, **prm)
    for _ in range(2):
        info = sim.step()
    print("single", n, opts, info, float(np.abs(sim.pressure.to_host()).max()), flush=True)
# ---- the fused-halo multi-rank step
k = 2
prm = dict(timestep=0.05, threshold=0.1, max_iterations=10, inflow=(5 * k, 20 * k, 27 * k, 32 * k), inflow_density=1.0, inflow_vy=20.0,
           obstacle=(70 * k, 80 * k, 25 * k, 35 * k))
for opts in (dict(cg_dynamic=0), dict(cg_dynamic=1), dict(cg_dynamic=1, cg_halo_mid=0, cg_early_load=0), dict(cg_dynamic=0, advect_kernel=4)):
    ctxs = [P.Context(0) for _ in range(2)]
    ranks = [dist.DistFluid(ctxs[r], 256, 256, r, 2, prm) for r in range(2)]
    ptrs = [r.window()[0] for r in ranks]
    for r in ranks:
        r.connect_local(ptrs); r.set_max_ctas(ctxs[0].num_sms() // 2)
        for kk, v in opts.items():
            r.ctx.set_option(kk, v)
    for _ in range(2):
        for r in ranks: r.step()
        print("loop-back", opts, [r.sync()["iterations"] for r in ranks], flush=True)
    for r in ranks: r.close()
