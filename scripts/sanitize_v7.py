"""compute-sanitizer run of the code paths added in round 1 v7: push all-reduce of the SM-resident CG kernel, light fences in
the streaming kernel, and the halo-flag protocol of the multi-rank step (loop-back: two ranks on one GPU)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import panopaea_b200 as P
from panopaea_b200 import fluid, dist
ctx = P.Context(0)
for n, opts in ((128, dict(cg_kernel=3, cg_push=1)), (128, dict(cg_kernel=3, cg_push=0)), (256, dict(cg_kernel=2, cg_fence=3))):
    for k, v in opts.items():
        ctx.set_option(k, v)
    prm = {k: v for k, v in fluid.smoke_params(128 if n == 128 else 256).items() if k not in ("h", "w")}
    prm["max_iterations"] = 30
    sim = fluid.DecFluid(h=n, w=n, ctx=ctx, **prm)
    for _ in range(2):
        info = sim.step()
    print("single", n, opts, info, float(np.abs(sim.pressure.to_host()).max()), flush=True)
k = 2
prm = dict(timestep=0.05, threshold=0.1, max_iterations=12, inflow=(5 * k, 20 * k, 27 * k, 32 * k), inflow_density=1.0, inflow_vy=20.0,
           obstacle=(70 * k, 80 * k, 25 * k, 35 * k))
for opts in (dict(cg_xflags=1, cg_dynamic=0, cg_halo_first=0), dict(cg_xflags=1, cg_dynamic=1, cg_halo_first=0),
             dict(cg_xflags=1, cg_dynamic=0, cg_halo_first=1), dict(cg_xflags=0, cg_dynamic=0, cg_halo_first=0)):
    ctxs = [P.Context(0) for _ in range(2)]
    ranks = [dist.DistFluid(ctxs[r], 256, 128, r, 2, prm) for r in range(2)]
    ptrs = [r.window()[0] for r in ranks]
    for r in ranks:
        r.connect_local(ptrs); r.set_max_ctas(ctxs[0].num_sms() // 2)
        for kk, v in opts.items():
            r.ctx.set_option(kk, v)
    for _ in range(2):
        for r in ranks: r.step()
        print("loop-back", opts, [r.sync()["iterations"] for r in ranks], flush=True)
    for r in ranks: r.close()
