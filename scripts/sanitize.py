"""Small run of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import panopaea_b200 as P
from panopaea_b200 import fluid, pcg, dist
ctx = P.Context(0)
for n, kern in ((128, 5), (128, 3), (128, 4), (256, 2), (128, 1)):
    ctx.set_option("cg_kernel", kern)
    sim = fluid.DecFluid(**fluid.smoke_params(n), ctx=ctx) if n == 128 else fluid.DecFluid(h=n, w=n, ctx=ctx, **{k: v for k, v in fluid.smoke_params(256).items() if k not in ("h", "w")})
    for _ in range(2):
        info = sim.step()
    print("kernel", kern, "n", n, info, float(np.abs(sim.pressure.to_host()).max()))
ctx.set_option("cg_kernel", 0)
# multigrid-preconditioned step: fused half-cycle kernels (regular and irregular tiles), single-CTA tail, unfused path
from panopaea_b200 import _lib
for n, fused in ((256, 1), (384, 1), (96, 1), (256, 0)):
    ctx.set_option("mg_fused", fused)
    k = n / 128.0
    sim = fluid.DecFluid(h=n, w=n, ctx=ctx, inflow=(int(5 * k), int(20 * k), int(54 * k), int(64 * k)), obstacle=(int(70 * k), int(80 * k), int(50 * k), int(70 * k)))
    sim.params.precond = _lib.PRECOND_MULTIGRID
    for _ in range(2):
        info = sim.step()
    print("multigrid n", n, "fused", fused, info)
ctx.set_option("mg_fused", 1)
# loop-back multi-rank step
ctxs = [P.Context(0) for _ in range(2)]
k = 2
prm = dict(timestep=0.05, threshold=0.1, max_iterations=20, inflow=(5 * k, 20 * k, 27 * k, 32 * k), inflow_density=1.0, inflow_vy=20.0, obstacle=(70 * k, 80 * k, 25 * k, 35 * k))
ranks = [dist.DistFluid(ctxs[r], 256, 128, r, 2, prm) for r in range(2)]
ptrs = [r.window()[0] for r in ranks]
for r in ranks:
    r.connect_local(ptrs); r.set_max_ctas(ctxs[0].num_sms() // 2)
for _ in range(2):
    for r in ranks: r.step()
    print([r.sync()["iterations"] for r in ranks])
