"""compute-sanitizer run of the Grid3d kernels (pano_grid3.cu): the one-pass advection, -div, projection and both CG kernels --
the cp.async plane ring of k3_cg_tile on interior, wall, obstacle and ragged tiles (static first tile + claimed tiles), and the
column kernel -- plus the wide-block -div / projection forms of the 2-D step."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import panopaea_b200 as P
from panopaea_b200 import fluid, grid3
ctx = P.Context(0)
for dims, opts in (((40, 50, 130), dict(cg3_kernel=0)), ((40, 50, 130), dict(cg3_kernel=0, cg3_zc=4)), ((24, 40, 200), dict(cg3_kernel=0, cg_blocks_per_sm=1)),
                   ((17, 20, 33), dict(cg3_kernel=1))):
    for k, v in opts.items():
        ctx.set_option(k, v)
    d, h, w = dims
    sim = grid3.DecFluid3(d, h, w, max_iterations=12, inflow=(d // 3, d // 2, 2, 8, w // 3, w // 2), obstacle=(d // 4, d // 2, h // 2, h // 2 + 4, w // 4, w // 2), ctx=ctx)
    for _ in range(2):
        info = sim.step()
    print("grid3", dims, opts, info, float(np.abs(sim.pressure.to_host()).max()), flush=True)
    for k in opts:
        ctx.set_option(k, 0)
# wide-block -div / projection (default from 1024 columns on; forced here on a small grid)
ctx.set_option("fused_wide", 8)
prm = {k: v for k, v in fluid.smoke_params(256).items() if k not in ("h", "w")}
prm["max_iterations"] = 8
sim2 = fluid.DecFluid(h=256, w=256 + 64, ctx=ctx, **prm)
for _ in range(2):
    info = sim2.step()
print("wide", info, float(np.abs(sim2.vel.view_linear()).max()), flush=True)
