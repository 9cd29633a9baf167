import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import panopaea_b200 as P
from panopaea_b200 import dist, fluid
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
h, w = 256, 256
k = 2
prm = dict(timestep=0.05, threshold=0.1, max_iterations=100, inflow=(5 * k, 20 * k, 27 * k, 32 * k), inflow_density=1.0,
           inflow_vy=20.0, obstacle=(70 * k, 80 * k, 25 * k, 35 * k))
ctxs = [P.Context(0) for _ in range(n)]
sms = ctxs[0].num_sms()
ranks = [dist.DistFluid(ctxs[r], h, w, r, n, prm) for r in range(n)]
ptrs = [r.window()[0] for r in ranks]
for r in ranks:
    r.connect_local(ptrs); r.set_max_ctas(sms // n)
t0 = time.time()
for i, r in enumerate(ranks):
    r.step(); print(f"rank {i} step enqueued at {time.time()-t0:.3f}s", flush=True)
for i, r in enumerate(ranks):
    try:
        print(i, r.sync(), f"{time.time()-t0:.3f}s", flush=True)
    except Exception as e:
        print(i, "ERR", e, f"{time.time()-t0:.3f}s", flush=True)
