import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import panopaea_b200 as P
from panopaea_b200 import fluid, _lib
ctx = P.Context(0)
L = _lib.load()
names = ["mailbox poll", "search update+bar", "stencil", "cta reduce 1", "grid allreduce 1", "x/r update+post", "cta reduce 2", "grid allreduce 2"]
for n in [int(a) for a in sys.argv[1:]] or [1024]:
    sim = fluid.DecFluid(**fluid.smoke_params(n), ctx=ctx)
    for _ in range(20): sim.step(want_info=False)
    ctx.set_option("cg_profile", 1)
    info = sim.step()
    out = (C.c_int64 * 8)()
    _lib.check(L.pano_ctx_cg_profile(ctx.handle, out))
    ctx.set_option("cg_profile", 0)
    its = max(1, info["applies"]); tot = sum(out)
    print(f"n={n} applies={its} total {tot/its:.0f} cycles/iter = {tot/its/1.965e3:.2f} us/iter")
    for nm, v in zip(names, out):
        print(f"   {nm:20s} {v/its:8.0f} cyc/iter  {100*v/tot:5.1f}%")
