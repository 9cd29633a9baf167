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
    if n > 1100:
        names = ["P1 tiles", "reduce + allreduce 1", "P2 tiles", "reduce + allreduce 2", "-", "-", "-", "-"]
    print(f"n={n} applies={its} total {tot/its:.0f} cycles/iter = {tot/its/1.965e3:.2f} us/iter")
    for nm, v in zip(names, out):
        print(f"   {nm:20s} {v/its:8.0f} cyc/iter  {100*v/tot:5.1f}%")
    if n > 1100:
        import numpy as np
        G = ctx.num_sms()
        arr = (C.c_int64 * (2 * G))()
        _lib.check(L.pano_ctx_cg_profile_ctas(ctx.handle, arr, 2 * G))
        a = np.array(arr[:], dtype=np.float64).reshape(2, G) / its
        for ph in range(2):
            v = a[ph]
            order = np.argsort(v)
            print(f"   P{ph+1} tile-loop cycles per CTA: min {v.min():.0f} p10 {np.percentile(v,10):.0f} median {np.median(v):.0f} p90 {np.percentile(v,90):.0f} max {v.max():.0f}; slowest CTAs {order[-6:].tolist()} fastest {order[:6].tolist()}")
            print("      56-tile CTAs (0..51) mean %.0f, 55-tile CTAs mean %.0f" % (v[:52].mean(), v[52:].mean()))
