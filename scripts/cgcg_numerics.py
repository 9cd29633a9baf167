#!/usr/bin/env python
"""Pre-validation (CPU, numpy) of the single-reduction CG (Chronopoulos-Gear) against the reference's CG (pcg.rs:14-82,
as restated by the oracle): same iterates up to rounding?  Run on the smoke-plume right-hand side and on a converging one."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import pano_oracle as O

O.build()
O.set_threading(O.ALL_PARALLEL)


def cgcg(A, b, max_it, thr):
    x = np.zeros_like(b)
    bmax = np.abs(b).max()
    if bmax < thr:
        return -1, x, b.copy(), b.copy()
    r = b.copy()
    w = A(r)
    gamma = float((r * r).sum()); delta = float((w * r).sum())
    alpha = gamma / delta; beta = 0.0
    p = np.zeros_like(b); s = np.zeros_like(b)
    for i in range(max_it):
        p = r + beta * p
        s = w + beta * s
        x = x + alpha * p
        r = r - alpha * s
        if np.abs(r).max() < thr:
            return i, x, r, p
        w = A(r)
        gamma_new = float((r * r).sum()); delta = float((w * r).sum())
        beta = gamma_new / gamma
        alpha = gamma_new / (delta - beta * gamma_new / alpha)
        gamma = gamma_new
    p = r + beta * p
    return max_it, x, r, p


for n, steps in ((512, 6), (1024, 4)):
    prm = O.smoke_params(n)
    S = O.FluidState(**prm)
    for _ in range(steps):
        out = S.step(want_rhs=True)
    b = out["rhs"].copy()
    ob = prm["obstacle"]
    A = lambda v: O.laplacian_closure(n, n, v, 0.05, ob)
    want = O.pcg_grid_laplacian(n, n, b, 100, 0.1, 0.05, ob)
    it, x, r, p = cgcg(A, b, 100, 0.1)
    print(f"plume {n}^2 step {steps}: iterations ref {want.iterations} cgcg {it}; final max|r| ref {want.final_residual:.6g} cgcg {np.abs(r).max():.6g}")
    print("   rel diff x %.3e  r %.3e  s %.3e" % (np.abs(x - want.x).max() / np.abs(want.x).max(), np.abs(r - want.residual).max() / np.abs(b).max(),
                                               np.abs(p - want.search).max() / np.abs(want.search).max()))
    tr = b - A(x)
    print("   true residual drift of cgcg: %.3e (relative to max|b|)" % (np.abs(r - tr).max() / np.abs(b).max()))
    S.close()

# converging systems (consistent rhs), various sizes: iteration counts must agree within +-2
rng = np.random.default_rng(0)
for h, w in ((40, 56), (128, 128), (200, 136), (520, 776)):
    ob = (h // 2, min(h, h // 2 + max(2, h // 12)), w // 3, min(w, w // 3 + max(3, w // 6)))
    b = O.laplacian_closure(h, w, rng.normal(size=(h, w)) * 400.0, 0.05, ob)
    A = lambda v: O.laplacian_closure(h, w, v, 0.05, ob)
    want = O.pcg_grid_laplacian(h, w, b, 100, 0.1, 0.05, ob)
    it, x, r, p = cgcg(A, b, 100, 0.1)
    dx = np.abs(x - want.x).max() / max(1.0, np.abs(want.x).max())
    print(f"consistent {h}x{w}: iterations ref {want.iterations} cgcg {it}; rel diff x {dx:.3e}")
