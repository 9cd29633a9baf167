"""Second, independent restatement of the reference's grid fluid step in numpy.

TEST INFRASTRUCTURE ONLY.  Its sole purpose is to catch transcription errors in
oracle/pano_oracle.c for the functions no reference test pins (advect,
advect_mac, the CG loop, the whole step): the two restatements were written
separately (this one array-at-a-time from the maths in SURVEY.md 8(a), the C
one loop-by-loop from the Rust) and tests/test_oracle_cross.py requires them to
agree.  Element-wise results agree bit for bit; dot products differ in
summation order (numpy pairwise vs. the 8-lane unrolled loop).

Citations: examples/dec_fluid.rs:46-141, 173-291; panopaea/src/dec/grid.rs:223-335;
panopaea/src/pcg.rs:14-82; panopaea/src/math/interp.rs:7-20.
"""
from __future__ import annotations

import numpy as np


def split(e, h, w):
    n0 = w * (h + 1)
    return e[:n0].reshape(h + 1, w), e[n0:].reshape(h, w + 1)


def lerp(a0, a1, s):                      # math/interp.rs:7-12
    return a0 * (1.0 - s) + a1 * s


def bilerp(a00, a01, a10, a11, s, t):     # math/interp.rs:15-20
    return lerp(lerp(a00, a01, s), lerp(a10, a11, s), t)


def hodge_1_dual(vy, vx):                 # dec/grid.rs:223-238
    return -vy, vx.copy()


def derivative_1_primal(ey, ex):          # dec/grid.rs:295-305
    return -ey[1:, :] + ey[:-1, :] - ex[:, :-1] + ex[:, 1:]


def derivative_0_dual(f, ey, ex):         # dec/grid.rs:318-334 (in place, interior only)
    ey[1:-1, :] = -(f[1:, :] - f[:-1, :])
    ex[:, 1:-1] = f[:, :-1] - f[:, 1:]


def advect(q, dt, vy, vx):                # dec_fluid.rs:173-211
    h, w = q.shape
    yy, xx = np.meshgrid(np.arange(h, dtype=q.dtype), np.arange(w, dtype=q.dtype), indexing="ij")
    ux = (vx[:, :-1] + vx[:, 1:]) / 2.0
    uy = (vy[:-1, :] + vy[1:, :]) / 2.0
    ppx = (xx + 0.5) + (-dt) * ux
    ppy = (yy + 0.5) + (-dt) * uy
    px = np.minimum(np.maximum(ppx - 0.5, 0.0), w - 1.00001)
    py = np.minimum(np.maximum(ppy - 0.5, 0.0), h - 1.00001)
    ix = np.floor(px).astype(np.int64)
    iy = np.floor(py).astype(np.int64)
    s = px - ix
    t = py - iy
    return bilerp(q[iy, ix], q[iy, ix + 1], q[iy + 1, ix], q[iy + 1, ix + 1], s, t)


def _gather_clamped(q, ppx_rel, ppy_rel):
    """index clamping rule of advect_mac (dec_fluid.rs:236-250)."""
    H, W = q.shape
    px = np.maximum(np.floor(ppx_rel), 0.0)
    py = np.maximum(np.floor(ppy_rel), 0.0)
    pxi = np.minimum(px, 1e15).astype(np.int64)
    pyi = np.minimum(py, 1e15).astype(np.int64)
    x0 = np.minimum(pxi, W - 1)
    x1 = np.minimum(pxi + 1, W - 1)
    y0 = np.minimum(pyi, H - 1)
    y1 = np.minimum(pyi + 1, H - 1)
    s = np.maximum(np.minimum(ppx_rel - pxi.astype(q.dtype), 1.0), 0.0)
    t = np.maximum(np.minimum(ppy_rel - pyi.astype(q.dtype), 1.0), 0.0)
    return bilerp(q[y0, x0], q[y0, x1], q[y1, x0], q[y1, x1], s, t)


def advect_mac(qy, qx, dt, vy, vx):       # dec_fluid.rs:213-291
    h, w = vx.shape[0], vy.shape[1]
    dtype = qx.dtype
    # x component, shape (h, w+1)
    yy, xx = np.meshgrid(np.arange(h, dtype=dtype), np.arange(w + 1, dtype=dtype), indexing="ij")
    xi = np.arange(w + 1)
    xc = np.minimum(xi, w - 1)
    xm = np.maximum(xi - 1, 0)
    vvy = (vy[:-1, xc] + vy[1:, xc] + vy[:-1, xm] + vy[1:, xm]) / 4.0
    ppx = (xx + 0.0) + (-dt) * vx
    ppy = (yy + 0.5) + (-dt) * vvy
    dx = _gather_clamped(qx, ppx - 0.0, ppy - 0.5)
    # y component, shape (h+1, w)
    yy, xx = np.meshgrid(np.arange(h + 1, dtype=dtype), np.arange(w, dtype=dtype), indexing="ij")
    yi = np.arange(h + 1)
    yc = np.minimum(yi, h - 1)
    ym = np.maximum(yi - 1, 0)
    vvx = (vx[yc, :-1] + vx[yc, 1:] + vx[ym, :-1] + vx[ym, 1:]) / 4.0
    ppx = (xx + 0.5) + (-dt) * vvx
    ppy = (yy + 0.0) + (-dt) * vy
    dy = _gather_clamped(qy, ppx - 0.5, ppy - 0.0)
    return dy, dx


def zero_rect(ey, ex, rect):
    y0, y1, x0, x1 = rect
    ey[y0:y1, x0:x1] = 0.0
    ex[y0:y1, x0:x1] = 0.0


def neg_divergence(vy, vx, obstacle):     # dec_fluid.rs:69-83
    ey, ex = hodge_1_dual(vy, vx)
    zero_rect(ey, ex, obstacle)
    return -derivative_1_primal(ey, ex)


def laplacian(p, dt, obstacle):           # dec_fluid.rs:100-119
    h, w = p.shape
    ey = np.zeros((h + 1, w), p.dtype)
    ex = np.zeros((h, w + 1), p.dtype)
    derivative_0_dual(p, ey, ex)
    zero_rect(ey, ex, obstacle)
    ey, ex = hodge_1_dual(ey, ex)
    return derivative_1_primal(ey, ex) * dt


def norm_max(a):
    return float(np.max(np.abs(a))) if a.size else 0.0


def pcg(b, max_iterations, threshold, apply_a):   # pcg.rs:14-82, identity preconditioner
    x = np.zeros_like(b)
    bmax = norm_max(b)
    if bmax < threshold:
        return x, -1, bmax
    r = b.copy()
    s = r.copy()
    sigma = float(np.dot(r.ravel(), r.ravel()))
    it, err = max_iterations, bmax
    for i in range(max_iterations):
        z = apply_a(s)
        alpha = sigma / float(np.dot(z.ravel(), s.ravel()))
        x = x + alpha * s
        r = r + (-alpha) * z
        err = norm_max(r)
        if err < threshold:
            it = i
            break
        sigma_new = float(np.dot(r.ravel(), r.ravel()))
        beta = sigma_new / sigma
        s = r + beta * s
        sigma = sigma_new
    return x, it, err


class FluidState:
    def __init__(self, h, w, timestep=0.05, threshold=0.1, max_iterations=100,
                 inflow=(5, 20, 54, 64), inflow_density=1.0, inflow_vy=20.0,
                 obstacle=(70, 80, 50, 70), dtype=np.float64):
        self.h, self.w, self.dt, self.threshold, self.max_iterations = h, w, timestep, threshold, max_iterations
        self.inflow, self.inflow_density, self.inflow_vy, self.obstacle = inflow, inflow_density, inflow_vy, obstacle
        self.vy = np.zeros((h + 1, w), dtype)
        self.vx = np.zeros((h, w + 1), dtype)
        self.density = np.zeros((h, w), dtype)
        self.pressure = np.zeros((h, w), dtype)

    def step(self):
        y0, y1, x0, x1 = self.inflow
        self.density[y0:y1, x0:x1] = self.inflow_density
        self.vy[y0:y1, x0:x1] = self.inflow_vy
        d = advect(self.density, self.dt, self.vy, self.vx)
        vy, vx = advect_mac(self.vy, self.vx, self.dt, self.vy, self.vx)
        self.density, self.vy, self.vx = d, vy, vx
        b = neg_divergence(self.vy, self.vx, self.obstacle)
        self.pressure, it, err = pcg(b, self.max_iterations, self.threshold,
                                     lambda s: laplacian(s, self.dt, self.obstacle))
        gy = np.zeros_like(self.vy)
        gx = np.zeros_like(self.vx)
        derivative_0_dual(self.pressure, gy, gx)
        self.vy = self.vy + self.dt * gy
        self.vx = self.vx + self.dt * gx
        self.vx[:, 0] = 0.0
        self.vx[:, -1] = 0.0
        self.vy[0, :] = 0.0
        self.vy[-1, :] = 0.0
        return dict(iterations=it, final_residual=err, rhs=b)


# ---------------------------------------------------------------------------------------------
# Multigrid / Jacobi preconditioners (no counterpart in the reference; trait seam pcg.rs:4-6).
# Array-at-a-time statement of the scheme that oracle/pano_oracle_mg.inc spells out loop by loop;
# tests/test_oracle_mg.py requires the two to agree to rounding (summation grouping differs).
def mg_level0_weights(h, w, obstacle):
    wy, wx = np.ones((h + 1, w)), np.ones((h, w + 1))
    wy[0, :] = 0
    wy[h, :] = 0
    wx[:, 0] = 0
    wx[:, w] = 0
    y0, y1, x0, x1 = obstacle
    wy[y0:y1, x0:x1] = 0
    wx[y0:y1, x0:x1] = 0
    return wy, wx


def mg_coarsen(wy, wx):
    h, w = wx.shape[0], wy.shape[1]
    hc, wc = (h + 1) // 2, (w + 1) // 2
    wyp = np.zeros((2 * hc + 1, 2 * wc))
    wyp[:h + 1, :w] = wy
    wxp = np.zeros((2 * hc, 2 * wc + 1))
    wxp[:h, :w + 1] = wx
    return 0.5 * (wyp[0::2, 0::2] + wyp[0::2, 1::2]), 0.5 * (wxp[0::2, 0::2] + wxp[1::2, 0::2])


def mg_apply_operator(wy, wx, u, dt):
    h, w = u.shape
    up = np.zeros((h + 2, w + 2))
    up[1:-1, 1:-1] = u
    c = up[1:-1, 1:-1]
    return dt * (wy[:-1] * (c - up[:-2, 1:-1]) + wy[1:] * (c - up[2:, 1:-1]) + wx[:, :-1] * (c - up[1:-1, :-2]) + wx[:, 1:] * (c - up[1:-1, 2:]))


class Multigrid:
    def __init__(self, h, w, timestep, obstacle=(0, 0, 0, 0), omega=0.8, nu=2, ncoarse=16, coarsest=8):
        self.dt, self.omega, self.nu, self.ncoarse = timestep, omega, nu, ncoarse
        self.W = [mg_level0_weights(h, w, obstacle)]
        while max(self.W[-1][1].shape[0], self.W[-1][0].shape[1]) > coarsest:
            self.W.append(mg_coarsen(*self.W[-1]))
        self.od = []
        for wy, wx in self.W:
            d = timestep * (wy[:-1] + wy[1:] + wx[:, :-1] + wx[:, 1:])
            od = np.zeros_like(d)
            od[d > 0] = omega / d[d > 0]
            self.od.append(od)

    def smooth(self, l, u, f, n, from_zero=False):
        wy, wx = self.W[l]
        for k in range(n):
            if from_zero and k == 0:
                u = self.od[l] * f
            else:
                u = u + self.od[l] * (f - mg_apply_operator(wy, wx, u, self.dt))
        return u

    def vcycle(self, l, f):
        wy, wx = self.W[l]
        if l == len(self.W) - 1:
            return self.smooth(l, None, f, self.ncoarse, True)
        u = self.smooth(l, None, f, self.nu, True)
        r = f - mg_apply_operator(wy, wx, u, self.dt)
        h, w = r.shape
        hc, wc = (h + 1) // 2, (w + 1) // 2
        rp = np.zeros((2 * hc, 2 * wc))
        rp[:h, :w] = r
        fc = rp[0::2, 0::2] + rp[0::2, 1::2] + rp[1::2, 0::2] + rp[1::2, 1::2]
        e = np.repeat(np.repeat(self.vcycle(l + 1, fc), 2, 0), 2, 1)[:h, :w]
        return self.smooth(l, u + e, f, self.nu)

    def apply(self, r):
        return self.vcycle(0, np.asarray(r, np.float64))

    def jacobi(self, r):
        wy, wx = self.W[0]
        d = self.dt * (wy[:-1] + wy[1:] + wx[:, :-1] + wx[:, 1:])
        out = np.zeros_like(d)
        out[d > 0] = np.asarray(r)[d > 0] / d[d > 0]
        return out
