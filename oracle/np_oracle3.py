"""Second, independent statement of the Grid3d path in numpy (array at a time).

TEST INFRASTRUCTURE ONLY.  The reference has no 3-D fluid code (only the struct Grid3d,
panopaea/src/domain/grid.rs:17-20, and the unused `trilinear`, panopaea/src/math/interp.rs:23-36): DESIGN.md 5c
defines the path as examples/dec_fluid.rs:46-141 with a z axis, oracle/pano_oracle3.inc states it loop by loop in
C, and this file states it again from the maths so that tests/test_oracle3.py can demand bit equality of the
element-wise passes (dot products differ in summation order).  PARITY UNPINNED by the reference.
"""
from __future__ import annotations

import numpy as np

from .np_oracle import bilerp, lerp, norm_max, pcg


def trilerp(a000, a001, a010, a011, a100, a101, a110, a111, s, t, u):   # math/interp.rs:23-36
    return lerp(bilerp(a000, a001, a010, a011, s, t), bilerp(a100, a101, a110, a111, s, t), u)


def _mesh(d, h, w):
    return np.meshgrid(np.arange(d, dtype=np.float64), np.arange(h, dtype=np.float64), np.arange(w, dtype=np.float64), indexing="ij")


def advect(q, dt, vz, vy, vx):
    d, h, w = q.shape
    zz, yy, xx = _mesh(d, h, w)
    ux = (vx[:, :, :-1] + vx[:, :, 1:]) / 2.0
    uy = (vy[:, :-1, :] + vy[:, 1:, :]) / 2.0
    uz = (vz[:-1] + vz[1:]) / 2.0
    px = np.minimum(np.maximum(((xx + 0.5) + (-dt) * ux) - 0.5, 0.0), w - 1.00001)
    py = np.minimum(np.maximum(((yy + 0.5) + (-dt) * uy) - 0.5, 0.0), h - 1.00001)
    pz = np.minimum(np.maximum(((zz + 0.5) + (-dt) * uz) - 0.5, 0.0), d - 1.00001)
    ix, iy, iz = (np.floor(p).astype(np.int64) for p in (px, py, pz))
    return trilerp(q[iz, iy, ix], q[iz, iy, ix + 1], q[iz, iy + 1, ix], q[iz, iy + 1, ix + 1], q[iz + 1, iy, ix], q[iz + 1, iy, ix + 1],
                   q[iz + 1, iy + 1, ix], q[iz + 1, iy + 1, ix + 1], px - ix, py - iy, pz - iz)


def _axis(rel, n):
    p = np.maximum(np.floor(rel), 0.0)
    pi = np.minimum(p, 1e15).astype(np.int64)
    return np.minimum(pi, n - 1), np.minimum(pi + 1, n - 1), np.maximum(np.minimum(rel - pi.astype(np.float64), 1.0), 0.0)


def _gather_clamped(q, relx, rely, relz):
    D, H, W = q.shape
    x0, x1, s = _axis(relx, W)
    y0, y1, t = _axis(rely, H)
    z0, z1, u = _axis(relz, D)
    return trilerp(q[z0, y0, x0], q[z0, y0, x1], q[z0, y1, x0], q[z0, y1, x1], q[z1, y0, x0], q[z1, y0, x1], q[z1, y1, x0], q[z1, y1, x1],
                   s, t, u)


def advect_mac(qz, qy, qx, dt, vz, vy, vx):
    d, h, w = vx.shape[0], vx.shape[1], vy.shape[2]
    ndt = -dt
    # x component (d, h, w+1)
    zz, yy, xx = _mesh(d, h, w + 1)
    xi = np.arange(w + 1)
    xc, xm = np.minimum(xi, w - 1), np.maximum(xi - 1, 0)
    vvy = (vy[:, :-1][:, :, xc] + vy[:, 1:][:, :, xc] + vy[:, :-1][:, :, xm] + vy[:, 1:][:, :, xm]) / 4.0
    vvz = (vz[:-1][:, :, xc] + vz[1:][:, :, xc] + vz[:-1][:, :, xm] + vz[1:][:, :, xm]) / 4.0
    dx = _gather_clamped(qx, ((xx + 0.0) + ndt * vx) - 0.0, ((yy + 0.5) + ndt * vvy) - 0.5, ((zz + 0.5) + ndt * vvz) - 0.5)
    # y component (d, h+1, w)
    zz, yy, xx = _mesh(d, h + 1, w)
    yi = np.arange(h + 1)
    yc, ym = np.minimum(yi, h - 1), np.maximum(yi - 1, 0)
    vvx = (vx[:, yc][:, :, :-1] + vx[:, yc][:, :, 1:] + vx[:, ym][:, :, :-1] + vx[:, ym][:, :, 1:]) / 4.0
    vvz = (vz[:-1][:, yc] + vz[1:][:, yc] + vz[:-1][:, ym] + vz[1:][:, ym]) / 4.0
    dy = _gather_clamped(qy, ((xx + 0.5) + ndt * vvx) - 0.5, ((yy + 0.0) + ndt * vy) - 0.0, ((zz + 0.5) + ndt * vvz) - 0.5)
    # z component (d+1, h, w)
    zz, yy, xx = _mesh(d + 1, h, w)
    zi = np.arange(d + 1)
    zc, zm = np.minimum(zi, d - 1), np.maximum(zi - 1, 0)
    vvx = (vx[zc][:, :, :-1] + vx[zc][:, :, 1:] + vx[zm][:, :, :-1] + vx[zm][:, :, 1:]) / 4.0
    vvy = (vy[zc][:, :-1] + vy[zc][:, 1:] + vy[zm][:, :-1] + vy[zm][:, 1:]) / 4.0
    dz = _gather_clamped(qz, ((xx + 0.5) + ndt * vvx) - 0.5, ((yy + 0.5) + ndt * vvy) - 0.5, ((zz + 0.0) + ndt * vz) - 0.0)
    return dz, dy, dx


def _zero_box(ez, ey, ex, box):
    z0, z1, y0, y1, x0, x1 = box
    for e in (ez, ey, ex):
        e[z0:z1, y0:y1, x0:x1] = 0.0


def _faces_to_cells(ez, ey, ex):
    return -ez[1:] + ez[:-1] - ey[:, 1:] + ey[:, :-1] - ex[:, :, :-1] + ex[:, :, 1:]


def neg_divergence(vz, vy, vx, obstacle):
    ez, ey, ex = -vz, -vy, vx.copy()
    _zero_box(ez, ey, ex, obstacle)
    return -_faces_to_cells(ez, ey, ex)


def _gradient(p):
    d, h, w = p.shape
    ez, ey, ex = np.zeros((d + 1, h, w)), np.zeros((d, h + 1, w)), np.zeros((d, h, w + 1))
    ez[1:-1] = -(p[1:] - p[:-1])
    ey[:, 1:-1] = -(p[:, 1:] - p[:, :-1])
    ex[:, :, 1:-1] = p[:, :, :-1] - p[:, :, 1:]
    return ez, ey, ex


def laplacian(p, dt, obstacle):
    ez, ey, ex = _gradient(p)
    _zero_box(ez, ey, ex, obstacle)
    return _faces_to_cells(-ez, -ey, ex) * dt


def project(vz, vy, vx, p, dt):
    ez, ey, ex = _gradient(p)
    vz, vy, vx = vz + dt * ez, vy + dt * ey, vx + dt * ex
    vz[0] = vz[-1] = 0.0
    vy[:, 0] = vy[:, -1] = 0.0
    vx[:, :, 0] = vx[:, :, -1] = 0.0
    return vz, vy, vx


def step(state, params):
    """state: dict(density, vz, vy, vx, pressure); params: oracle.pano_oracle3.smoke_params-style dict.  One loop pass."""
    dt = params["timestep"]
    z0, z1, y0, y1, x0, x1 = params["inflow"]
    state["density"][z0:z1, y0:y1, x0:x1] = params["inflow_density"]
    state["vy"][z0:z1, y0:y1, x0:x1] = params["inflow_vy"]
    q = advect(state["density"], dt, state["vz"], state["vy"], state["vx"])
    vz, vy, vx = advect_mac(state["vz"], state["vy"], state["vx"], dt, state["vz"], state["vy"], state["vx"])
    b = neg_divergence(vz, vy, vx, params["obstacle"])
    x, it, err = pcg(b, params["max_iterations"], params["threshold"], lambda s: laplacian(s, dt, params["obstacle"]))
    vz, vy, vx = project(vz, vy, vx, x, dt)
    state.update(density=q, vz=vz, vy=vy, vx=vx, pressure=x)
    return dict(iterations=it, final_residual=err, rhs=b)
