"""ctypes front end of the Grid3d checker (oracle/pano_oracle3.inc, compiled into libpano_oracle.so).

TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED: the reference has no 3-D fluid path (only the struct Grid3d,
panopaea/src/domain/grid.rs:17-20, and the unused `trilinear`, panopaea/src/math/interp.rs:23-36); this is the
specification of DESIGN.md 5c, cross-checked against the independent numpy statement oracle/np_oracle3.py.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from .pano_oracle import PcgResult, _p, lib


class Box(C.Structure):
    _fields_ = [(k, C.c_size_t) for k in ("z0", "z1", "y0", "y1", "x0", "x1")]


class Params3(C.Structure):
    _fields_ = [("d", C.c_size_t), ("h", C.c_size_t), ("w", C.c_size_t), ("timestep", C.c_double), ("threshold", C.c_double),
                ("max_iterations", C.c_size_t), ("inflow", Box), ("inflow_density", C.c_double), ("inflow_vy", C.c_double),
                ("obstacle", Box)]


class LapCtx3(C.Structure):
    _fields_ = [("d", C.c_size_t), ("h", C.c_size_t), ("w", C.c_size_t), ("timestep", C.c_double), ("obstacle", Box),
                ("cell_temp", C.c_void_p), ("face_temp", C.c_void_p), ("face_primal_temp", C.c_void_p)]


def num_faces(d, h, w):
    return (d + 1) * h * w + d * (h + 1) * w + d * h * (w + 1)


def split(faces, d, h, w):
    """(vz (d+1,h,w), vy (d,h+1,w), vx (d,h,w+1)) views of a flat face buffer."""
    nz, ny = (d + 1) * h * w, d * (h + 1) * w
    return faces[:nz].reshape(d + 1, h, w), faces[nz:nz + ny].reshape(d, h + 1, w), faces[nz + ny:].reshape(d, h, w + 1)


def join(vz, vy, vx):
    return np.concatenate([np.ascontiguousarray(a, np.float64).ravel() for a in (vz, vy, vx)])


def _check_box(d, h, w, box):
    z0, z1, y0, y1, x0, x1 = box
    if z1 > z0 and y1 > y0 and x1 > x0 and (z1 > d or y1 > h or x1 > w or min(z0, y0, x0) < 0):
        raise IndexError(f"box {box} exceeds the {d}x{h}x{w} grid")


def trilinear(*a):
    return float(lib().orc_trilinear(*[C.c_double(v) for v in a]))


def _sz(*v):
    return [C.c_size_t(x) for x in v]


def advect(d, h, w, q, timestep, vel):
    q, vel = np.ascontiguousarray(q, np.float64), np.ascontiguousarray(vel, np.float64)
    dst = np.zeros(d * h * w)
    lib().orc3_advect(*_sz(d, h, w), _p(dst), _p(q), C.c_double(timestep), _p(vel))
    return dst.reshape(d, h, w)


def advect_mac(d, h, w, src, timestep, vel):
    src, vel = np.ascontiguousarray(src, np.float64), np.ascontiguousarray(vel, np.float64)
    dst = np.zeros_like(src)
    lib().orc3_advect_mac(*_sz(d, h, w), _p(dst), _p(src), C.c_double(timestep), _p(vel))
    return dst


def neg_divergence(d, h, w, vel, obstacle=(0,) * 6):
    _check_box(d, h, w, obstacle)
    vel = np.ascontiguousarray(vel, np.float64)
    b, tmp = np.zeros(d * h * w), np.zeros(num_faces(d, h, w))
    lib().orc3_neg_divergence(*_sz(d, h, w), _p(b), _p(vel), C.byref(Box(*obstacle)), _p(tmp))
    return b.reshape(d, h, w)


def _lap_ctx(d, h, w, timestep, obstacle):
    n, nf = d * h * w, num_faces(d, h, w)
    keep = (np.zeros(n), np.zeros(nf), np.zeros(nf))
    return LapCtx3(d, h, w, timestep, Box(*obstacle), *[_p(a) for a in keep]), keep


def laplacian_closure(d, h, w, p, timestep, obstacle=(0,) * 6):
    _check_box(d, h, w, obstacle)
    p = np.ascontiguousarray(p, np.float64)
    ctx, _keep = _lap_ctx(d, h, w, timestep, obstacle)
    out = np.zeros(d * h * w)
    lib().orc3_laplacian_closure(C.byref(ctx), _p(out), _p(p))
    return out.reshape(d, h, w)


def project(d, h, w, vel, p, timestep):
    vel = np.ascontiguousarray(vel, np.float64).copy()
    p = np.ascontiguousarray(p, np.float64)
    ct, ft = np.zeros(d * h * w), np.zeros(num_faces(d, h, w))
    lib().orc3_project(*_sz(d, h, w), _p(vel), _p(p), C.c_double(timestep), _p(ct), _p(ft))
    return vel


def pcg(d, h, w, b, max_iterations, threshold, timestep, obstacle=(0,) * 6):
    """pcg.rs:14-82 (the 2-D oracle's own loop) driven with the 7-point closure."""
    _check_box(d, h, w, obstacle)
    b = np.ascontiguousarray(b, np.float64)
    L = lib()
    n = d * h * w
    x, r, aux, s = (np.zeros(n) for _ in range(4))
    ctx, _keep = _lap_ctx(d, h, w, timestep, obstacle)
    info = (C.c_long * 2)()
    fres = C.c_double(0)
    L.orc_pcg_f64(C.c_size_t(n), _p(x), _p(b), C.c_size_t(max_iterations), C.c_double(threshold), _p(r), _p(aux), _p(s),
                  C.cast(L.orc3_laplacian_closure, C.c_void_p), C.byref(ctx), info, C.byref(fres))
    sh = (d, h, w)
    return PcgResult(x.reshape(sh), r.reshape(sh), s.reshape(sh), aux.reshape(sh), int(info[0]), int(info[1]), float(fres.value))


def smoke_params(n: int):
    """The smoke plume of SURVEY.md 8(d) with a z axis: k = n/128; the inflow and the obstacle keep their (y, x) rectangles and
    take the x extents along z too (the plume rises along y in the middle of the (z, x) cross-section)."""
    if n % 32:
        raise ValueError("n must be a multiple of 32")
    k = n / 128.0
    r = lambda v: int(round(v * k))
    return dict(d=n, h=n, w=n, timestep=0.05, threshold=0.1, max_iterations=100,
                inflow=(r(54), r(64), r(5), r(20), r(54), r(64)), inflow_density=1.0, inflow_vy=20.0,
                obstacle=(r(50), r(70), r(70), r(80), r(50), r(70)))


class FluidState3:
    FIELDS = dict(vel=0, pressure=1, density=2, vel_temp=3, temp=4, residual=5, auxiliary=6, search=7)

    def __init__(self, d, h, w, timestep=0.05, threshold=0.1, max_iterations=100, inflow=(0,) * 6, inflow_density=1.0,
                 inflow_vy=20.0, obstacle=(0,) * 6):
        _check_box(d, h, w, inflow)
        _check_box(d, h, w, obstacle)
        self.d, self.h, self.w = d, h, w
        self._params = Params3(d, h, w, timestep, threshold, max_iterations, Box(*inflow), inflow_density, inflow_vy, Box(*obstacle))
        self._L = lib()
        self._s = C.c_void_p(self._L.orc3_state_new(C.byref(self._params)))

    def field(self, name):
        ptr = self._L.orc3_state_field(self._s, C.c_int(self.FIELDS[name]))
        faces = name in ("vel", "vel_temp")
        n = num_faces(self.d, self.h, self.w) if faces else self.d * self.h * self.w
        a = np.frombuffer((C.c_double * n).from_address(ptr), dtype=np.float64)
        return a if faces else a.reshape(self.d, self.h, self.w)

    def step(self, want_rhs=False):
        info = (C.c_long * 2)()
        fres = C.c_double(0)
        rhs = np.zeros(self.d * self.h * self.w) if want_rhs else None
        self._L.orc3_step(self._s, info, C.byref(fres), _p(rhs) if want_rhs else None)
        out = dict(iterations=int(info[0]), applies=int(info[1]), final_residual=float(fres.value))
        if want_rhs:
            out["rhs"] = rhs.reshape(self.d, self.h, self.w)
        return out

    def close(self):
        if self._s:
            self._L.orc3_state_free(self._s)
            self._s = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
