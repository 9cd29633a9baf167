"""domain::Grid3d (panopaea/src/domain/grid.rs:17-20: the struct only) with the dec_fluid loop body carried to
(z, y, x) -- SURVEY.md 8(f) rank 4, defined in DESIGN.md 5c.  The reference's other 3-D item, math::trilinear
(panopaea/src/math/interp.rs:23-36), is `trilinear` below, argument order kept."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import Box, PcgInfo, Step3Params, check
from .dec import _Field

CELL3, FACE3 = 3, 4
COMP_VZ = 3


class Grid3d:
    def __init__(self, dim, ctx=None):
        self._dim = (int(dim[0]), int(dim[1]), int(dim[2]))   # (z, y, x)
        self._ctx = ctx

    @classmethod
    def new(cls, dim, ctx=None):
        return cls(dim, ctx)

    def dim(self):
        return self._dim

    @property
    def ctx(self):
        if self._ctx is None:
            from .context import default_context
            self._ctx = default_context()
        return self._ctx

    def num_cells(self):
        d, h, w = self._dim
        return d * h * w

    def num_faces(self):
        d, h, w = self._dim
        return (d + 1) * h * w + d * (h + 1) * w + d * h * (w + 1)

    def new_cells(self):
        return CellField3(self)

    def new_faces(self):
        return FaceField3(self)


class _Field3(_Field):
    def __init__(self, grid):
        self.grid = grid
        self.dtype = np.dtype(np.float64)
        d, h, w = grid.dim()
        self._L = _lib.load()
        hnd = C.c_void_p()
        check(self._L.pano_field3_new(grid.ctx.handle, self.KIND, d, h, w, C.byref(hnd)))
        self._h = hnd

    def fill_box(self, box, value, comp=_lib.COMP_ALL):
        check(self._L.pano_field3_fill_box(self._h, comp, Box(*box), float(value)))

    def fill_rect(self, *a, **k):
        raise TypeError("a Grid3d field: use fill_box")


class CellField3(_Field3):
    KIND = CELL3

    def to_host(self):
        return self.view_linear().reshape(self.grid.dim())


class FaceField3(_Field3):
    KIND = FACE3

    def split(self):
        """(vz (d+1,h,w), vy (d,h+1,w), vx (d,h,w+1)) copied to the host."""
        d, h, w = self.grid.dim()
        flat = self.view_linear()
        nz, ny = (d + 1) * h * w, d * (h + 1) * w
        return flat[:nz].reshape(d + 1, h, w), flat[nz:nz + ny].reshape(d, h + 1, w), flat[nz + ny:].reshape(d, h, w + 1)

    def upload_split(self, vz, vy, vx):
        return self.upload(np.concatenate([np.ascontiguousarray(a, np.float64).ravel() for a in (vz, vy, vx)]))


def trilinear(a000, a001, a010, a011, a100, a101, a110, a111, s, t, u):
    return float(_lib.load().pano_trilinear(*[float(v) for v in (a000, a001, a010, a011, a100, a101, a110, a111, s, t, u)]))


def advect(dst, src, timestep, vel):
    check(_lib.load().pano_advect3(dst.handle, src.handle, float(timestep), vel.handle))


def advect_mac(dst, src, timestep, vel):
    check(_lib.load().pano_advect3_mac(dst.handle, src.handle, float(timestep), vel.handle))


def advect_all(q_dst, vel_dst, q_src, vel, timestep):
    check(_lib.load().pano_advect3_all(q_dst.handle, vel_dst.handle, q_src.handle, vel.handle, float(timestep)))


def neg_divergence(b, vel, obstacle=(0,) * 6, want_max=True):
    out = C.c_double()
    check(_lib.load().pano_neg_divergence3(b.handle, vel.handle, Box(*obstacle), C.byref(out) if want_max else None))
    return out.value if want_max else None


def laplacian_apply(z, s, timestep, obstacle=(0,) * 6):
    check(_lib.load().pano_laplacian3_apply(z.handle, s.handle, float(timestep), Box(*obstacle)))


def project(vel, pressure, timestep):
    check(_lib.load().pano_project3(vel.handle, pressure.handle, float(timestep)))


def pcg_solve(x, b, max_iterations, threshold, residual, auxiliary, search, timestep, obstacle=(0,) * 6, want_info=True,
              precond=_lib.PRECOND_IDENTITY):
    """pcg::precond_conjugate_gradient(&(), x, b, max_iterations, threshold, residual, auxiliary, search, A) with the 7-point closure."""
    info = PcgInfo()
    check(_lib.load().pano_pcg3_solve(precond, x.handle, b.handle, int(max_iterations), float(threshold), residual.handle,
                                      auxiliary.handle, search.handle, float(timestep), Box(*obstacle),
                                      C.byref(info) if want_info else None))
    return info.as_dict() if want_info else None


def smoke_params(n: int):
    """The synthetic plume of SURVEY.md 8(d) with a z axis (k = n/128): the (y, x) rectangles of examples/dec_fluid.rs:51-52, 72-73,
    their x extents repeated along z."""
    if n % 32:
        raise ValueError("n must be a multiple of 32")
    k = n / 128.0
    r = lambda v: int(round(v * k))
    return dict(d=n, h=n, w=n, timestep=0.05, threshold=0.1, max_iterations=100,
                inflow=(r(54), r(64), r(5), r(20), r(54), r(64)), inflow_density=1.0, inflow_vy=20.0,
                obstacle=(r(50), r(70), r(70), r(80), r(50), r(70)))


class DecFluid3:
    """examples/dec_fluid.rs main() on a Grid3d: the same fields (:29-41) and one `step()` per loop pass."""

    def __init__(self, d, h, w, timestep=0.05, threshold=0.1, max_iterations=100, inflow=(0,) * 6, inflow_density=1.0,
                 inflow_vy=20.0, obstacle=(0,) * 6, ctx=None):
        self.grid = g = Grid3d((d, h, w), ctx)
        self.vel, self.pressure, self.density = g.new_faces(), g.new_cells(), g.new_cells()
        self.vel_temp, self.temp = g.new_faces(), g.new_cells()
        self.auxiliary, self.residual, self.search = g.new_cells(), g.new_cells(), g.new_cells()
        self.params = Step3Params(timestep, threshold, max_iterations, _lib.PRECOND_IDENTITY, Box(*inflow), inflow_density, inflow_vy,
                                  Box(*obstacle))
        self._L = _lib.load()

    def step(self, want_info=True):
        info = PcgInfo()
        check(self._L.pano_fluid3_step(C.byref(self.params), self.density.handle, self.vel.handle, self.pressure.handle,
                                       self.temp.handle, self.vel_temp.handle, self.residual.handle, self.auxiliary.handle,
                                       self.search.handle, C.byref(info) if want_info else None))
        return info.as_dict() if want_info else None


def fluid3_step_host(ctx, params, d, h, w, density, vel, pressure=None, want_info=True):
    """pano_fluid3_step_host: the host owns the fields (contiguous float64 arrays: density (d,h,w), vel flat faces, optional pressure);
    they are updated in place.  Use pinned buffers (pano_host_alloc) for speed; any contiguous array works."""
    info = PcgInfo()
    check(_lib.load().pano_fluid3_step_host(ctx.handle, C.byref(params), d, h, w, density.ctypes.data_as(C.c_void_p),
                                            vel.ctypes.data_as(C.c_void_p),
                                            pressure.ctypes.data_as(C.c_void_p) if pressure is not None else None,
                                            C.byref(info) if want_info else None))
    return info.as_dict() if want_info else None
