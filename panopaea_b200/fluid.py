"""The example-local functions of examples/dec_fluid.rs (advect :173, advect_mac :213), the fused
passes built from them, and DecFluid: the example's main loop (:26-167) on device-resident fields."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import PcgInfo, Rect, StepParams, check
from .domain import Grid2d


def advect(dst, src, timestep, vel):
    """pub fn advect(dst: &mut Simplex2<f64>, src: &Simplex2<f64>, timestep: f64, vel: &Simplex1<f64>)"""
    check(_lib.load().pano_advect(dst.handle, src.handle, float(timestep), vel.handle))


def advect_mac(dst, src, timestep, vel):
    """pub fn advect_mac(dst: &mut Simplex1<f64>, src: &Simplex1<f64>, timestep: f64, vel: &Simplex1<f64>)"""
    check(_lib.load().pano_advect_mac(dst.handle, src.handle, float(timestep), vel.handle))


def advect_all(q_dst, vel_dst, q_src, vel, timestep):
    check(_lib.load().pano_advect_all(q_dst.handle, vel_dst.handle, q_src.handle, vel.handle, float(timestep)))


def neg_divergence(b, vel, obstacle=(0, 0, 0, 0), want_max=True):
    out = C.c_double()
    check(_lib.load().pano_neg_divergence(b.handle, vel.handle, Rect(*obstacle), C.byref(out) if want_max else None))
    return out.value if want_max else None


def laplacian_apply(z, s, timestep, obstacle=(0, 0, 0, 0)):
    check(_lib.load().pano_laplacian_apply(z.handle, s.handle, float(timestep), Rect(*obstacle)))


def project(vel, pressure, timestep):
    check(_lib.load().pano_project(vel.handle, pressure.handle, float(timestep)))


def density_to_u8(density, lower=-2.0, upper=2.0):
    h, w = density.grid.dim()
    out = np.empty((h, w), np.uint8)
    check(_lib.load().pano_density_to_u8(density.handle, float(lower), float(upper), out.ctypes.data_as(C.c_void_p)))
    return out


def export_png(path, density_u8):
    """util::png::export (panopaea_utils/src/png.rs:6-18) for the gray image returned by density_to_u8:
    RGB8 with the three channels equal, as examples/dec_fluid.rs:150-154 builds it.  The vertical flip
    already happened on the device.  Pure-Python encoder (zlib + struct): the PNG stays a host job."""
    import struct
    import zlib
    img = np.ascontiguousarray(density_u8, np.uint8)
    h, w = img.shape
    rgb = np.repeat(img[:, :, None], 3, axis=2)
    raw = b"".join(b"\x00" + rgb[y].tobytes() for y in range(h))

    def chunk(tag, data):
        body = tag + data
        return struct.pack(">I", len(data)) + body + struct.pack(">I", zlib.crc32(body) & 0xFFFFFFFF)

    png = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 2, 0, 0, 0)) + \
        chunk(b"IDAT", zlib.compress(raw, 6)) + chunk(b"IEND", b"")
    with open(path, "wb") as f:
        f.write(png)


def smoke_params(n: int):
    """Synthetic smoke plume of SURVEY.md 8(d): the shipped example scaled by k = n/128
    (n = 128 is examples/dec_fluid.rs:27-44, 51-54, 72-73 exactly)."""
    if n % 128:
        raise ValueError("n must be a multiple of 128")
    k = n // 128
    return dict(h=n, w=n, timestep=0.05, threshold=0.1, max_iterations=100,
                inflow=(5 * k, 20 * k, 54 * k, 64 * k), inflow_density=1.0, inflow_vy=20.0,
                obstacle=(70 * k, 80 * k, 50 * k, 70 * k))


class DecFluid:
    """examples/dec_fluid.rs main(): fields :29-41, literals :43-44, one `step()` per loop pass."""

    def __init__(self, h=128, w=128, timestep=0.05, threshold=0.1, max_iterations=100,
                 inflow=(5, 20, 54, 64), inflow_density=1.0, inflow_vy=20.0, obstacle=(70, 80, 50, 70), ctx=None):
        self.grid = Grid2d((h, w), ctx)
        g = self.grid
        self.vel, self.pressure, self.density = g.new_simplex_1(), g.new_simplex_2(), g.new_simplex_2()
        self.vel_temp, self.temp = g.new_simplex_1(), g.new_simplex_2()
        self.auxiliary, self.residual, self.search = g.new_simplex_2(), g.new_simplex_2(), g.new_simplex_2()
        self.params = StepParams(timestep, threshold, max_iterations, _lib.PRECOND_IDENTITY, Rect(*inflow),
                                 inflow_density, inflow_vy, Rect(*obstacle))
        self._L = _lib.load()

    def step(self, want_info=True):
        info = PcgInfo()
        check(self._L.pano_fluid_step(C.byref(self.params), self.density.handle, self.vel.handle, self.pressure.handle,
                                      self.temp.handle, self.vel_temp.handle, self.residual.handle,
                                      self.auxiliary.handle, self.search.handle, C.byref(info) if want_info else None))
        return info.as_dict() if want_info else None

    def step_composed(self):
        """The same loop body spelled with the reference's own call sequence (one device kernel
        per reference call, the generic CG driver, the Laplacian as a closure): the cross-check
        of the fused path and the proof that the drop-in API composes like the Rust crate."""
        from . import pcg
        g, p = self.grid, self.params
        inflow = (p.inflow.y0, p.inflow.y1, p.inflow.x0, p.inflow.x1)
        obstacle = (p.obstacle.y0, p.obstacle.y1, p.obstacle.x0, p.obstacle.x1)
        dt = p.timestep
        if not hasattr(self, "vel_primal_temp"):
            self.vel_primal_temp, self.pressure_temp = g.new_simplex_1(), g.new_simplex_2()
        self.density.fill_rect(inflow, p.inflow_density)                         # :48-57
        self.vel.fill_rect(inflow, p.inflow_vy, _lib.COMP_VY)
        advect(self.temp, self.density, dt, self.vel)                            # :59
        advect_mac(self.vel_temp, self.vel, dt, self.vel)                        # :60
        self.density.assign(self.temp)                                           # :62
        self.vel.assign(self.vel_temp)                                           # :63
        self.vel_temp.fill(0.0)                                                  # :65
        self.temp.fill(0.0)                                                      # :66
        g.hodge_1_dual(self.vel_temp, self.vel)                                  # :69
        self.vel_temp.fill_rect(obstacle, 0.0)                                   # :70-78
        g.derivative_1_primal(self.temp, self.vel_temp)                          # :80
        self.temp.scale(-1.0)                                                    # :81-83
        self.vel_temp.fill(0.0)                                                  # :89

        def closure(laplacian, pp):                                              # :100-119
            g.hodge_2_primal(self.pressure_temp, pp)
            g.derivative_0_dual(self.vel_temp, self.pressure_temp)
            self.vel_temp.fill_rect(obstacle, 0.0)
            g.hodge_1_dual(self.vel_primal_temp, self.vel_temp)
            g.derivative_1_primal(laplacian, self.vel_primal_temp)
            laplacian.scale(dt)

        info = pcg.precond_conjugate_gradient((), self.pressure, self.temp, p.max_iterations, p.threshold,
                                              self.residual, self.auxiliary, self.search, closure)   # :91-119
        g.hodge_2_primal(self.pressure_temp, self.pressure)                      # :124
        g.derivative_0_dual(self.vel_temp, self.pressure_temp)                   # :125
        self.vel.scaled_add(dt, self.vel_temp)                                   # :126
        h, w = g.dim()
        vx_rect_l, vx_rect_r = (0, h, 0, 1), (0, h, w, w + 1)                    # :132-135
        self.vel.fill_rect(vx_rect_l, 0.0, _lib.COMP_VX)
        self.vel.fill_rect(vx_rect_r, 0.0, _lib.COMP_VX)
        self.vel.fill_rect((0, 1, 0, w), 0.0, _lib.COMP_VY)                      # :137-140
        self.vel.fill_rect((h, h + 1, 0, w), 0.0, _lib.COMP_VY)
        return info
