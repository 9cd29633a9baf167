"""ctypes front end of the CPU oracle (oracle/pano_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  Never imported by the
product package ``panopaea_b200``.

The oracle is a C restatement ("port") of the Rust reference, which cannot be
built in this image.  See the header of pano_oracle.c for what pins it.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libpano_oracle.so")

SERIAL, REFERENCE_FAITHFUL, ALL_PARALLEL = 0, 1, 2


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in ("pano_oracle.c", "pano_oracle_body.inc", "pano_oracle_mg.inc", "pano_oracle3.inc", "Makefile")]
    stale = (not os.path.exists(_SO)) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs)
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-s"] + (["-B"] if force else []), check=True)
    return _SO


class Rect(C.Structure):
    _fields_ = [("y0", C.c_size_t), ("y1", C.c_size_t), ("x0", C.c_size_t), ("x1", C.c_size_t)]


def _params_type(real):
    class Params(C.Structure):
        _fields_ = [("h", C.c_size_t), ("w", C.c_size_t), ("timestep", real), ("threshold", real),
                    ("max_iterations", C.c_size_t), ("inflow", Rect), ("inflow_density", real),
                    ("inflow_vy", real), ("obstacle", Rect)]
    return Params


def _lap_ctx_type(real):
    class LapCtx(C.Structure):
        _fields_ = [("h", C.c_size_t), ("w", C.c_size_t), ("timestep", real), ("obstacle", Rect),
                    ("pressure_temp", C.c_void_p), ("vel_temp", C.c_void_p), ("vel_primal_temp", C.c_void_p)]
    return LapCtx


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.orc_max_threads.restype = C.c_int
        _lib.orc_get_threading.restype = C.c_int
        for sfx, real in (("f64", C.c_double), ("f32", C.c_float)):
            getattr(_lib, f"orc_dot_linear_{sfx}").restype = real
            getattr(_lib, f"orc_norm_max_{sfx}").restype = real
            getattr(_lib, f"orc_state_new_{sfx}").restype = C.c_void_p
            getattr(_lib, f"orc_state_field_{sfx}").restype = C.c_void_p
            getattr(_lib, f"orc_num_elem_1_{sfx}").restype = C.c_size_t
        _lib.orc_mg_new.restype = C.c_void_p
        _lib.orc_trilinear.restype = C.c_double
        _lib.orc_trilinear.argtypes = [C.c_double] * 11
        _lib.orc3_num_faces.restype = C.c_size_t
        _lib.orc3_state_new.restype = C.c_void_p
        _lib.orc3_state_field.restype = C.c_void_p
        _lib.orc_mg_level_wy.restype = C.c_void_p
        _lib.orc_mg_level_wx.restype = C.c_void_p
    return _lib


def set_threading(mode: int, num_threads: int | None = None) -> int:
    L = lib()
    L.orc_set_threading(C.c_int(mode))
    if num_threads is not None:
        L.orc_set_num_threads(C.c_int(num_threads))
    return L.orc_max_threads()


def max_threads() -> int:
    return lib().orc_max_threads()


def _sfx(dtype):
    dtype = np.dtype(dtype)
    if dtype == np.float64:
        return "f64", C.c_double
    if dtype == np.float32:
        return "f32", C.c_float
    raise TypeError(dtype)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _arr(a, dtype=None):
    a = np.ascontiguousarray(a, dtype=dtype)
    return a


def num_elem_1(h, w):
    return (h + 1) * w + h * (w + 1)


def split(edges, h, w):
    """(vy (h+1,w), vx (h,w+1)) views of a flat Simplex1 buffer (dec/grid.rs:48-52)."""
    n0 = w * (h + 1)
    return edges[:n0].reshape(h + 1, w), edges[n0:].reshape(h, w + 1)


def join(vy, vx):
    return np.concatenate([np.ascontiguousarray(vy).ravel(), np.ascontiguousarray(vx).ravel()])


def _fn(name, dtype):
    s, real = _sfx(dtype)
    return getattr(lib(), f"{name}_{s}"), real


def _unary_s1(name, h, w, src):
    src = _arr(src)
    f, _ = _fn(name, src.dtype)
    dst = np.zeros_like(src)
    f(C.c_size_t(h), C.c_size_t(w), _p(dst), _p(src))
    return dst


def hodge_1_primal(h, w, primal):
    return _unary_s1("orc_hodge_1_primal", h, w, primal)


def hodge_1_dual(h, w, dual):
    return _unary_s1("orc_hodge_1_dual", h, w, dual)


def hodge_2_primal(h, w, primal):
    return _unary_s1("orc_hodge_2_primal", h, w, primal)


def hodge_0_primal(h, w, primal, out=None):
    primal = _arr(primal)
    f, _ = _fn("orc_hodge_0_primal", primal.dtype)
    dst = np.zeros_like(primal) if out is None else out
    f(C.c_size_t(h), C.c_size_t(w), _p(dst), _p(primal))
    return dst


def hodge_2_dual(h, w, dual, out=None):
    dual = _arr(dual)
    f, _ = _fn("orc_hodge_2_dual", dual.dtype)
    dst = np.zeros_like(dual) if out is None else out
    f(C.c_size_t(h), C.c_size_t(w), _p(dst), _p(dual))
    return dst


def derivative_0_dual(h, w, faces, out=None):
    """out (flat Simplex1) keeps its boundary edges: the reference leaves them untouched."""
    faces = _arr(faces)
    f, _ = _fn("orc_derivative_0_dual", faces.dtype)
    edges = np.zeros(num_elem_1(h, w), dtype=faces.dtype) if out is None else out
    f(C.c_size_t(h), C.c_size_t(w), _p(edges), _p(faces))
    return edges


def derivative_1_primal(h, w, edges):
    edges = _arr(edges)
    f, _ = _fn("orc_derivative_1_primal", edges.dtype)
    faces = np.zeros(h * w, dtype=edges.dtype)
    f(C.c_size_t(h), C.c_size_t(w), _p(faces), _p(edges))
    return faces.reshape(h, w)


def derivative_0_primal(h, w, vertices):
    vertices = _arr(vertices)
    f, _ = _fn("orc_derivative_0_primal", vertices.dtype)
    edges = np.zeros(num_elem_1(h, w), dtype=vertices.dtype)
    f(C.c_size_t(h), C.c_size_t(w), _p(edges), _p(vertices))
    return edges


def dot_linear(a, b):
    a, b = _arr(a), _arr(b)
    f, _ = _fn("orc_dot_linear", a.dtype)
    return float(f(C.c_size_t(a.size), _p(a), _p(b)))


def norm_max(a):
    a = _arr(a)
    f, _ = _fn("orc_norm_max", a.dtype)
    return float(f(C.c_size_t(a.size), _p(a)))


def scaled_add(y, alpha, x):
    y, x = _arr(y).copy(), _arr(x)
    f, real = _fn("orc_scaled_add", y.dtype)
    f(C.c_size_t(y.size), _p(y), real(alpha), _p(x))
    return y


def advect(h, w, q, timestep, vel):
    q, vel = _arr(q), _arr(vel)
    f, real = _fn("orc_advect", q.dtype)
    dst = np.zeros(h * w, dtype=q.dtype)
    f(C.c_size_t(h), C.c_size_t(w), _p(dst), _p(q), real(timestep), _p(vel))
    return dst.reshape(h, w)


def advect_mac(h, w, src, timestep, vel):
    src, vel = _arr(src), _arr(vel)
    f, real = _fn("orc_advect_mac", src.dtype)
    dst = np.zeros_like(src)
    f(C.c_size_t(h), C.c_size_t(w), _p(dst), _p(src), real(timestep), _p(vel))
    return dst


def _check_rect(h, w, rect):
    y0, y1, x0, x1 = rect
    if y1 > y0 and x1 > x0 and (y1 > h or x1 > w or y0 < 0 or x0 < 0):
        raise IndexError(f"rectangle {rect} exceeds the {h}x{w} grid: the reference's vy[(y,x)] / vx[(y,x)] would panic")


def laplacian_closure(h, w, p, timestep, obstacle=(0, 0, 0, 0)):
    """A(p) of examples/dec_fluid.rs:100-119 (vel_temp boundary edges zero as after :89)."""
    _check_rect(h, w, obstacle)
    p = _arr(p)
    sfx, real = _sfx(p.dtype)
    n1 = num_elem_1(h, w)
    pt, vt, vpt = np.zeros(h * w, p.dtype), np.zeros(n1, p.dtype), np.zeros(n1, p.dtype)
    ctx = _lap_ctx_type(real)(h, w, timestep, Rect(*obstacle), _p(pt), _p(vt), _p(vpt))
    out = np.zeros(h * w, p.dtype)
    getattr(lib(), f"orc_laplacian_closure_{sfx}")(C.byref(ctx), _p(out), _p(p))
    return out.reshape(h, w)


@dataclass
class PcgResult:
    x: np.ndarray
    residual: np.ndarray
    search: np.ndarray
    auxiliary: np.ndarray
    iterations: int      # index i at the break (pcg.rs:61); max_iterations if exhausted; -1 early-out
    applies: int
    final_residual: float


def pcg_grid_laplacian(h, w, b, max_iterations, threshold, timestep, obstacle=(0, 0, 0, 0)):
    """pcg.rs:14-82 driven with the dec_fluid Laplacian closure."""
    _check_rect(h, w, obstacle)
    b = _arr(b)
    sfx, real = _sfx(b.dtype)
    L = lib()
    n, n1 = h * w, num_elem_1(h, w)
    x, r, aux, s = (np.zeros(n, b.dtype) for _ in range(4))
    pt, vt, vpt = np.zeros(n, b.dtype), np.zeros(n1, b.dtype), np.zeros(n1, b.dtype)
    ctx = _lap_ctx_type(real)(h, w, timestep, Rect(*obstacle), _p(pt), _p(vt), _p(vpt))
    info = (C.c_long * 2)()
    fres = real(0)
    closure = C.cast(getattr(L, f"orc_laplacian_closure_{sfx}"), C.c_void_p)
    getattr(L, f"orc_pcg_{sfx}")(C.c_size_t(n), _p(x), _p(b), C.c_size_t(max_iterations), real(threshold),
                                 _p(r), _p(aux), _p(s), closure, C.byref(ctx), info, C.byref(fres))
    return PcgResult(x.reshape(h, w), r.reshape(h, w), s.reshape(h, w), aux.reshape(h, w),
                     int(info[0]), int(info[1]), float(fres.value))


class Multigrid:
    """The multigrid preconditioner specified in oracle/pano_oracle_mg.inc (f64).  apply(src) -> dst;
    jacobi(src) is the diagonal preconditioner built from the same face weights."""

    def __init__(self, h, w, timestep, obstacle=(0, 0, 0, 0)):
        _check_rect(h, w, obstacle)
        self.h, self.w = h, w
        self._L = lib()
        self._m = C.c_void_p(self._L.orc_mg_new(C.c_size_t(h), C.c_size_t(w), C.c_double(timestep), C.byref(Rect(*obstacle))))

    @property
    def levels(self):
        return int(self._L.orc_mg_levels(self._m))

    def level_dim(self, l):
        hh, ww = C.c_size_t(), C.c_size_t()
        self._L.orc_mg_level_dim(self._m, l, C.byref(hh), C.byref(ww))
        return hh.value, ww.value

    def level_weights(self, l):
        hh, ww = self.level_dim(l)
        wy = np.ctypeslib.as_array(C.cast(self._L.orc_mg_level_wy(self._m, l), C.POINTER(C.c_double)), shape=((hh + 1) * ww,)).copy()
        wx = np.ctypeslib.as_array(C.cast(self._L.orc_mg_level_wx(self._m, l), C.POINTER(C.c_double)), shape=(hh * (ww + 1),)).copy()
        return wy.reshape(hh + 1, ww), wx.reshape(hh, ww + 1)

    def apply(self, src):
        src = np.ascontiguousarray(src, np.float64)
        dst = np.zeros(self.h * self.w)
        self._L.orc_mg_apply(self._m, _p(dst), _p(src))
        return dst.reshape(self.h, self.w)

    def jacobi(self, src):
        src = np.ascontiguousarray(src, np.float64)
        dst = np.zeros(self.h * self.w)
        self._L.orc_jacobi_apply(self._m, _p(dst), _p(src))
        return dst.reshape(self.h, self.w)

    def close(self):
        if self._m:
            self._L.orc_mg_free(self._m)
            self._m = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def pcg_grid_laplacian_precond(h, w, b, max_iterations, threshold, timestep, obstacle=(0, 0, 0, 0), precond="multigrid"):
    """pcg.rs:14-82 with the dec_fluid Laplacian closure and a non-identity Preconditioner (f64):
    "multigrid" | "jacobi" | "identity"."""
    _check_rect(h, w, obstacle)
    b = np.ascontiguousarray(b, np.float64)
    L = lib()
    n, n1 = h * w, num_elem_1(h, w)
    x, r, aux, s = (np.zeros(n) for _ in range(4))
    pt, vt, vpt = np.zeros(n), np.zeros(n1), np.zeros(n1)
    ctx = _lap_ctx_type(C.c_double)(h, w, timestep, Rect(*obstacle), _p(pt), _p(vt), _p(vpt))
    info = (C.c_long * 2)()
    fres = C.c_double(0)
    closure = C.cast(L.orc_laplacian_closure_f64, C.c_void_p)
    mg = Multigrid(h, w, timestep, obstacle) if precond != "identity" else None
    apply = {"multigrid": L.orc_mg_apply, "jacobi": L.orc_jacobi_apply}.get(precond)
    L.orc_pcg_precond(C.c_size_t(n), _p(x), _p(b), C.c_size_t(max_iterations), C.c_double(threshold), _p(r), _p(aux), _p(s),
                      closure, C.byref(ctx), C.cast(apply, C.c_void_p) if apply else None, mg._m if mg else None, info, C.byref(fres))
    return PcgResult(x.reshape(h, w), r.reshape(h, w), s.reshape(h, w), aux.reshape(h, w),
                     int(info[0]), int(info[1]), float(fres.value))


def smoke_params(n: int, dtype=np.float64):
    """Synthetic smoke plume of SURVEY.md 8(d): the shipped example scaled by k = n/128.
    At n = 128 this is examples/dec_fluid.rs:27-44, 51-54, 72-73 bit for bit."""
    if n % 128:
        raise ValueError("n must be a multiple of 128")
    k = n // 128
    return dict(h=n, w=n, timestep=0.05, threshold=0.1, max_iterations=100,
                inflow=(5 * k, 20 * k, 54 * k, 64 * k), inflow_density=1.0, inflow_vy=20.0,
                obstacle=(70 * k, 80 * k, 50 * k, 70 * k))


class FluidState:
    """State of examples/dec_fluid.rs main (fields :29-41) advanced by orc_step."""
    FIELDS = dict(vel=0, pressure=1, density=2, vel_temp=3, temp=4, residual=5, auxiliary=6, search=7)

    def __init__(self, h, w, timestep=0.05, threshold=0.1, max_iterations=100,
                 inflow=(5, 20, 54, 64), inflow_density=1.0, inflow_vy=20.0,
                 obstacle=(70, 80, 50, 70), dtype=np.float64):
        self.sfx, self.real = _sfx(dtype)
        self.dtype = np.dtype(dtype)
        self.h, self.w = h, w
        _check_rect(h, w, inflow)
        _check_rect(h, w, obstacle)
        P = _params_type(self.real)
        self._params = P(h, w, timestep, threshold, max_iterations, Rect(*inflow), inflow_density,
                         inflow_vy, Rect(*obstacle))
        self._L = lib()
        self._s = C.c_void_p(getattr(self._L, f"orc_state_new_{self.sfx}")(C.byref(self._params)))

    def field(self, name):
        which = self.FIELDS[name]
        ptr = getattr(self._L, f"orc_state_field_{self.sfx}")(self._s, C.c_int(which))
        n = num_elem_1(self.h, self.w) if name in ("vel", "vel_temp") else self.h * self.w
        buf = (self.real * n).from_address(ptr)
        a = np.frombuffer(buf, dtype=self.dtype)
        return a if name in ("vel", "vel_temp") else a.reshape(self.h, self.w)

    def step(self, want_rhs=False):
        info = (C.c_long * 2)()
        fres = self.real(0)
        rhs = np.zeros(self.h * self.w, self.dtype) if want_rhs else None
        getattr(self._L, f"orc_step_{self.sfx}")(self._s, info, C.byref(fres), _p(rhs) if want_rhs else None)
        out = dict(iterations=int(info[0]), applies=int(info[1]), final_residual=float(fres.value))
        if want_rhs:
            out["rhs"] = rhs.reshape(self.h, self.w)
        return out

    def close(self):
        if self._s:
            getattr(self._L, f"orc_state_free_{self.sfx}")(self._s)
            self._s = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
