"""dec::grid::Simplex0/1/2 (panopaea/src/dec/grid.rs:10, 37-62, 76) as handles to
device-resident fields, plus the LinearView / LinearViewReal surface
(panopaea/src/math/linear_view.rs:5-31) the CG loop and the example use."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import check, Rect


def _dtype_code(dtype):
    dtype = np.dtype(dtype)
    if dtype == np.float64:
        return _lib.F64
    if dtype == np.float32:
        return _lib.F32
    raise TypeError(f"unsupported dtype {dtype}: the path is f64 (f32 for the generic operators)")


class _Field:
    KIND = None

    def __init__(self, grid, dtype=np.float64):
        self.grid = grid
        self.dtype = np.dtype(dtype)
        h, w = grid.dim()
        self._L = _lib.load()
        hnd = C.c_void_p()
        check(self._L.pano_field_new(grid.ctx.handle, self.KIND, _dtype_code(dtype), h, w, C.byref(hnd)))
        self._h = hnd

    @property
    def handle(self):
        return self._h

    def __len__(self):
        n = C.c_size_t()
        check(self._L.pano_field_info(self._h, None, None, None, None, C.byref(n)))
        return n.value

    # ---- host <-> device over the flat view
    def upload(self, host):
        a = np.ascontiguousarray(host, dtype=self.dtype).ravel()
        check(self._L.pano_field_upload(self._h, a.ctypes.data_as(C.c_void_p), a.size))
        return self

    def view_linear(self):
        """The flat slice, copied to the host (math/linear_view.rs:7)."""
        out = np.empty(len(self), self.dtype)
        check(self._L.pano_field_download(self._h, out.ctypes.data_as(C.c_void_p), out.size))
        return out

    # ---- ndarray methods the path uses on view_linear_mut()
    def fill(self, value):
        check(self._L.pano_field_fill(self._h, float(value)))

    def assign(self, src):
        check(self._L.pano_field_assign(self._h, src.handle))

    def scaled_add(self, alpha, rhs):
        check(self._L.pano_field_scaled_add(self._h, float(alpha), rhs.handle))

    def scale(self, alpha):
        check(self._L.pano_field_scale(self._h, float(alpha)))

    def xpby(self, a, beta):
        check(self._L.pano_field_xpby(self._h, a.handle, float(beta)))

    def swap(self, other):
        check(self._L.pano_field_swap(self._h, other.handle))

    # ---- LinearViewReal (math/linear_view.rs:12-30)
    def dot_linear(self, rhs):
        out = C.c_double()
        check(self._L.pano_field_dot(self._h, rhs.handle, C.byref(out)))
        return out.value

    def norm_max(self):
        out = C.c_double()
        check(self._L.pano_field_norm_max(self._h, C.byref(out)))
        return out.value

    def fill_rect(self, rect, value, comp=_lib.COMP_ALL):
        check(self._L.pano_field_fill_rect(self._h, comp, Rect(*rect), float(value)))

    def free(self):
        if self._h:
            self._L.pano_field_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Simplex0(_Field):
    KIND = _lib.SIMPLEX0

    def to_host(self):
        h, w = self.grid.dim()
        return self.view_linear().reshape(h + 1, w + 1)


class Simplex2(_Field):
    KIND = _lib.SIMPLEX2

    def dim(self):
        return self.grid.dim()

    def to_host(self):
        return self.view_linear().reshape(self.grid.dim())


class Simplex1(_Field):
    KIND = _lib.SIMPLEX1

    def dim(self):
        return self.grid.dim()

    def split(self):
        """(vertical (h+1, w), horizontal (h, w+1)) copied to the host (dec/grid.rs:48-53)."""
        h, w = self.grid.dim()
        flat = self.view_linear()
        n0 = w * (h + 1)
        return flat[:n0].reshape(h + 1, w), flat[n0:].reshape(h, w + 1)

    def upload_split(self, vy, vx):
        return self.upload(np.concatenate([np.ascontiguousarray(vy, self.dtype).ravel(),
                                           np.ascontiguousarray(vx, self.dtype).ravel()]))
