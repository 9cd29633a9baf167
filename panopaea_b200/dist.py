"""Slab-decomposed step over several GPUs: thin wrapper of the pano_dist_* C ABI.

One process per GPU in production (`DistFluid.connect_ipc` takes the all-gathered IPC handles);
`connect_local` wires several ranks that live in one process (loop-back tests on a single GPU).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import PcgInfo, Rect, StepParams, check

DENSITY, VY, VX, PRESSURE = 0, 1, 2, 3


def slab_range(h, rank, nranks):
    """Rows [y0, y1) of an h-row grid owned by `rank` (balanced split, same formula as the library)."""
    return h * rank // nranks, h * (rank + 1) // nranks


def make_params(timestep=0.05, threshold=0.1, max_iterations=100, inflow=(5, 20, 54, 64), inflow_density=1.0, inflow_vy=20.0,
                obstacle=(70, 80, 50, 70), **_ignored):
    return StepParams(timestep, threshold, max_iterations, _lib.PRECOND_IDENTITY, Rect(*inflow), inflow_density, inflow_vy,
                      Rect(*obstacle))


class DistFluid:
    def __init__(self, ctx, h, w, rank, nranks, params):
        self._L = _lib.load()
        self.ctx, self.h, self.w, self.rank, self.nranks = ctx, h, w, rank, nranks
        self.y0, self.y1 = slab_range(h, rank, nranks)
        self.params = params if isinstance(params, StepParams) else make_params(**params)
        hnd = C.c_void_p()
        check(self._L.pano_dist_create(ctx.handle, h, w, rank, nranks, C.byref(self.params), C.byref(hnd)))
        self._h = hnd

    # ---- wiring
    def window(self):
        p, n = C.c_void_p(), C.c_size_t()
        check(self._L.pano_dist_window(self._h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def ipc_handle(self) -> bytes:
        buf = C.create_string_buffer(_lib.IPC_HANDLE_BYTES)
        check(self._L.pano_dist_ipc_handle(self._h, buf))
        return buf.raw

    def connect_ipc(self, handles):
        blob = b"".join(handles)
        assert len(blob) == _lib.IPC_HANDLE_BYTES * self.nranks
        check(self._L.pano_dist_connect(self._h, 1, C.c_char_p(blob)))

    def connect_local(self, window_ptrs):
        arr = (C.c_void_p * self.nranks)(*window_ptrs)
        check(self._L.pano_dist_connect(self._h, 0, arr))

    def set_max_ctas(self, n):
        check(self._L.pano_dist_set_max_ctas(self._h, int(n)))

    # ---- data
    def _rows(self, which):
        return (self.y1 - self.y0) + (1 if which == VY and self.rank == self.nranks - 1 else 0)

    def _pitch(self, which):
        return self.w + 1 if which == VX else self.w

    def upload(self, which, global_array):
        """global_array: the WHOLE field ((h,w), (h+1,w) or (h,w+1)); this rank takes its rows."""
        a = np.ascontiguousarray(global_array, np.float64)
        rows = np.ascontiguousarray(a[self.y0:self.y0 + self._rows(which)])
        check(self._L.pano_dist_upload(self._h, which, rows.ctypes.data_as(C.c_void_p)))

    def download(self, which):
        out = np.empty((self._rows(which), self._pitch(which)))
        n = C.c_size_t()
        check(self._L.pano_dist_download(self._h, which, out.ctypes.data_as(C.c_void_p), C.byref(n)))
        assert n.value == out.shape[0]
        return out

    # ---- stepping (collective)
    def step(self):
        check(self._L.pano_dist_step(self._h))

    def step_host(self, density_rows, vy_rows, vx_rows):
        """pano_dist_step_host: this rank's rows (C pointers to pinned host buffers, layout of upload / download) in, one step, rows
        out -- the density under the solve -- and the synchronisation; returns the solver info."""
        info = PcgInfo()
        check(self._L.pano_dist_step_host(self._h, density_rows, vy_rows, vx_rows, C.byref(info)))
        return info.as_dict()

    def solve(self):
        """The pressure solve alone, on the right-hand side of the last step (collective, asynchronous)."""
        check(self._L.pano_dist_solve(self._h))

    def sync(self):
        info = PcgInfo()
        check(self._L.pano_dist_sync(self._h, C.byref(info)))
        return info.as_dict()

    def close(self):
        if self._h:
            self._L.pano_dist_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def gather_local(ranks, which):
    """Assemble a global field from ranks living in this process."""
    return np.concatenate([r.download(which) for r in ranks], axis=0)
