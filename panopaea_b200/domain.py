"""domain::Grid2d (panopaea/src/domain/grid.rs:2-15) with the Manifold2d methods the
reference implements on it (panopaea/src/dec/manifold.rs:19-84, dec/grid.rs:343-371)."""
from __future__ import annotations

import numpy as np

from . import _lib
from ._lib import check


class Grid2d:
    def __init__(self, dim, ctx=None):
        self._dim = (int(dim[0]), int(dim[1]))   # (y, x)
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

    # ---- Manifold2d: counts and allocators (dec/grid.rs:346-371)
    def num_elem_0(self):
        return (self._dim[0] + 1) * (self._dim[1] + 1)

    def num_elem_1(self):
        h, w = self._dim
        return (h + 1) * w + h * (w + 1)

    def num_elem_2(self):
        return self._dim[0] * self._dim[1]

    def new_simplex_0(self, dtype=np.float64):
        from .dec import Simplex0
        return Simplex0(self, dtype)

    def new_simplex_1(self, dtype=np.float64):
        from .dec import Simplex1
        return Simplex1(self, dtype)

    def new_simplex_2(self, dtype=np.float64):
        from .dec import Simplex2
        return Simplex2(self, dtype)

    # ---- Manifold2d operators (dec/manifold.rs:46-83): (destination, source), as in the reference
    def _op(self, name, dst, src):
        check(getattr(_lib.load(), name)(dst.handle, src.handle))

    def derivative_0_primal(self, d_src, src):
        self._op("pano_derivative_0_primal", d_src, src)

    def derivative_0_dual(self, d_src, src):
        self._op("pano_derivative_0_dual", d_src, src)

    def derivative_1_primal(self, d_src, src):
        self._op("pano_derivative_1_primal", d_src, src)

    def derivative_1_dual(self, d_src, src):
        self._op("pano_derivative_1_dual", d_src, src)      # raises: unimplemented!() in the reference

    def hodge_0_primal(self, dual, primal):
        self._op("pano_hodge_0_primal", dual, primal)

    def hodge_2_dual(self, primal, dual):
        self._op("pano_hodge_2_dual", primal, dual)

    def hodge_1_primal(self, dual, primal):
        self._op("pano_hodge_1_primal", dual, primal)

    def hodge_1_dual(self, primal, dual):
        self._op("pano_hodge_1_dual", primal, dual)

    def hodge_2_primal(self, dual, primal):
        self._op("pano_hodge_2_primal", dual, primal)

    def hodge_0_dual(self, primal, dual):
        self._op("pano_hodge_0_dual", primal, dual)
