"""panopaea_b200 -- B200-native drop-in for the grid fluid step of msiglreith/panopaea.

The product is the CUDA library (csrc/, C ABI in include/panopaea_b200.h).  This package is
the thin host-side mirror of the reference's Rust interface for that path -- same names,
argument order and panicking behaviour (as exceptions) -- used by the tests and bench.py:

    panopaea_b200.domain.Grid2d                 panopaea/src/domain/grid.rs:2-15
    panopaea_b200.dec.Simplex0/1/2, Manifold2d methods on Grid2d
                                                panopaea/src/dec/{grid,manifold}.rs
    panopaea_b200.pcg.precond_conjugate_gradient panopaea/src/pcg.rs:14-82
    panopaea_b200.fluid.advect / advect_mac / DecFluid
                                                examples/dec_fluid.rs
    panopaea_b200.grid3.Grid3d / trilinear / DecFluid3
                                                panopaea/src/domain/grid.rs:17-20, math/interp.rs:23-36 (+ DESIGN.md 5c)
"""
from . import _lib  # noqa: F401
from ._lib import PanoError  # noqa: F401
from .context import Context, default_context  # noqa: F401
from .domain import Grid2d  # noqa: F401
from .dec import Simplex0, Simplex1, Simplex2  # noqa: F401
from . import pcg, fluid, grid3  # noqa: F401
from .grid3 import Grid3d  # noqa: F401

__all__ = ["Context", "default_context", "Grid2d", "Simplex0", "Simplex1", "Simplex2", "pcg", "fluid", "grid3", "Grid3d", "PanoError"]
