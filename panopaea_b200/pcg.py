"""pcg (panopaea/src/pcg.rs): the Preconditioner trait, its identity impl for `()`, and
precond_conjugate_gradient with the reference's argument order.

Two drivers:
  * precond_conjugate_gradient(...)  -- generic: any operator closure `a(dst, src)`, any
    preconditioner object; the loop of pcg.rs:32-80 runs on the host over device primitives
    (each dot / max-norm is a device reduction read back, like the reference's scalars).
  * solve_grid_laplacian(...)        -- the hot path: operator fixed to the dec_fluid closure,
    the whole solve is one persistent CUDA kernel with all scalars on the device.
"""
from __future__ import annotations

import ctypes as C

from . import _lib
from ._lib import PcgInfo, Rect, check


class Identity:
    """impl Preconditioner<L> for ()  (pcg.rs:8-12)."""

    def apply(self, dst, src):
        dst.assign(src)


class Jacobi:
    """Preconditioner for the dec_fluid operator A = timestep * Laplacian(open faces): dst = src / diag(A).
    Not in the reference (its trait has the `()` impl only); specified in DESIGN.md 5b."""
    kind = _lib.PRECOND_JACOBI

    def __init__(self, grid, timestep, obstacle=(0, 0, 0, 0)):
        self.grid, self.timestep, self.obstacle = grid, float(timestep), tuple(obstacle)

    def apply(self, dst, src):
        check(_lib.load().pano_jacobi_apply(dst.handle, src.handle, self.timestep, Rect(*self.obstacle)))


class Multigrid:
    """Preconditioner: one geometric-multigrid V-cycle for the same operator (csrc/pano_mg.cu).  Symmetric positive
    definite, so pcg.rs:14-82 applies unchanged; brings the solve of the smoke plume from >100 iterations to 2-3."""
    kind = _lib.PRECOND_MULTIGRID

    def __init__(self, grid, timestep, obstacle=(0, 0, 0, 0)):
        self.grid, self.timestep, self.obstacle = grid, float(timestep), tuple(obstacle)
        h, w = grid.dim()
        self._L = _lib.load()
        self._h = C.c_void_p()
        check(self._L.pano_mg_create(grid.ctx.handle, h, w, self.timestep, Rect(*self.obstacle), C.byref(self._h)))

    def levels(self):
        n, t = C.c_int(), C.c_int()
        check(self._L.pano_mg_levels(self._h, C.byref(n), C.byref(t)))
        return n.value, t.value

    def apply(self, dst, src):
        check(self._L.pano_mg_apply(self._h, dst.handle, src.handle))

    def close(self):
        if self._h:
            self._L.pano_mg_destroy(self._h)
            self._h = None


def precond_conjugate_gradient(preconditioner, x, b, max_iterations, threshold, residual, auxiliary, search, a):
    """pcg.rs:14-82.  `preconditioner` None or () means the identity.  Returns what the reference prints:
    dict(iterations=index at the break | max_iterations | -1 on early out, final_residual=...)."""
    if preconditioner is None or preconditioner == ():
        preconditioner = Identity()
    x.fill(0.0)                                               # :32
    bmax = b.norm_max()
    if bmax < threshold:                                      # :35-38
        return dict(iterations=-1, applies=0, final_residual=bmax, rhs_max=bmax)
    residual.assign(b)                                        # :40
    preconditioner.apply(auxiliary, residual)                 # :41
    search.assign(auxiliary)                                  # :42
    sigma = auxiliary.dot_linear(residual)                    # :46
    it, applies, err = max_iterations, 0, bmax
    for i in range(max_iterations):                           # :48
        a(auxiliary, search)                                  # :51
        applies += 1
        alpha = sigma / auxiliary.dot_linear(search)          # :53
        x.scaled_add(alpha, search)                           # :55
        residual.scaled_add(-alpha, auxiliary)                # :56
        err = residual.norm_max()                             # :58
        if err < threshold:                                   # :60-63
            it = i
            break
        preconditioner.apply(auxiliary, residual)             # :65
        sigma_new = auxiliary.dot_linear(residual)            # :67
        beta = sigma_new / sigma                              # :68
        search.xpby(auxiliary, beta)                          # :72-77  search = aux + beta*search
        sigma = sigma_new                                     # :79
    return dict(iterations=it, applies=applies, final_residual=err, rhs_max=bmax)


def solve_grid_laplacian(x, b, max_iterations, threshold, residual, auxiliary, search, timestep,
                         obstacle=(0, 0, 0, 0), preconditioner=None, want_info=True):
    """precond_conjugate_gradient with `a` = the closure of examples/dec_fluid.rs:100-119, fused."""
    if preconditioner in (None, ()) or isinstance(preconditioner, Identity):
        kind = _lib.PRECOND_IDENTITY
    elif isinstance(preconditioner, (Jacobi, Multigrid)):
        kind = preconditioner.kind
        if float(timestep) != preconditioner.timestep or tuple(obstacle) != preconditioner.obstacle:
            raise ValueError("the preconditioner was built for another operator (timestep / obstacle differ)")
    else:
        raise NotImplementedError("fused solve: preconditioner must be (), Identity, Jacobi or Multigrid; "
                                  "use precond_conjugate_gradient for an arbitrary object")
    info = PcgInfo()
    check(_lib.load().pano_pcg_solve(kind, x.handle, b.handle, int(max_iterations), float(threshold),
                                     residual.handle, auxiliary.handle, search.handle, float(timestep), Rect(*obstacle),
                                     C.byref(info) if want_info else None))
    return info.as_dict() if want_info else None
