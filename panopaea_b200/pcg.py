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
    if preconditioner not in (None, ()) and not isinstance(preconditioner, Identity):
        raise NotImplementedError("only the identity preconditioner `()` exists in the reference (pcg.rs:8-12)")
    info = PcgInfo()
    check(_lib.load().pano_pcg_solve(_lib.PRECOND_IDENTITY, x.handle, b.handle, int(max_iterations), float(threshold),
                                     residual.handle, auxiliary.handle, search.handle, float(timestep), Rect(*obstacle),
                                     C.byref(info) if want_info else None))
    return info.as_dict() if want_info else None
