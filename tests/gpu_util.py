"""Shared helpers of the -m gpu parity tests (CUDA path through the C ABI vs the CPU oracle)."""
import numpy as np

import panopaea_b200 as P
from panopaea_b200 import fluid

_ctx = None


def ctx():
    global _ctx
    if _ctx is None:
        _ctx = P.Context(0)
    return _ctx


def grid(h, w):
    return P.Grid2d((h, w), ctx())


def s2(g, a=None, dtype=np.float64):
    f = g.new_simplex_2(dtype)
    if a is not None:
        f.upload(a)
    return f


def s1(g, a=None, dtype=np.float64):
    f = g.new_simplex_1(dtype)
    if a is not None:
        f.upload(a)
    return f


def s0(g, a=None, dtype=np.float64):
    f = g.new_simplex_0(dtype)
    if a is not None:
        f.upload(a)
    return f


def rand_inputs(h, w, vmax, seed=0, dtype=np.float64):
    rng = np.random.default_rng(seed)
    q = rng.uniform(-1.0, 1.0, (h, w)).astype(dtype)
    vel = rng.uniform(-vmax, vmax, (h + 1) * w + h * (w + 1)).astype(dtype)
    return q, vel


def consistent_rhs(oracle, h, w, obstacle, seed=0, scale=400.0, dt=0.05):
    """b in the range of the singular Neumann operator (so CG converges)."""
    rng = np.random.default_rng(seed)
    return oracle.laplacian_closure(h, w, rng.normal(size=(h, w)) * scale, dt, obstacle)


def default_obstacle(h, w):
    return (h // 2, min(h, h // 2 + max(2, h // 12)), w // 3, min(w, w // 3 + max(3, w // 6)))
