"""-m gpu: the Jacobi / multigrid preconditioners (csrc/pano_mg.cu) against their specification
(oracle/pano_oracle_mg.inc).  apply() is element-wise deterministic arithmetic, so it must be BIT-EXACT;
the preconditioned solve differs only in the summation order of the dot products."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SHAPES = [(128, 128, (70, 80, 50, 70)), (17, 33, (5, 9, 10, 20)), (300, 200, (100, 130, 50, 90)), (9, 7, (0, 0, 0, 0)),
          (8, 8, (2, 4, 2, 4)), (1, 40, (0, 0, 0, 0)), (33, 2, (0, 0, 0, 0)), (65, 129, (0, 65, 60, 61)), (32, 32, (0, 0, 0, 0)),
          (33, 33, (3, 30, 3, 30)), (1024, 1024, (560, 640, 400, 560)), (1000, 1500, (1, 999, 700, 701))]


@pytest.mark.parametrize("h,w,ob", SHAPES)
def test_multigrid_apply_bit_exact(oracle, h, w, ob):
    from tests import gpu_util as U
    from panopaea_b200 import pcg
    grid = U.grid(h, w)
    r = np.random.default_rng(7).normal(size=(h, w)) * 3.0
    R, Z = U.s2(grid, r), grid.new_simplex_2()
    M = pcg.Multigrid(grid, 0.05, ob)
    ref = oracle.Multigrid(h, w, 0.05, ob)
    assert M.levels()[0] == ref.levels
    M.apply(Z, R)
    assert np.array_equal(Z.to_host(), ref.apply(r))
    M.apply(Z, R)                                       # the object is reusable, the result does not depend on its scratch
    assert np.array_equal(Z.to_host(), ref.apply(r))
    J = pcg.Jacobi(grid, 0.05, ob)
    J.apply(Z, R)
    assert np.array_equal(Z.to_host(), ref.jacobi(r))
    M.close()


def test_apply_rejects_aliasing_and_foreign_grids():
    from tests import gpu_util as U
    import panopaea_b200 as P
    from panopaea_b200 import pcg
    grid = U.grid(40, 40)
    M = pcg.Multigrid(grid, 0.05)
    a = grid.new_simplex_2()
    with pytest.raises(P.PanoError):
        M.apply(a, a)
    other = U.grid(41, 40).new_simplex_2()
    with pytest.raises(P.PanoError):
        M.apply(other, other)
    with pytest.raises(P.PanoError):
        pcg.Multigrid(grid, 0.05, (0, 41, 0, 1))       # rectangle outside the grid: an index panic in the reference
    M.close()


@pytest.mark.parametrize("n,kind", [(128, "multigrid"), (256, "multigrid"), (1024, "multigrid"), (128, "jacobi")])
def test_preconditioned_solve_matches_oracle(oracle, n, kind):
    """pano_pcg_solve with a non-identity kind == pcg.rs:14-82 with that Preconditioner object."""
    from tests import gpu_util as U
    from panopaea_b200 import pcg
    prm = oracle.smoke_params(n)
    S = oracle.FluidState(**prm)
    oracle.set_threading(oracle.ALL_PARALLEL)
    for _ in range(3):
        S.step()
    oracle.set_threading(oracle.SERIAL)
    b = S.field("temp").reshape(n, n).copy()
    S.close()
    ob = prm["obstacle"]
    want = oracle.pcg_grid_laplacian_precond(n, n, b, 300, 0.1, 0.05, ob, kind)
    grid = U.grid(n, n)
    x, B, r, aux, s = grid.new_simplex_2(), U.s2(grid, b), grid.new_simplex_2(), grid.new_simplex_2(), grid.new_simplex_2()
    P = (pcg.Multigrid if kind == "multigrid" else pcg.Jacobi)(grid, 0.05, ob)
    got = pcg.solve_grid_laplacian(x, B, 300, 0.1, r, aux, s, 0.05, ob, preconditioner=P)
    assert abs(got["iterations"] - want.iterations) <= 2, (got, want.iterations)
    if got["iterations"] == want.iterations:
        assert np.abs(x.to_host() - want.x).max() <= 1e-6 * np.abs(want.x).max()
    res = b - oracle.laplacian_closure(n, n, x.to_host(), 0.05, ob)
    assert np.abs(res).max() < 0.1
    assert abs(np.abs(res).max() - got["final_residual"]) <= 1e-9 * max(1.0, np.abs(b).max())
    # the generic driver with the same object composes exactly like the Rust crate (trait object + closure)
    from panopaea_b200 import fluid
    x2 = grid.new_simplex_2()
    info2 = pcg.precond_conjugate_gradient(P, x2, B, 300, 0.1, r, aux, s, lambda dst, src: fluid.laplacian_apply(dst, src, 0.05, ob))
    assert info2["iterations"] == got["iterations"]
    assert np.array_equal(x2.to_host(), x.to_host())
    if kind == "multigrid":
        assert got["iterations"] <= 4
        P.close()


def test_step_with_multigrid_converges_where_identity_does_not(oracle):
    """dec_fluid's loop with precond = multigrid: every solve reaches the threshold (the reference's CG at 512^2
    stops at its 100-iteration cap with max|r| still above it), and the projected field is divergence-free to it."""
    from tests import gpu_util as U
    from panopaea_b200 import _lib, fluid
    n = 512
    sim = fluid.DecFluid(**fluid.smoke_params(n), ctx=U.ctx())
    sim.params.precond = _lib.PRECOND_MULTIGRID
    plain = fluid.DecFluid(**fluid.smoke_params(n), ctx=U.ctx())
    for _ in range(4):
        a, b = sim.step(), plain.step()
    assert 0 <= a["iterations"] <= 5 and a["final_residual"] < 0.1
    assert b["iterations"] == 100 and b["final_residual"] > 0.1
    div = sim.grid.new_simplex_2()
    worst = fluid.neg_divergence(div, sim.vel, fluid.smoke_params(n)["obstacle"])
    assert worst < 0.1 + 1e-9


def test_host_buffer_step_with_multigrid():
    """pano_fluid_step_host with precond = multigrid gives the fields of the device-resident step with the same
    preconditioner (same kernels, same order), and reports the host-driven loop's info."""
    import ctypes as C
    from tests import gpu_util as U
    from panopaea_b200 import _lib, fluid
    n = 256
    prm = fluid.smoke_params(n)
    params = _lib.StepParams(prm["timestep"], prm["threshold"], prm["max_iterations"], _lib.PRECOND_MULTIGRID, _lib.Rect(*prm["inflow"]),
                             prm["inflow_density"], prm["inflow_vy"], _lib.Rect(*prm["obstacle"]))
    sim = fluid.DecFluid(**prm, ctx=U.ctx())
    sim.params.precond = _lib.PRECOND_MULTIGRID
    density, vel, pressure = np.zeros((n, n)), np.zeros((n + 1) * n + n * (n + 1)), np.zeros((n, n))
    L = _lib.load()
    for _ in range(4):
        info = _lib.PcgInfo()
        _lib.check(L.pano_fluid_step_host(U.ctx().handle, C.byref(params), n, n, density.ctypes.data_as(C.c_void_p),
                                          vel.ctypes.data_as(C.c_void_p), pressure.ctypes.data_as(C.c_void_p), C.byref(info)))
        want = sim.step()
        assert info.iterations == want["iterations"] and 0 <= info.iterations <= 4
        assert info.final_residual == want["final_residual"] < 0.1
        assert np.array_equal(density, sim.density.to_host())
        assert np.array_equal(vel, sim.vel.view_linear())
        assert np.array_equal(pressure, sim.pressure.to_host())


def test_multi_gpu_step_rejects_preconditioners():
    """The slab-decomposed step implements the reference's solver (identity) only; asking for more is an error, not a fallback."""
    from tests import gpu_util as U
    import panopaea_b200 as P
    from panopaea_b200 import _lib, dist
    prm = dict(timestep=0.05, threshold=0.1, max_iterations=20, inflow=(5, 20, 27, 32), inflow_density=1.0, inflow_vy=20.0, obstacle=(70, 80, 25, 35))
    params = dist.make_params(**prm)
    params.precond = _lib.PRECOND_MULTIGRID
    with pytest.raises(P.PanoError):
        dist.DistFluid(U.ctx(), 128, 64, 0, 2, params)
