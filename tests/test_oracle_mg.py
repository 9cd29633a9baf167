"""The preconditioner specification (oracle/pano_oracle_mg.inc, C) against its independent numpy statement
(oracle/np_oracle.py) and against the properties a CG preconditioner must have.  The reference has no
preconditioner besides `()` (pcg.rs:8-12), so these two files ARE the definition the CUDA path is held to."""
import numpy as np
import pytest

from oracle import np_oracle as NP

SHAPES = [(128, 128, (70, 80, 50, 70)), (17, 33, (5, 9, 10, 20)), (300, 200, (100, 130, 50, 90)), (9, 7, (0, 0, 0, 0)),
          (8, 8, (2, 4, 2, 4)), (1, 40, (0, 0, 0, 0)), (33, 2, (0, 0, 0, 0)), (65, 129, (0, 65, 60, 61))]


@pytest.mark.parametrize("h,w,ob", SHAPES)
def test_vcycle_matches_numpy_statement(oracle, h, w, ob):
    m, P = oracle.Multigrid(h, w, 0.05, ob), NP.Multigrid(h, w, 0.05, ob)
    assert m.levels == len(P.W)
    for l in range(m.levels):
        wy, wx = m.level_weights(l)
        assert np.array_equal(wy, P.W[l][0]) and np.array_equal(wx, P.W[l][1])      # dyadic weights: exact
    r = np.random.default_rng(0).normal(size=(h, w))
    z, z2 = m.apply(r), P.apply(r)
    assert np.abs(z - z2).max() <= 1e-13 * max(1e-300, np.abs(z2).max())
    j, j2 = m.jacobi(r), P.jacobi(r)
    assert np.abs(j - j2).max() <= 1e-15 * max(1e-300, np.abs(j2).max())


@pytest.mark.parametrize("h,w,ob", SHAPES[:5])
def test_vcycle_is_symmetric_positive_and_linear(oracle, h, w, ob):
    m = oracle.Multigrid(h, w, 0.05, ob)
    rng = np.random.default_rng(1)
    a, b = rng.normal(size=(h, w)), rng.normal(size=(h, w))
    Ma, Mb = m.apply(a), m.apply(b)
    assert abs((Ma * b).sum() - (a * Mb).sum()) <= 1e-10 * abs((Ma * b).sum())      # <Ma, b> = <a, Mb>
    assert (Ma * a).sum() > 0 and (Mb * b).sum() > 0
    assert np.array_equal(m.apply(4.0 * a), 4.0 * Ma)                               # exact for a power of two


def test_pcg_with_multigrid_converges_in_a_few_iterations(oracle):
    n = 256
    prm = oracle.smoke_params(n)
    S = oracle.FluidState(**prm)
    for _ in range(3):
        S.step()
    b = S.field("temp").reshape(n, n).copy()          # the rhs of the last solve
    S.close()
    ident = oracle.pcg_grid_laplacian(n, n, b, 400, 0.1, 0.05, prm["obstacle"])
    same = oracle.pcg_grid_laplacian_precond(n, n, b, 400, 0.1, 0.05, prm["obstacle"], "identity")
    assert same.iterations == ident.iterations and np.array_equal(same.x, ident.x)   # apply = NULL is pcg.rs:8-12
    mg = oracle.pcg_grid_laplacian_precond(n, n, b, 400, 0.1, 0.05, prm["obstacle"], "multigrid")
    jac = oracle.pcg_grid_laplacian_precond(n, n, b, 400, 0.1, 0.05, prm["obstacle"], "jacobi")
    assert ident.iterations > 100 and jac.iterations > 100
    assert 0 <= mg.iterations <= 4
    res = b - oracle.laplacian_closure(n, n, mg.x, 0.05, prm["obstacle"])
    assert np.abs(res).max() < 0.1 and abs(np.abs(res).max() - mg.final_residual) < 1e-9
    tight = oracle.pcg_grid_laplacian_precond(n, n, b, 400, 1e-8, 0.05, prm["obstacle"], "multigrid")
    assert tight.iterations <= 16
