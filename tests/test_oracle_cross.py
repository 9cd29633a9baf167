"""The C oracle against the independent numpy restatement (oracle/np_oracle.py):
the only pin available for advect, advect_mac, CG and the whole step, which the
reference does not test."""
import numpy as np
import pytest

from oracle import np_oracle as NP

CASES = [(5, 5, 30.0), (3, 3, 30.0), (17, 33, 30.0), (33, 17, 200.0), (2, 2, 50.0), (64, 48, 200.0)]


def _inputs(h, w, vmax, seed):
    rng = np.random.default_rng(seed)
    return rng.uniform(-1, 1, (h, w)), rng.uniform(-vmax, vmax, (h + 1) * w + h * (w + 1))


@pytest.mark.parametrize("h,w,vmax", CASES)
def test_advect_bit_exact(oracle, h, w, vmax):
    q, vel = _inputs(h, w, vmax, 3)
    vy, vx = oracle.split(vel, h, w)
    assert np.array_equal(oracle.advect(h, w, q, 0.05, vel), NP.advect(q, 0.05, vy, vx))


@pytest.mark.parametrize("h,w,vmax", CASES)
def test_advect_mac_bit_exact(oracle, h, w, vmax):
    _, vel = _inputs(h, w, vmax, 4)
    vy, vx = oracle.split(vel, h, w)
    dy, dx = NP.advect_mac(vy, vx, 0.05, vy, vx)
    oy, ox = oracle.split(oracle.advect_mac(h, w, vel, 0.05, vel), h, w)
    assert np.array_equal(oy, dy) and np.array_equal(ox, dx)


@pytest.mark.parametrize("h,w", [(3, 3), (5, 7), (16, 16), (17, 33)])
def test_laplacian_and_divergence(oracle, h, w):
    q, vel = _inputs(h, w, 5.0, 5)
    obstacle = (h // 2, min(h, h // 2 + 2), w // 3, min(w, w // 3 + 3))
    assert np.array_equal(oracle.laplacian_closure(h, w, q, 0.05, obstacle), NP.laplacian(q, 0.05, obstacle))
    vy, vx = oracle.split(vel, h, w)
    e = oracle.hodge_1_dual(h, w, vel)
    ey, ex = oracle.split(e, h, w)
    y0, y1, x0, x1 = obstacle
    ey[y0:y1, x0:x1] = 0
    ex[y0:y1, x0:x1] = 0
    assert np.array_equal(-oracle.derivative_1_primal(h, w, e), NP.neg_divergence(vy, vx, obstacle))


def test_laplacian_is_symmetric_psd(oracle):
    h, w = 9, 11
    obstacle = (3, 6, 2, 7)
    n = h * w
    A = np.zeros((n, n))
    for k in range(n):
        e = np.zeros(n)
        e[k] = 1.0
        A[:, k] = oracle.laplacian_closure(h, w, e.reshape(h, w), 1.0, obstacle).ravel()
    assert np.array_equal(A, A.T)
    assert np.all(np.abs(A.sum(axis=1)) == 0)
    assert np.linalg.eigvalsh(A).min() > -1e-12


def test_pcg_matches(oracle):
    h, w = 24, 20
    rng = np.random.default_rng(7)
    obstacle = (10, 14, 5, 12)
    # a right-hand side in the range of the (singular, Neumann) operator
    b = NP.laplacian(rng.normal(size=(h, w)) * 400, 0.05, obstacle)
    r = oracle.pcg_grid_laplacian(h, w, b, 100, 0.1, 0.05, obstacle)
    x, it, err = NP.pcg(b, 100, 0.1, lambda s: NP.laplacian(s, 0.05, obstacle))
    assert abs(r.iterations - it) <= 1
    assert 0 < r.iterations < 100 and r.applies == r.iterations + 1
    if r.iterations == it:
        assert np.allclose(r.x, x, rtol=1e-9, atol=1e-9)
    # early-out (pcg.rs:35-38): x stays zero
    r0 = oracle.pcg_grid_laplacian(h, w, b * 1e-4, 100, 0.1, 0.05, obstacle)
    assert r0.iterations == -1 and r0.applies == 0 and not r0.x.any()


def test_dec_fluid_steps(oracle):
    S = oracle.FluidState(**oracle.smoke_params(128))
    N = NP.FluidState(**oracle.smoke_params(128))
    for i in range(30):
        a, b = S.step(want_rhs=True), N.step()
        assert a["iterations"] == b["iterations"], i
        assert np.allclose(a["rhs"], b["rhs"], rtol=0, atol=1e-9)
    vy, vx = oracle.split(S.field("vel"), 128, 128)
    assert np.allclose(vy, N.vy, rtol=0, atol=1e-9) and np.allclose(vx, N.vx, rtol=0, atol=1e-9)
    assert np.allclose(S.field("density"), N.density, rtol=0, atol=1e-10)
