"""Pins the CPU oracle against the reference's own golden vectors
(panopaea/src/dec/grid.rs:428-482, 484-515) and the committed oracle fixtures."""
import json
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def ref():
    with open(os.path.join(GOLD, "reference_vectors.json")) as f:
        return json.load(f)


def test_grid_2d_divergence(oracle, ref):
    """reference test grid_2d_divergence: hodge_1_dual -> derivative_1_primal, f32, eps 1e-3."""
    g = ref["divergence_5x5_f32"]
    vel = oracle.join(np.array(g["vy"], np.float32), np.array(g["vx"], np.float32))
    div = oracle.derivative_1_primal(5, 5, oracle.hodge_1_dual(5, 5, vel))
    assert div.dtype == np.float32
    assert np.all(np.abs(div.ravel() - np.array(g["div"], np.float32)) < g["eps"])


def test_grid_2d_laplacian(oracle, ref):
    """reference test grid_2d_laplacian: hodge_2_primal -> derivative_0_dual -> hodge_1_dual -> derivative_1_primal."""
    g = ref["laplacian_3x3_f64"]
    p = np.array(g["faces"], np.float64)
    e = oracle.derivative_0_dual(3, 3, oracle.hodge_2_primal(3, 3, p.ravel()))
    lap = oracle.derivative_1_primal(3, 3, oracle.hodge_1_dual(3, 3, e))
    assert np.all(np.abs(lap.ravel() - np.array(g["laplacian"])) < g["eps"])
    # the fused closure with dt = 1 and no obstacle is the same operator
    assert np.array_equal(oracle.laplacian_closure(3, 3, p, 1.0), lap)


def test_grid_2d_gradient_derived(oracle, ref):
    """The reference's grid_2d_gradient asserts nothing; this pins the derived values."""
    g = ref["gradient_3x3_f64_derived"]
    e = oracle.derivative_0_dual(3, 3, np.array(g["faces"], np.float64))
    e0, e1 = oracle.split(e, 3, 3)
    assert np.array_equal(e0, np.array(g["e0"], np.float64))
    assert np.array_equal(e1, np.array(g["e1"], np.float64))


def test_derivative_0_dual_leaves_boundary_untouched(oracle):
    h, w = 4, 6
    out = np.full(oracle.num_elem_1(h, w), 7.0)
    oracle.derivative_0_dual(h, w, np.arange(h * w, dtype=np.float64), out=out)
    vy, vx = oracle.split(out, h, w)
    assert np.all(vy[0] == 7.0) and np.all(vy[h] == 7.0)
    assert np.all(vx[:, 0] == 7.0) and np.all(vx[:, w] == 7.0)
    assert not np.any(vy[1:h] == 7.0) and not np.any(vx[:, 1:w] == 7.0)


def test_hodge_1_pair_is_inverse(oracle):
    h, w = 3, 5
    e = np.random.default_rng(1).normal(size=oracle.num_elem_1(h, w))
    assert np.array_equal(oracle.hodge_1_dual(h, w, oracle.hodge_1_primal(h, w, e)), -e)
    vy, vx = oracle.split(oracle.hodge_1_dual(h, w, e), h, w)
    ey, ex = oracle.split(e, h, w)
    assert np.array_equal(vy, -ey) and np.array_equal(vx, ex)


def test_norm_max_and_dot(oracle):
    a = np.array([1.0, -7.5, 3.0, 0.0, 2.0, -1.0, 4.0, 6.0, -6.5, 1.5, 2.5])
    assert oracle.norm_max(a) == 7.5
    assert oracle.norm_max(np.zeros(0)) == 0.0
    b = np.arange(a.size, dtype=np.float64)
    assert oracle.dot_linear(a, b) == pytest.approx(float(a @ b), rel=1e-15)
    # 8-lane order: check against a literal restatement
    p = [0.0] * 8
    for i in range(8):
        p[i] += a[i] * b[i]
    s = 0.0
    for i in range(4):
        s += p[i] + p[i + 4]
    for i in range(8, a.size):
        s += a[i] * b[i]
    assert oracle.dot_linear(a, b) == s


def test_committed_fixtures(oracle):
    """oracle_vectors.npz (made by tests/golden/make_golden.py) still reproduces."""
    from tests.golden.make_golden import inputs
    z = np.load(os.path.join(GOLD, "oracle_vectors.npz"))
    for (h, w, vmax) in [(5, 5, 30.0), (3, 3, 30.0), (17, 33, 30.0), (17, 33, 200.0), (33, 17, 200.0)]:
        q, vel = inputs(h, w, vmax)
        tag = f"{h}x{w}_v{int(vmax)}"
        assert np.array_equal(oracle.advect(h, w, q, 0.05, vel), z[f"advect_{tag}"])
        assert np.array_equal(oracle.advect_mac(h, w, vel, 0.05, vel), z[f"advect_mac_{tag}"])
        obstacle = (h // 2, min(h, h // 2 + 2), w // 3, min(w, w // 3 + 3))
        assert np.array_equal(oracle.laplacian_closure(h, w, q, 0.05, obstacle), z[f"lap_{tag}"])
    S = oracle.FluidState(**oracle.smoke_params(128))
    its = [S.step()["iterations"] for _ in range(25)]
    assert its == list(z["dec_fluid_128_iterations"])
    assert np.array_equal(S.field("density")[::16], z["dec_fluid_128_density_rows"])
    assert np.array_equal(S.field("pressure")[::16], z["dec_fluid_128_pressure_rows"])


def test_threading_variants_agree(oracle):
    """reference-faithful threading is bit-identical to serial; all-parallel only reorders dots."""
    outs = []
    for mode in (oracle.SERIAL, oracle.REFERENCE_FAITHFUL, oracle.ALL_PARALLEL):
        oracle.set_threading(mode, 4)
        S = oracle.FluidState(**oracle.smoke_params(128))
        its = [S.step()["iterations"] for _ in range(5)]
        outs.append((its, S.field("density").copy(), S.field("vel").copy()))
    oracle.set_threading(oracle.SERIAL)
    assert outs[0][0] == outs[1][0]
    assert np.array_equal(outs[0][1], outs[1][1]) and np.array_equal(outs[0][2], outs[1][2])
    assert all(abs(a - b) <= 2 for a, b in zip(outs[0][0], outs[2][0]))
    assert np.allclose(outs[0][1], outs[2][1], rtol=0, atol=1e-7)
