"""-m gpu: every Manifold2d operator and flat-view primitive of the C ABI against the oracle,
starting with the reference's own unit tests (panopaea/src/dec/grid.rs:428-534) re-run on the GPU."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SIZES = [(1, 1), (2, 3), (3, 3), (5, 5), (17, 33), (128, 128), (200, 1000)]


@pytest.fixture(scope="module")
def ref():
    with open(os.path.join(GOLD, "reference_vectors.json")) as f:
        return json.load(f)


def test_ref_grid_2d_divergence(ref):
    """grid_2d_divergence (dec/grid.rs:428-482): f32, hodge_1_dual -> derivative_1_primal, eps 1e-3."""
    from tests import gpu_util as U
    g = ref["divergence_5x5_f32"]
    grid = U.grid(5, 5)
    vel = grid.new_simplex_1(np.float32).upload_split(g["vy"], g["vx"])
    vel_primal = grid.new_simplex_1(np.float32)
    divergence = grid.new_simplex_2(np.float32)
    grid.hodge_1_dual(vel_primal, vel)
    grid.derivative_1_primal(divergence, vel_primal)
    assert np.all(np.abs(divergence.view_linear() - np.array(g["div"], np.float32)) < g["eps"])


def test_ref_grid_2d_laplacian(ref):
    """grid_2d_laplacian (dec/grid.rs:484-515): f64, the four-operator chain, eps 1e-3."""
    from tests import gpu_util as U
    g = ref["laplacian_3x3_f64"]
    grid = U.grid(3, 3)
    faces_primal = U.s2(grid, np.array(g["faces"]))
    faces_dual, edges_dual, edges_primal, laplacian = grid.new_simplex_2(), grid.new_simplex_1(), grid.new_simplex_1(), grid.new_simplex_2()
    grid.hodge_2_primal(faces_dual, faces_primal)
    grid.derivative_0_dual(edges_dual, faces_dual)
    grid.hodge_1_dual(edges_primal, edges_dual)
    grid.derivative_1_primal(laplacian, edges_primal)
    assert np.all(np.abs(laplacian.view_linear() - np.array(g["laplacian"])) < g["eps"])
    # and the fused operator with dt = 1, no obstacle
    from panopaea_b200 import fluid
    z = grid.new_simplex_2()
    fluid.laplacian_apply(z, faces_primal, 1.0)
    assert np.array_equal(z.view_linear(), np.array(g["laplacian"]))


def test_ref_grid_2d_gradient(ref):
    """grid_2d_gradient (dec/grid.rs:517-534) asserts nothing in the reference; derived values here."""
    from tests import gpu_util as U
    g = ref["gradient_3x3_f64_derived"]
    grid = U.grid(3, 3)
    gradient = grid.new_simplex_1()
    grid.derivative_0_dual(gradient, U.s2(grid, np.array(g["faces"])))
    e0, e1 = gradient.split()
    assert np.array_equal(e0, np.array(g["e0"], float)) and np.array_equal(e1, np.array(g["e1"], float))


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("h,w", SIZES)
def test_operators_bit_exact(oracle, dtype, h, w):
    from tests import gpu_util as U
    grid = U.grid(h, w)
    rng = np.random.default_rng(h * 1000 + w)
    e = rng.normal(size=grid.num_elem_1()).astype(dtype)
    f = rng.normal(size=(h, w)).astype(dtype)
    v = rng.normal(size=(h + 1, w + 1)).astype(dtype)
    E, F, V = U.s1(grid, e, dtype), U.s2(grid, f, dtype), U.s0(grid, v, dtype)
    out1, out2, out0 = grid.new_simplex_1(dtype), grid.new_simplex_2(dtype), grid.new_simplex_0(dtype)

    grid.hodge_1_dual(out1, E)
    assert np.array_equal(out1.view_linear(), oracle.hodge_1_dual(h, w, e))
    grid.hodge_1_primal(out1, E)
    assert np.array_equal(out1.view_linear(), oracle.hodge_1_primal(h, w, e))
    grid.hodge_2_primal(out2, F)
    assert np.array_equal(out2.to_host(), f)
    grid.hodge_0_dual(out2, F)
    assert np.array_equal(out2.to_host(), f)
    grid.derivative_1_primal(out2, E)
    assert np.array_equal(out2.to_host(), oracle.derivative_1_primal(h, w, e))
    grid.derivative_0_primal(out1, V)
    assert np.array_equal(out1.view_linear(), oracle.derivative_0_primal(h, w, v))
    # derivative_0_dual leaves boundary edges untouched: pre-fill with a sentinel
    out1.fill(7.0)
    grid.derivative_0_dual(out1, F)
    expect = oracle.derivative_0_dual(h, w, f, out=np.full(grid.num_elem_1(), 7.0, dtype))
    assert np.array_equal(out1.view_linear(), expect)
    # Hodge<Simplex0> with its addressing quirk; untouched entries keep the sentinel
    for name in ("hodge_0_primal", "hodge_2_dual"):
        out0.fill(-3.0)
        getattr(grid, name)(out0, V)
        expect0 = getattr(oracle, name)(h, w, v, out=np.full((h + 1, w + 1), -3.0, dtype))
        assert np.array_equal(out0.to_host(), expect0), name


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_flat_view_ops(oracle, dtype):
    from tests import gpu_util as U
    grid = U.grid(37, 53)
    rng = np.random.default_rng(5)
    a = rng.normal(size=(37, 53)).astype(dtype)
    b = rng.normal(size=(37, 53)).astype(dtype)
    A, B = U.s2(grid, a, dtype), U.s2(grid, b, dtype)
    A.scaled_add(0.37, B)
    a = oracle.scaled_add(a, 0.37, b).reshape(37, 53)
    assert np.array_equal(A.to_host(), a)
    A.xpby(B, -1.25)
    a = (b + dtype(-1.25) * a).astype(dtype)
    assert np.array_equal(A.to_host(), a)
    A.scale(-1.0)
    a = -a
    assert np.array_equal(A.to_host(), a)
    tol = 1e-12 if dtype == np.float64 else 2e-3
    assert A.dot_linear(B) == pytest.approx(float(np.dot(a.ravel().astype(np.float64), b.ravel().astype(np.float64))), rel=tol, abs=tol)
    assert A.norm_max() == oracle.norm_max(a)
    A.fill(2.5)
    assert np.all(A.to_host() == 2.5)
    A.fill(0.0)
    assert not A.to_host().any()
    A.fill_rect((3, 9, 10, 20), 4.0)
    ref = np.zeros((37, 53), dtype)
    ref[3:9, 10:20] = 4.0
    assert np.array_equal(A.to_host(), ref)
    # Simplex1: same (y, x) rectangle in vy and vx, or one component
    E = grid.new_simplex_1(dtype)
    E.fill_rect((3, 9, 10, 20), 1.0)
    vy, vx = E.split()
    assert vy[3:9, 10:20].all() and vx[3:9, 10:20].all() and vy.sum() == 60 and vx.sum() == 60
    from panopaea_b200 import _lib
    E.fill_rect((0, 2, 0, 53), 5.0, _lib.COMP_VY)
    vy, vx = E.split()
    assert np.all(vy[0:2] == 5.0) and vx.sum() == 60
    # swap is O(1) and exchanges contents
    A.swap(B)
    assert np.array_equal(A.to_host(), b) and np.array_equal(B.to_host(), ref)


def test_large_reductions(oracle):
    from tests import gpu_util as U
    grid = U.grid(1024, 1024)
    rng = np.random.default_rng(6)
    a = rng.normal(size=(1024, 1024))
    b = rng.normal(size=(1024, 1024))
    A, B = U.s2(grid, a), U.s2(grid, b)
    assert A.dot_linear(B) == pytest.approx(oracle.dot_linear(a, b), rel=1e-10)
    assert A.norm_max() == oracle.norm_max(a)
    # deterministic run to run
    assert A.dot_linear(B) == A.dot_linear(B)


def test_errors_mirror_reference_panics():
    from tests import gpu_util as U
    import panopaea_b200 as P
    from panopaea_b200 import _lib
    g1, g2 = U.grid(4, 4), U.grid(4, 5)
    with pytest.raises(P.PanoError) as e:
        g1.derivative_1_dual(g1.new_simplex_0(), g1.new_simplex_1())      # unimplemented!() dec/grid.rs:308-312
    assert e.value.code == _lib.ERR_UNIMPLEMENTED
    with pytest.raises(P.PanoError) as e:
        g1.hodge_1_dual(g1.new_simplex_1(), g2.new_simplex_1())           # ndarray Zip shape panic
    assert e.value.code == _lib.ERR_SHAPE
    with pytest.raises(P.PanoError) as e:
        g1.derivative_1_primal(g1.new_simplex_1(), g1.new_simplex_1())    # wrong simplex kind
    assert e.value.code == _lib.ERR_SHAPE
    with pytest.raises(P.PanoError) as e:
        g1.new_simplex_2().fill_rect((0, 5, 0, 4), 1.0)                   # index out of bounds
    assert e.value.code == _lib.ERR_SHAPE
    with pytest.raises(P.PanoError):
        g1.new_simplex_2().upload(np.zeros(15))
    with pytest.raises(P.PanoError) as e:
        g1.hodge_1_dual(g1.new_simplex_1(np.float32), g1.new_simplex_1(np.float64))
    assert e.value.code == _lib.ERR_SHAPE
