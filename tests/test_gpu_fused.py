"""-m gpu: the fused passes (advect_all, -divergence, Laplacian, projection) against the oracle.
Element-wise work is required to be BIT-EXACT (the library is built with --fmad=false), which is
stricter than the 1e-5 relative bound of BASELINE.json."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
# (h, w, vmax): vmax 30 -> <=1.5-cell backtrace (every border clamp), 200 -> 10 cells, 1e4 -> leaves any tile
CASES = [(2, 2, 50.0), (3, 3, 30.0), (5, 5, 30.0), (17, 33, 30.0), (17, 33, 200.0), (33, 17, 200.0),
         (128, 128, 30.0), (128, 128, 1e4), (257, 511, 200.0), (1024, 1024, 30.0), (64, 48, 1e12), (40, 40, 1e300),
         (70, 130, 1e-300)]


@pytest.mark.parametrize("h,w,vmax", CASES)
def test_advect_bit_exact(oracle, h, w, vmax):
    from tests import gpu_util as U
    from panopaea_b200 import fluid
    grid = U.grid(h, w)
    q, vel = U.rand_inputs(h, w, vmax, seed=1)
    Q, V = U.s2(grid, q), U.s1(grid, vel)
    dq, dv = grid.new_simplex_2(), grid.new_simplex_1()
    fluid.advect(dq, Q, 0.05, V)
    fluid.advect_mac(dv, V, 0.05, V)
    want_q, want_v = oracle.advect(h, w, q, 0.05, vel), oracle.advect_mac(h, w, vel, 0.05, vel)
    assert np.array_equal(dq.to_host(), want_q)
    assert np.array_equal(dv.view_linear(), want_v)
    # the one-pass kernel gives the same bits
    dq2, dv2 = grid.new_simplex_2(), grid.new_simplex_1()
    fluid.advect_all(dq2, dv2, Q, V, 0.05)
    assert np.array_equal(dq2.to_host(), want_q) and np.array_equal(dv2.view_linear(), want_v)


# The TMA-staged persistent advection kernel (pano_advect_tma.cu), forced with advect_kernel=4 on grids small enough for the
# oracle: vmax 30 -> backtraces below 2 cells (every gather from shared memory), 60 -> 3 cells (many cells fall back to global
# memory), 1e4 / 1e300 -> every cell falls back / the > 2^32-cell path; non-multiples of the tile (ragged tiles), the minimum size.
TMA_CASES = [(128, 256, 30.0), (256, 256, 30.0), (256, 256, 60.0), (130, 258, 30.0), (514, 390, 39.9), (514, 390, 200.0),
             (256, 512, 1e4), (128, 256, 1e300), (1024, 1024, 30.0), (1024, 1024, 45.0), (2048, 2048, 35.0)]


@pytest.mark.parametrize("dynamic", [1, 0])
@pytest.mark.parametrize("h,w,vmax", TMA_CASES)
def test_advect_tma_bit_exact(oracle, h, w, vmax, dynamic):
    from tests import gpu_util as U
    from panopaea_b200 import fluid
    if dynamic == 0 and h * w > 300000:
        pytest.skip("static tile lists: small cases only")
    grid = U.grid(h, w)
    q, vel = U.rand_inputs(h, w, vmax, seed=h + w)
    Q, V = U.s2(grid, q), U.s1(grid, vel)
    oracle.set_threading(oracle.ALL_PARALLEL)
    try:
        want_q, want_v = oracle.advect(h, w, q, 0.05, vel), oracle.advect_mac(h, w, vel, 0.05, vel)
    finally:
        oracle.set_threading(oracle.SERIAL)
    U.ctx().set_option("advect_kernel", 4)
    U.ctx().set_option("advect_dynamic", dynamic)
    try:
        for rep in range(3):                                  # successive launches alternate between the two tile counters
            dq, dv = grid.new_simplex_2(), grid.new_simplex_1()
            dq.fill(-7.0)
            dv.fill(-7.0)
            n0 = U.ctx().launch_count()
            fluid.advect_all(dq, dv, Q, V, 0.05)
            assert np.array_equal(dq.to_host(), want_q), rep
            assert np.array_equal(dv.view_linear(), want_v), rep
    finally:
        U.ctx().set_option("advect_kernel", 0)
        U.ctx().set_option("advect_dynamic", 1)


def test_advect_tma_smooth_flow_and_auto_choice(oracle):
    """A smooth flow (the regime of the smoke plume: neighbouring cells trace back to neighbouring cells) on a grid where the
    auto choice is the TMA kernel, and the same answer from the marching kernel."""
    from tests import gpu_util as U
    from panopaea_b200 import fluid
    h, w = 1536, 2048
    grid = U.grid(h, w)
    yy, xx = np.mgrid[0:h + 1, 0:w + 1]
    vy = 30.0 * np.sin(xx[:, :w] / 37.0) * np.cos(yy[:, :w] / 53.0)
    vx = 30.0 * np.cos(xx[:h, :] / 41.0 + 1.0) * np.sin(yy[:h, :] / 29.0)
    vel = oracle.join(vy, vx)
    q = np.sin(xx[:h, :w] / 11.0) + np.cos(yy[:h, :w] / 7.0)
    Q, V = U.s2(grid, q), U.s1(grid, vel)
    out = []
    for kernel in (0, 3):
        U.ctx().set_option("advect_kernel", kernel)
        try:
            dq, dv = grid.new_simplex_2(), grid.new_simplex_1()
            fluid.advect_all(dq, dv, Q, V, 0.05)
            out.append((dq.to_host(), dv.view_linear().copy()))
        finally:
            U.ctx().set_option("advect_kernel", 0)
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1])
    oracle.set_threading(oracle.ALL_PARALLEL)
    try:
        assert np.array_equal(out[0][0], oracle.advect(h, w, q, 0.05, vel))
        assert np.array_equal(out[0][1], oracle.advect_mac(h, w, vel, 0.05, vel))
    finally:
        oracle.set_threading(oracle.SERIAL)


def test_advect_mac_distinct_source(oracle):
    """advect_mac(dst, src, dt, vel) with src != vel (the signature allows it; the example passes vel twice)."""
    from tests import gpu_util as U
    from panopaea_b200 import fluid
    h, w = 40, 56
    grid = U.grid(h, w)
    _, vel = U.rand_inputs(h, w, 60.0, seed=2)
    _, src = U.rand_inputs(h, w, 1.0, seed=3)
    dv = grid.new_simplex_1()
    fluid.advect_mac(dv, U.s1(grid, src), 0.05, U.s1(grid, vel))
    assert np.array_equal(dv.view_linear(), oracle.advect_mac(h, w, src, 0.05, vel))


def test_advect_f32(oracle):
    from tests import gpu_util as U
    from panopaea_b200 import fluid
    h, w = 31, 45
    grid = U.grid(h, w)
    q, vel = U.rand_inputs(h, w, 30.0, seed=4, dtype=np.float32)
    dq, dv = grid.new_simplex_2(np.float32), grid.new_simplex_1(np.float32)
    fluid.advect_all(dq, dv, U.s2(grid, q, np.float32), U.s1(grid, vel, np.float32), 0.05)
    assert np.array_equal(dq.to_host(), oracle.advect(h, w, q, 0.05, vel))
    assert np.array_equal(dv.view_linear(), oracle.advect_mac(h, w, vel, 0.05, vel))


def test_committed_golden_vectors():
    """tests/golden/oracle_vectors.npz (oracle outputs frozen at commit time)."""
    from tests import gpu_util as U
    from tests.golden.make_golden import inputs
    from panopaea_b200 import fluid
    z = np.load(os.path.join(GOLD, "oracle_vectors.npz"))
    for (h, w, vmax) in [(5, 5, 30.0), (3, 3, 30.0), (17, 33, 30.0), (17, 33, 200.0), (33, 17, 200.0)]:
        q, vel = inputs(h, w, vmax)
        tag = f"{h}x{w}_v{int(vmax)}"
        grid = U.grid(h, w)
        Q, V = U.s2(grid, q), U.s1(grid, vel)
        dq, dv, lap = grid.new_simplex_2(), grid.new_simplex_1(), grid.new_simplex_2()
        fluid.advect_all(dq, dv, Q, V, 0.05)
        assert np.array_equal(dq.to_host(), z[f"advect_{tag}"])
        assert np.array_equal(dv.view_linear(), z[f"advect_mac_{tag}"])
        fluid.laplacian_apply(lap, Q, 0.05, (h // 2, min(h, h // 2 + 2), w // 3, min(w, w // 3 + 3)))
        assert np.array_equal(lap.to_host(), z[f"lap_{tag}"])


@pytest.mark.parametrize("h,w", [(2, 2), (3, 3), (5, 7), (17, 33), (128, 128), (300, 200), (1024, 1024)])
def test_divergence_laplacian_project_bit_exact(oracle, h, w):
    from tests import gpu_util as U
    from panopaea_b200 import fluid
    grid = U.grid(h, w)
    q, vel = U.rand_inputs(h, w, 5.0, seed=7)
    for obstacle in [(0, 0, 0, 0), U.default_obstacle(h, w), (0, 1, 0, 2), (h - 1, h, w - 2, w)]:
        y0, y1, x0, x1 = obstacle
        # -divergence (dec_fluid.rs:69-83)
        e = oracle.hodge_1_dual(h, w, vel)
        ey, ex = oracle.split(e, h, w)
        ey[y0:y1, x0:x1] = 0
        ex[y0:y1, x0:x1] = 0
        want_b = -oracle.derivative_1_primal(h, w, e)
        B = grid.new_simplex_2()
        bmax = fluid.neg_divergence(B, U.s1(grid, vel), obstacle)
        assert np.array_equal(B.to_host(), want_b)
        assert bmax == oracle.norm_max(want_b)
        # Laplacian closure (dec_fluid.rs:100-119)
        Z = grid.new_simplex_2()
        fluid.laplacian_apply(Z, U.s2(grid, q), 0.05, obstacle)
        assert np.array_equal(Z.to_host(), oracle.laplacian_closure(h, w, q, 0.05, obstacle))
    # projection + walls (dec_fluid.rs:124-141), composed in the oracle exactly as the example does
    vt = oracle.derivative_0_dual(h, w, oracle.hodge_2_primal(h, w, q.ravel()))
    want = oracle.scaled_add(vel, 0.05, vt)
    vy, vx = oracle.split(want, h, w)
    vx[:, 0] = 0
    vx[:, w] = 0
    vy[0, :] = 0
    vy[h, :] = 0
    V = U.s1(grid, vel)
    fluid.project(V, U.s2(grid, q), 0.05)
    assert np.array_equal(V.view_linear(), want)


def test_density_to_u8():
    from tests import gpu_util as U
    from panopaea_b200 import fluid
    h, w = 20, 30
    grid = U.grid(h, w)
    d = np.random.default_rng(8).uniform(-3, 3, (h, w))
    img = fluid.density_to_u8(U.s2(grid, d), -2.0, 2.0)
    want = ((np.clip(d, -2.0, 2.0) - (-2.0)) / 4.0 * 255.0).astype(np.uint8)[::-1]
    assert np.array_equal(img, want)


def test_export_png_roundtrip(tmp_path):
    """the output step after the path (dec_fluid.rs:143-164): device transfer + flip, PNG encode on the host"""
    import struct
    import zlib
    from tests import gpu_util as U
    from panopaea_b200 import fluid
    h, w = 24, 40
    grid = U.grid(h, w)
    d = np.random.default_rng(10).uniform(-3, 3, (h, w))
    img = fluid.density_to_u8(U.s2(grid, d))
    path = tmp_path / "density_0.png"
    fluid.export_png(str(path), img)
    raw = path.read_bytes()
    assert raw[:8] == b"\x89PNG\r\n\x1a\n"
    assert struct.unpack(">II", raw[16:24]) == (w, h)
    i = raw.index(b"IDAT")
    n = struct.unpack(">I", raw[i - 4:i])[0]
    pix = np.frombuffer(zlib.decompress(raw[i + 4:i + 4 + n]), np.uint8).reshape(h, 1 + 3 * w)[:, 1:].reshape(h, w, 3)
    assert np.array_equal(pix[:, :, 0], img) and np.array_equal(pix[:, :, 1], img) and np.array_equal(pix[:, :, 2], img)


def test_full_size_properties():
    """4096^2 (BASELINE configs[3]): size-independent properties instead of an oracle run."""
    from tests import gpu_util as U
    from panopaea_b200 import fluid
    n = 4096
    grid = U.grid(n, n)
    k = n // 128
    obstacle = (70 * k, 80 * k, 50 * k, 70 * k)
    rng = np.random.default_rng(9)
    p = rng.normal(size=(n, n))
    q = rng.normal(size=(n, n))
    P_, Q_ = U.s2(grid, p), U.s2(grid, q)
    Ap, Aq = grid.new_simplex_2(), grid.new_simplex_2()
    fluid.laplacian_apply(Ap, P_, 0.05, obstacle)
    fluid.laplacian_apply(Aq, Q_, 0.05, obstacle)
    # symmetry <Ap, q> == <p, Aq>, null space (constants), rows sum to zero
    lhs, rhs = Ap.dot_linear(Q_), P_.dot_linear(Aq)
    assert lhs == pytest.approx(rhs, rel=1e-9)
    ap = Ap.to_host()
    assert abs(ap.sum()) < 1e-6 * np.abs(ap).sum()
    assert not ap[70 * k + 1:80 * k - 1, 50 * k + 1:70 * k - 1].any()          # cells sealed inside the obstacle
    ones = U.s2(grid, np.ones((n, n)))
    fluid.laplacian_apply(Aq, ones, 0.05, obstacle)
    assert Aq.norm_max() == 0.0
    # linearity: A(2p) == 2 A(p) exactly (power-of-two scaling is exact in binary floating point)
    P_.scale(2.0)
    fluid.laplacian_apply(Aq, P_, 0.05, obstacle)
    assert np.array_equal(Aq.to_host(), 2.0 * ap)
    # advecting a constant field with zero velocity at the walls returns the constant up to a few ulp
    vel = grid.new_simplex_1()
    vel.upload(rng.uniform(-30, 30, grid.num_elem_1()))
    c = U.s2(grid, np.full((n, n), 3.25))
    out = grid.new_simplex_2()
    fluid.advect(out, c, 0.05, vel)
    assert np.allclose(out.to_host(), 3.25, rtol=2e-15, atol=0)
    # divergence of the projected gradient field: -div(project(0, p)) == A(p)/dt * dt  (no obstacle in project)
    vel.fill(0.0)
    P_.scale(0.5)
    fluid.project(vel, P_, 0.05)
    b = grid.new_simplex_2()
    fluid.neg_divergence(b, vel, (0, 0, 0, 0))
    fluid.laplacian_apply(Aq, P_, 0.05, (0, 0, 0, 0))
    # project gives v = dt*grad; -div v = -dt*lap = -A(p) away from the walls' zeroed edges (already closed in A)
    assert np.allclose(b.to_host(), -Aq.to_host(), rtol=1e-12, atol=1e-12)
