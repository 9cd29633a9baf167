"""-m gpu: the Grid3d path (panopaea/src/domain/grid.rs:17-20 + math/interp.rs:23-36, defined in DESIGN.md 5c) through
the C ABI against its CPU checker (oracle/pano_oracle3.inc, "parity unpinned": the reference has no 3-D fluid code)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _g3(d, h, w):
    import panopaea_b200 as P
    from tests import gpu_util as U
    return P.Grid3d((d, h, w), U.ctx())


def _close(a, b, rel):
    return np.abs(a - b).max() <= rel * max(np.abs(b).max(), 1e-300)


SHAPES = [(2, 2, 2), (3, 4, 5), (9, 17, 12), (16, 8, 33), (5, 40, 70), (33, 31, 32), (64, 64, 64)]


def test_trilinear_matches_reference_expression():
    from oracle import pano_oracle3 as O3
    from panopaea_b200 import grid3
    rng = np.random.default_rng(0)
    for _ in range(20):
        a = rng.uniform(-3, 3, 11)
        assert grid3.trilinear(*a) == O3.trilinear(*a)


@pytest.mark.parametrize("vmax", [30.0, 200.0, 1e4, 1e12, 1e300])
@pytest.mark.parametrize("d,h,w", SHAPES)
def test_advect3_bit_exact(d, h, w, vmax):
    from oracle import pano_oracle3 as O3
    from panopaea_b200 import grid3
    g = _g3(d, h, w)
    rng = np.random.default_rng(d * 1000 + h)
    q = rng.uniform(-1, 1, (d, h, w))
    vel = rng.uniform(-vmax, vmax, O3.num_faces(d, h, w))
    src = rng.uniform(-2, 2, vel.size)
    fq, fv, fs = g.new_cells().upload(q), g.new_faces().upload(vel), g.new_faces().upload(src)
    dq, dv = g.new_cells(), g.new_faces()
    grid3.advect(dq, fq, 0.05, fv)
    assert np.array_equal(dq.to_host(), O3.advect(d, h, w, q, 0.05, vel))
    grid3.advect_mac(dv, fs, 0.05, fv)
    assert np.array_equal(dv.view_linear(), O3.advect_mac(d, h, w, src, 0.05, vel))
    dq2, dv2 = g.new_cells(), g.new_faces()
    grid3.advect_all(dq2, dv2, fq, fv, 0.05)                   # the fused pass of the step: self-advection
    assert np.array_equal(dq2.to_host(), O3.advect(d, h, w, q, 0.05, vel))
    assert np.array_equal(dv2.view_linear(), O3.advect_mac(d, h, w, vel, 0.05, vel))


@pytest.mark.parametrize("d,h,w", SHAPES)
def test_divergence_laplacian_projection_bit_exact(d, h, w):
    from oracle import pano_oracle3 as O3
    from panopaea_b200 import grid3
    g = _g3(d, h, w)
    rng = np.random.default_rng(5)
    p = rng.uniform(-3, 3, (d, h, w))
    vel = rng.uniform(-5, 5, O3.num_faces(d, h, w))
    fp, fv, out = g.new_cells().upload(p), g.new_faces().upload(vel), g.new_cells()
    for ob in [(0,) * 6, (d // 2, min(d, d // 2 + 2), h // 2, min(h, h // 2 + 2), w // 3, min(w, w // 3 + 3)), (0, 1, 0, 2, 0, 2),
               (d - 1, d, h - 2, h, w - 1, w), (0, d, 0, h, 0, w)]:
        bmax = grid3.neg_divergence(out, fv, ob)
        want = O3.neg_divergence(d, h, w, vel, ob)
        assert np.array_equal(out.to_host(), want) and bmax == np.abs(want).max()
        grid3.laplacian_apply(out, fp, 0.05, ob)
        assert np.array_equal(out.to_host(), O3.laplacian_closure(d, h, w, p, 0.05, ob))
    grid3.project(fv, fp, 0.05)
    assert np.array_equal(fv.view_linear(), O3.project(d, h, w, vel, p, 0.05))


@pytest.mark.parametrize("kernel", [0, 1])
@pytest.mark.parametrize("zc", [0, 4, 64])
@pytest.mark.parametrize("d,h,w", [(3, 4, 5), (9, 17, 12), (16, 8, 33), (33, 31, 32), (64, 64, 64), (70, 40, 96), (20, 50, 130), (5, 16, 64), (2, 2, 2)])
def test_pcg3_matches_oracle(d, h, w, zc, kernel):
    """pcg.rs:14-82 with the 7-point closure: iteration count within +-2 (north_star), x / r / s as the reference leaves them.
    kernel 0: auto (shared-memory plane tiles for even widths, the column kernel otherwise); 1: the column kernel everywhere."""
    from oracle import np_oracle3 as NP3
    from oracle import pano_oracle3 as O3
    from panopaea_b200 import grid3
    from tests import gpu_util as U
    U.ctx().set_option("cg3_zc", zc)
    U.ctx().set_option("cg3_kernel", kernel)
    try:
        g = _g3(d, h, w)
        ob = (d // 2, min(d, d // 2 + 2), h // 2, min(h, h // 2 + 2), w // 3, min(w, w // 3 + 3))
        b = NP3.laplacian(np.random.default_rng(3).normal(size=(d, h, w)) * 400.0, 0.05, ob)
        for max_it, thr in ((100, 0.1), (7, 1e-9), (100, 1e30)):
            want = O3.pcg(d, h, w, b, max_it, thr, 0.05, ob)
            x, r, aux, s, fb = (g.new_cells() for _ in range(5))
            for f in (r, s):
                f.fill(7.0)                                    # the early-out must leave the scratch untouched
            fb.upload(b)
            info = grid3.pcg_solve(x, fb, max_it, thr, r, aux, s, 0.05, ob)
            assert abs(info["iterations"] - want.iterations) <= 2, (info, want.iterations)
            if want.iterations == -1:
                assert info["iterations"] == -1 and not x.to_host().any() and (r.to_host() == 7.0).all() and (s.to_host() == 7.0).all()
                continue
            assert info["rhs_max"] == np.abs(b).max()
            if info["iterations"] == want.iterations:
                assert info["applies"] == want.applies
                floor = 1e-9 * np.abs(b).max()                  # a converged residual is rounding noise: compare against the scale of b there
                assert _close(x.to_host(), want.x, 1e-6)
                assert np.abs(r.to_host() - want.residual).max() <= max(1e-6 * np.abs(want.residual).max(), floor)
                assert np.abs(s.to_host() - want.search).max() <= max(1e-6 * np.abs(want.search).max(), floor)
                assert abs(info["final_residual"] - want.final_residual) <= 1e-6 * max(1.0, want.final_residual)
                # the returned residual really is b - A x
                res = b - O3.laplacian_closure(d, h, w, x.to_host(), 0.05, ob)
                assert np.abs(res - r.to_host()).max() <= 1e-9 * np.abs(b).max()
    finally:
        U.ctx().set_option("cg3_zc", 0)
        U.ctx().set_option("cg3_kernel", 0)


def test_pcg3_rejects_what_the_2d_entry_points_reject():
    import panopaea_b200 as P
    from panopaea_b200 import fluid, grid3
    from tests import gpu_util as U
    g = _g3(4, 4, 4)
    c = [g.new_cells() for _ in range(5)]
    with pytest.raises(P.PanoError):
        grid3.pcg_solve(c[0], c[1], 10, 0.1, c[2], c[3], c[2], 0.05)          # aliasing
    with pytest.raises(P.PanoError):
        grid3.pcg_solve(c[0], c[1], 10, 0.1, c[2], c[3], c[4], 0.05, (0, 9, 0, 1, 0, 1))   # box outside the grid
    with pytest.raises(P.PanoError):
        grid3.pcg_solve(c[0], c[1], 10, 0.1, c[2], c[3], c[4], 0.05, precond=2)     # only `()` exists on a Grid3d
    g2 = U.grid(4, 4)
    with pytest.raises(P.PanoError):
        fluid.laplacian_apply(c[0], c[1], 0.05)                                   # a 2-D operator on 3-D fields
    with pytest.raises(P.PanoError):
        grid3.laplacian_apply(U.s2(g2), U.s2(g2), 0.05)
    with pytest.raises(P.PanoError):
        c[0].fill_box((0, 5, 0, 1, 0, 1), 1.0)


@pytest.mark.parametrize("n", [32, 64, 128])
def test_step3_resynchronised(n):
    """Per-step parity of the whole loop body on the 3-D smoke plume: the device state is overwritten with the checker's before
    every step.  Advection and -div bit-exact, iteration count within +-2, pressure / velocity within 1e-5 relative."""
    from oracle import pano_oracle as O
    from oracle import pano_oracle3 as O3
    from panopaea_b200 import grid3
    from tests import gpu_util as U
    sim = grid3.DecFluid3(**grid3.smoke_params(n), ctx=U.ctx())
    ref = O3.FluidState3(**O3.smoke_params(n))
    O.set_threading(O.ALL_PARALLEL if n >= 64 else O.SERIAL)
    try:
        for i in range(8 if n <= 64 else 3):
            sim.density.upload(ref.field("density"))
            sim.vel.upload(ref.field("vel"))
            g, o = sim.step(), ref.step(want_rhs=True)
            assert abs(g["iterations"] - o["iterations"]) <= 2, (i, g, o)
            assert np.array_equal(sim.density.to_host(), ref.field("density")), i
            assert g["rhs_max"] == np.abs(o["rhs"]).max()
            if g["iterations"] == o["iterations"]:
                assert _close(sim.pressure.to_host(), ref.field("pressure"), 1e-5), i
                assert _close(sim.vel.view_linear(), ref.field("vel"), 1e-5), i
        assert ref.field("density").max() > 0.5 and np.abs(ref.field("vel")).max() > 1.0
    finally:
        O.set_threading(O.SERIAL)


def test_handles_may_be_released_in_any_order():
    """pano_ctx_destroy with fields still alive: the context's device state lives until the last field goes (garbage-collected
    hosts release handles in no particular order; nothing may dangle across the boundary)."""
    import ctypes as C
    from panopaea_b200 import _lib
    L = _lib.load()
    ctx = C.c_void_p()
    assert L.pano_ctx_create(0, None, C.byref(ctx)) == 0
    f2, f3 = C.c_void_p(), C.c_void_p()
    assert L.pano_field_new(ctx, _lib.SIMPLEX2, _lib.F64, 64, 64, C.byref(f2)) == 0
    assert L.pano_field3_new(ctx, 3, 8, 8, 8, C.byref(f3)) == 0
    assert L.pano_ctx_destroy(ctx) == 0            # deferred: two fields alive
    assert L.pano_field_fill(f2, 3.0) == 0         # the fields still work
    out = C.c_double()
    assert L.pano_field_norm_max(f2, C.byref(out)) == 0 and out.value == 3.0
    assert L.pano_field_free(f2) == 0
    assert L.pano_field_free(f3) == 0              # the last one takes the context with it


def test_step3_256_vs_oracle():
    """The bench size of the Grid3d leg (256^3) against the all-parallel checker on identical inputs: the device runs a developed
    plume (10 free steps), its state is handed to the checker, and both advance it by one step -- advection and -div bit-exact,
    iteration count within +-2, pressure / velocity within 1e-5 relative (the bars of north_star)."""
    from oracle import pano_oracle as O
    from oracle import pano_oracle3 as O3
    from panopaea_b200 import grid3
    from tests import gpu_util as U
    n = 256
    sim = grid3.DecFluid3(**grid3.smoke_params(n), ctx=U.ctx())
    for _ in range(10):
        sim.step(want_info=False)
    ref = O3.FluidState3(**O3.smoke_params(n))
    ref.field("density")[...] = sim.density.to_host()
    ref.field("vel")[...] = sim.vel.view_linear()
    assert np.abs(ref.field("vel")).max() > 5.0 and ref.field("density").max() > 0.5        # a developed plume, not a zero field
    O.set_threading(O.ALL_PARALLEL)
    try:
        g, o = sim.step(), ref.step(want_rhs=True)
    finally:
        O.set_threading(O.SERIAL)
    assert abs(g["iterations"] - o["iterations"]) <= 2, (g, o)
    assert g["rhs_max"] == np.abs(o["rhs"]).max()
    assert np.array_equal(sim.density.to_host(), ref.field("density"))
    if g["iterations"] == o["iterations"]:
        assert _close(sim.pressure.to_host(), ref.field("pressure"), 1e-5)
        assert _close(sim.vel.view_linear(), ref.field("vel"), 1e-5)
        assert g["final_residual"] == pytest.approx(o["final_residual"], rel=1e-5)
    ref.close()


def test_fluid3_step_host_matches_resident_step():
    """pano_fluid3_step_host (host arrays in and out, density download under the solve) against the device-resident step: same bits."""
    from panopaea_b200 import grid3
    from tests import gpu_util as U
    n = 64
    prm = grid3.smoke_params(n)
    sim = grid3.DecFluid3(**prm, ctx=U.ctx())
    d = np.zeros((n, n, n))
    v = np.zeros(sim.grid.num_faces())
    p = np.zeros((n, n, n))
    for i in range(5):
        want = sim.step()
        got = grid3.fluid3_step_host(U.ctx(), sim.params, n, n, n, d, v, p if i % 2 == 0 else None)
        assert got == want, (i, got, want)
        assert np.array_equal(d, sim.density.to_host()) and np.array_equal(v, sim.vel.view_linear())
        if i % 2 == 0:
            assert np.array_equal(p, sim.pressure.to_host())
    assert d.max() > 0.5
