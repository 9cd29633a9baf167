"""-m gpu: the slab-decomposed multi-GPU step, exercised on ONE GPU by running several ranks in this
process ("loop-back": each rank has its own context/stream and a share of the SMs; the ranks address
each other's windows directly instead of through CUDA IPC).  The code path -- ghost-row pushes with
flags, in-kernel halo stores, cross-rank reduction stage -- is the one the real multi-process run uses."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _make_ranks(n, h, w, prm):
    import panopaea_b200 as P
    from panopaea_b200 import dist
    ctxs = [P.Context(0) for _ in range(n)]
    sms = ctxs[0].num_sms()
    ranks = [dist.DistFluid(ctxs[r], h, w, r, n, prm) for r in range(n)]
    ptrs = [r.window()[0] for r in ranks]
    for r in ranks:
        r.connect_local(ptrs)
        r.set_max_ctas(max(1, sms // n))
    return ranks


def _step_all(ranks):
    for r in ranks:
        r.step()          # asynchronous: every rank's kernels are queued before anybody waits
    return [r.sync() for r in ranks]


# variants of the cross-rank protocol inside the CG kernel (pano_sm100.cuh): halo flags (default) with static or dynamic tile
# lists, halo tiles first with the flags raised early, and the fenced root exchange the flags replaced
# ... each with the two-reduction kernel (sr = 0, k_cg_stream) and with the single-reduction kernel (sr = 1, k_cg_sr: the default)
# sr = 2: the single-reduction kernel inside the fused-halo step (the default: ONE ghost-row exchange per step, the rest recomputed
# or mirrored from inside the solver); sr = 1: the same kernel with the four exchanges of the two-reduction path
# sr = 3: sr = 2 with "dist_overlap": interior rows advected while the ghost rows are in flight, edge strips after the wait
@pytest.mark.parametrize("dynamic,halo_first,xflags,sr", [(0, 0, 1, 2), (1, 0, 1, 2), (0, 0, 0, 2), (1, 0, 0, 2), (0, 0, 1, 3), (1, 0, 1, 3), (0, 0, 1, 1), (1, 0, 0, 1),
                                                          (0, 0, 1, 0), (1, 0, 1, 0), (0, 1, 1, 0), (0, 0, 0, 0), (1, 0, 0, 0), (0, 1, 0, 0)])
@pytest.mark.parametrize("nranks,h,w", [(2, 256, 256), (4, 256, 128), (3, 250, 192)])
def test_loopback_matches_single_gpu(nranks, h, w, dynamic, halo_first, xflags, sr):
    from tests import gpu_util as U
    from panopaea_b200 import dist, fluid
    k = 2
    prm = dict(timestep=0.05, threshold=0.1, max_iterations=100, inflow=(5 * k, 20 * k, 27 * k, 32 * k), inflow_density=1.0,
               inflow_vy=20.0, obstacle=(70 * k, 80 * k, 25 * k, 35 * k))
    single = fluid.DecFluid(h=h, w=w, ctx=U.ctx(), **prm)
    ranks = _make_ranks(nranks, h, w, prm)
    for r in ranks:
        r.ctx.set_option("cg_dynamic", dynamic)      # 1: tiles claimed from a counter (by default only above 24 tiles per CTA)
        r.ctx.set_option("cg_halo_first", halo_first)
        r.ctx.set_option("cg_xflags", xflags)
        r.ctx.set_option("cg_single_reduction", 1 if sr else 0)
        r.ctx.set_option("dist_fused_halos", 1 if sr >= 2 else 0)
        r.ctx.set_option("dist_overlap", 1 if sr == 3 else 0)
    for step in range(6):
        want = single.step()
        infos = _step_all(ranks)
        assert all(i == infos[0] for i in infos), "every rank must see the same solver outcome"
        assert abs(infos[0]["iterations"] - want["iterations"]) <= 1, (step, infos[0], want)
        # the first step is bit-exact outside the solver; later steps inherit the solver's reduction-order rounding
        if step == 0:
            assert infos[0]["rhs_max"] == want["rhs_max"]
        else:
            assert infos[0]["rhs_max"] == pytest.approx(want["rhs_max"], rel=1e-9)
        if infos[0]["iterations"] != want["iterations"]:
            break
        d = dist.gather_local(ranks, dist.DENSITY)
        if step == 0:
            assert np.array_equal(d, single.density.to_host())
        assert np.abs(d - single.density.to_host()).max() <= 1e-9
        vy, vx = single.vel.split()
        p = single.pressure.to_host()
        for got, ref in ((dist.gather_local(ranks, dist.VY), vy), (dist.gather_local(ranks, dist.VX), vx),
                         (dist.gather_local(ranks, dist.PRESSURE), p)):
            assert np.abs(got - ref).max() <= 1e-8 * max(1.0, np.abs(ref).max()), step


def test_loopback_against_oracle(oracle):
    from panopaea_b200 import dist
    n = 256
    k = 2
    prm = dict(h=n, w=n, timestep=0.05, threshold=0.1, max_iterations=100, inflow=(5 * k, 20 * k, 54 * k, 64 * k),
               inflow_density=1.0, inflow_vy=20.0, obstacle=(70 * k, 80 * k, 50 * k, 70 * k))
    ref = oracle.FluidState(**prm)
    ranks = _make_ranks(2, n, n, {kk: v for kk, v in prm.items() if kk not in ("h", "w")})
    for step in range(4):
        o = ref.step()
        infos = _step_all(ranks)
        assert abs(infos[0]["iterations"] - o["iterations"]) <= 2
        if step == 0:
            assert np.array_equal(dist.gather_local(ranks, dist.DENSITY), ref.field("density"))
        assert np.abs(dist.gather_local(ranks, dist.DENSITY) - ref.field("density")).max() <= 1e-7
        if infos[0]["iterations"] != o["iterations"]:
            break
        ovy, ovx = oracle.split(ref.field("vel"), n, n)
        assert np.abs(dist.gather_local(ranks, dist.VY) - ovy).max() <= 1e-5 * np.abs(ovy).max()
        assert np.abs(dist.gather_local(ranks, dist.PRESSURE) - ref.field("pressure")).max() <= 1e-5 * np.abs(ref.field("pressure")).max()


def test_upload_download_roundtrip():
    from panopaea_b200 import dist
    h, w = 96, 64
    prm = dict(inflow=(1, 2, 1, 2), obstacle=(0, 0, 0, 0))
    ranks = _make_ranks(3, h, w, prm)
    rng = np.random.default_rng(0)
    fields = {dist.DENSITY: rng.normal(size=(h, w)), dist.VY: rng.normal(size=(h + 1, w)), dist.VX: rng.normal(size=(h, w + 1))}
    for which, a in fields.items():
        for r in ranks:
            r.upload(which, a)
        assert np.array_equal(dist.gather_local(ranks, which), a)


def test_backtrace_longer_than_ghost_zone_is_reported():
    import panopaea_b200 as P
    from panopaea_b200 import dist
    h, w = 128, 64
    prm = dict(inflow=(1, 2, 1, 2), inflow_vy=0.0, inflow_density=0.0, obstacle=(0, 0, 0, 0))
    ranks = _make_ranks(2, h, w, prm)
    vy = np.full((h + 1, w), 400.0)          # 20 cells per step > 8 ghost rows
    for r in ranks:
        r.upload(dist.VY, vy)
    for r in ranks:
        r.step()
    errs = 0
    for r in ranks:
        try:
            r.sync()
        except P.PanoError:
            errs += 1
    assert errs >= 1


@pytest.mark.parametrize("sr", [1, 0])
def test_dist_solve_repeats_the_last_solve(sr):
    """pano_dist_solve (BASELINE configs[4], the Poisson solve alone): again on the right-hand side of the last step, from a
    zero guess (pcg.rs:32) -- the same bits as the solve inside the step."""
    from panopaea_b200 import dist
    k = 2
    prm = dict(timestep=0.05, threshold=0.1, max_iterations=100, inflow=(5 * k, 20 * k, 54 * k, 64 * k), inflow_density=1.0,
               inflow_vy=20.0, obstacle=(70 * k, 80 * k, 50 * k, 70 * k))
    ranks = _make_ranks(2, 256, 256, prm)
    for r in ranks:
        r.ctx.set_option("cg_single_reduction", sr)
    for _ in range(3):
        infos = _step_all(ranks)
    p0 = dist.gather_local(ranks, dist.PRESSURE)
    for _ in range(2):
        for r in ranks:
            r.solve()
        again = [r.sync() for r in ranks]
        assert again == infos
        assert np.array_equal(dist.gather_local(ranks, dist.PRESSURE), p0)


@pytest.mark.parametrize("nranks,h,w", [(2, 512, 512), (3, 768, 384)])
def test_loopback_tma_advection(nranks, h, w):
    """The slabs on the TMA-staged advection kernel (pano_advect_tma.cu; forced: the auto choice needs bigger slabs): tensor maps
    over the stored rows of the window, ghost rows as tile halos, vx super-rows paired relative to the window."""
    from tests import gpu_util as U
    from panopaea_b200 import dist, fluid
    k = h // 128
    kx = w // 128
    prm = dict(timestep=0.05, threshold=0.1, max_iterations=40, inflow=(5 * k, 20 * k, 54 * kx, 64 * kx), inflow_density=1.0,
               inflow_vy=20.0, obstacle=(70 * k, 80 * k, 50 * kx, 70 * kx))
    single = fluid.DecFluid(h=h, w=w, ctx=U.ctx(), **prm)
    ranks = _make_ranks(nranks, h, w, prm)
    for r in ranks:
        r.ctx.set_option("advect_kernel", 4)
    for step in range(5):
        want = single.step()
        infos = _step_all(ranks)
        assert abs(infos[0]["iterations"] - want["iterations"]) <= 1
        d = dist.gather_local(ranks, dist.DENSITY)
        if step == 0:
            assert np.array_equal(d, single.density.to_host())
            assert infos[0]["rhs_max"] == want["rhs_max"]
        assert np.abs(d - single.density.to_host()).max() <= 1e-9
        if infos[0]["iterations"] != want["iterations"]:
            break
        vy, vx = single.vel.split()
        assert np.abs(dist.gather_local(ranks, dist.VY) - vy).max() <= 1e-8 * max(1.0, np.abs(vy).max())
        assert np.abs(dist.gather_local(ranks, dist.VX) - vx).max() <= 1e-8 * max(1.0, np.abs(vx).max())


@pytest.mark.parametrize("dynamic", [0, 1])
@pytest.mark.parametrize("nranks,h,w", [(2, 256, 256), (3, 480, 192)])
def test_loopback_halo_tiles_mid_pass(nranks, h, w, dynamic):
    """cg_halo_mid = 1: the tile rows mirrored into the neighbours sit in the middle of every pass and the halo flags go out right
    behind them (pano_cg_sr.cu: tile_at).  Tile order and flag timing change; the results only by the rounding of the regrouped
    partial sums."""
    from panopaea_b200 import dist
    k = 2
    prm = dict(timestep=0.05, threshold=0.1, max_iterations=60, inflow=(5 * k, 20 * k, 27 * k, 32 * k), inflow_density=1.0,
               inflow_vy=20.0, obstacle=(70 * k, 80 * k, 25 * k, 35 * k))
    out = []
    for mid in (0, 1):
        ranks = _make_ranks(nranks, h, w, prm)
        for r in ranks:
            r.ctx.set_option("cg_dynamic", dynamic)
            r.ctx.set_option("cg_halo_mid", mid)
        infos = None
        for _ in range(4):
            infos = _step_all(ranks)
        out.append((infos, [dist.gather_local(ranks, f) for f in (dist.DENSITY, dist.VY, dist.VX, dist.PRESSURE)]))
        for r in ranks:
            r.close()
    assert [i["iterations"] for i in out[0][0]] == [i["iterations"] for i in out[1][0]]
    for a, b in zip(out[0][1], out[1][1]):
        assert np.abs(a - b).max() <= 1e-9 * max(1.0, np.abs(a).max())


def test_step_host_two_ranks_threads():
    """pano_dist_step_host (host rows in, step, host rows out, synchronising): the call is collective, so the two loop-back ranks are
    driven from two threads (ctypes releases the GIL).  Compared with pano_fluid_step_host on one GPU, which it mirrors."""
    import ctypes as C
    import threading
    from tests import gpu_util as U
    from panopaea_b200 import _lib, dist, fluid
    h = w = 256
    k = 2
    prm = dict(timestep=0.05, threshold=0.1, max_iterations=100, inflow=(5 * k, 20 * k, 27 * k, 32 * k), inflow_density=1.0,
               inflow_vy=20.0, obstacle=(70 * k, 80 * k, 25 * k, 35 * k))
    L = _lib.load()
    single = fluid.DecFluid(h=h, w=w, ctx=U.ctx(), **prm)
    ranks = _make_ranks(2, h, w, prm)
    bufs = []
    for r in ranks:
        mine = []
        for which in (dist.DENSITY, dist.VY, dist.VX):
            n = r._rows(which) * r._pitch(which)
            p = C.c_void_p()
            _lib.check(L.pano_host_alloc(n * 8, C.byref(p)))
            a = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_double)), shape=(n,))
            a[:] = 0.0
            mine.append((p, a))
        bufs.append(mine)
    infos = [None, None]

    def run(i):
        infos[i] = ranks[i].step_host(*[p for p, _ in bufs[i]])

    for step in range(4):
        want = single.step()
        ts = [threading.Thread(target=run, args=(i,)) for i in range(2)]
        for t in ts:
            t.start()
        for t in ts:
            t.join(timeout=120)
        assert all(not t.is_alive() for t in ts)
        assert infos[0] == infos[1] and abs(infos[0]["iterations"] - want["iterations"]) <= 1, (step, infos, want)
        if infos[0]["iterations"] != want["iterations"]:
            break
        y0 = [r.y0 for r in ranks] + [h]
        d = np.concatenate([bufs[i][0][1].reshape(-1, w) for i in range(2)])
        vy = np.concatenate([bufs[i][1][1].reshape(-1, w) for i in range(2)])
        vx = np.concatenate([bufs[i][2][1].reshape(-1, w + 1) for i in range(2)])
        svy, svx = single.vel.split()
        assert d.shape == (h, w) and vy.shape == (h + 1, w) and vx.shape == (h, w + 1), y0
        assert np.abs(d - single.density.to_host()).max() <= 1e-9
        assert np.abs(vy - svy).max() <= 1e-8 * max(1.0, np.abs(svy).max()) and np.abs(vx - svx).max() <= 1e-8 * max(1.0, np.abs(svx).max())
    for mine in bufs:
        for p, _ in mine:
            L.pano_host_free(p)
