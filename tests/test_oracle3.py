"""Grid3d checker (oracle/pano_oracle3.inc) against the independent numpy statement (oracle/np_oracle3.py).

The reference has no 3-D fluid code: `trilinear` (panopaea/src/math/interp.rs:23-36) is the one pinned item; everything else
is the specification of DESIGN.md 5c, and these two statements of it must agree bit for bit on the element-wise passes."""
import numpy as np
import pytest

from oracle import np_oracle3 as NP3
from oracle import pano_oracle3 as O3

CASES = [(2, 2, 2, 30.0), (3, 4, 5, 30.0), (5, 3, 2, 200.0), (9, 17, 12, 30.0), (16, 8, 33, 200.0)]


def _inputs(d, h, w, vmax, seed):
    rng = np.random.default_rng(seed)
    return rng.uniform(-1, 1, (d, h, w)), rng.uniform(-vmax, vmax, O3.num_faces(d, h, w))


def test_trilinear_is_the_reference_expression():
    rng = np.random.default_rng(0)
    for _ in range(50):
        a = rng.uniform(-3, 3, 8)
        s, t, u = rng.uniform(0, 1, 3)
        lin = lambda a0, a1, k: a0 * (1.0 - k) + a1 * k                                      # interp.rs:7-12
        bil = lambda a00, a01, a10, a11: lin(lin(a00, a01, s), lin(a10, a11, s), t)             # interp.rs:15-20
        want = lin(bil(a[0], a[1], a[2], a[3]), bil(a[4], a[5], a[6], a[7]), u)                 # interp.rs:31-35
        assert O3.trilinear(*a, s, t, u) == want
    assert O3.trilinear(1, 1, 1, 1, 1, 1, 1, 1, 0.3, 0.6, 0.9) == 1.0
    assert O3.trilinear(0, 1, 0, 1, 0, 1, 0, 1, 0.25, 0.5, 0.75) == 0.25                        # s runs along x
    assert O3.trilinear(0, 0, 1, 1, 0, 0, 1, 1, 0.25, 0.5, 0.75) == 0.5                         # t along y
    assert O3.trilinear(0, 0, 0, 0, 1, 1, 1, 1, 0.25, 0.5, 0.75) == 0.75                        # u along z


@pytest.mark.parametrize("d,h,w,vmax", CASES)
def test_advect_bit_exact(d, h, w, vmax):
    q, vel = _inputs(d, h, w, vmax, 3)
    vz, vy, vx = O3.split(vel, d, h, w)
    assert np.array_equal(O3.advect(d, h, w, q, 0.05, vel), NP3.advect(q, 0.05, vz, vy, vx))


@pytest.mark.parametrize("d,h,w,vmax", CASES)
def test_advect_mac_bit_exact(d, h, w, vmax):
    _, vel = _inputs(d, h, w, vmax, 4)
    _, src = _inputs(d, h, w, 1.0, 5)
    vz, vy, vx = O3.split(vel, d, h, w)
    qz, qy, qx = O3.split(src, d, h, w)
    got = O3.split(O3.advect_mac(d, h, w, src, 0.05, vel), d, h, w)
    want = NP3.advect_mac(qz, qy, qx, 0.05, vz, vy, vx)
    for g, w_ in zip(got, want):
        assert np.array_equal(g, w_)


def test_advect_reduces_to_2d_on_a_z_uniform_field(oracle):
    """Every z plane equal and vz = 0: each plane must reproduce the 2-D reference functions (dec_fluid.rs:173-291) bit for bit."""
    d, h, w = 4, 9, 11
    rng = np.random.default_rng(7)
    q2, vel2 = rng.uniform(-1, 1, (h, w)), rng.uniform(-40, 40, (h + 1) * w + h * (w + 1))
    vy2, vx2 = oracle.split(vel2, h, w)
    q3 = np.broadcast_to(q2, (d, h, w)).copy()
    vel3 = O3.join(np.zeros((d + 1, h, w)), np.broadcast_to(vy2, (d, h + 1, w)), np.broadcast_to(vx2, (d, h, w + 1)))
    a3 = O3.advect(d, h, w, q3, 0.05, vel3)
    m3 = O3.split(O3.advect_mac(d, h, w, vel3, 0.05, vel3), d, h, w)
    a2 = oracle.advect(h, w, q2, 0.05, vel2)
    my2, mx2 = oracle.split(oracle.advect_mac(h, w, vel2, 0.05, vel2), h, w)
    for z in range(d):
        assert np.array_equal(a3[z], a2) and np.array_equal(m3[1][z], my2) and np.array_equal(m3[2][z], mx2)
    assert not m3[0].any()


@pytest.mark.parametrize("d,h,w", [(2, 2, 2), (3, 4, 5), (8, 8, 8), (7, 12, 9)])
def test_divergence_laplacian_projection(d, h, w):
    q, vel = _inputs(d, h, w, 5.0, 6)
    ob = (d // 2, min(d, d // 2 + 2), h // 2, min(h, h // 2 + 2), w // 3, min(w, w // 3 + 3))
    vz, vy, vx = O3.split(vel, d, h, w)
    assert np.array_equal(O3.neg_divergence(d, h, w, vel, ob), NP3.neg_divergence(vz, vy, vx, ob))
    assert np.array_equal(O3.laplacian_closure(d, h, w, q, 0.05, ob), NP3.laplacian(q, 0.05, ob))
    got = O3.split(O3.project(d, h, w, vel, q, 0.05), d, h, w)
    for g, w_ in zip(got, NP3.project(vz, vy, vx, q, 0.05)):
        assert np.array_equal(g, w_)


def test_laplacian_reduces_to_the_reference_2d_closure(oracle):
    """A pressure that does not depend on z: the 7-point closure equals the reference's 5-point closure (dec_fluid.rs:100-119,
    pinned by the reference's own Laplacian vector in tests/test_oracle_golden.py)."""
    d, h, w = 3, 6, 7
    p2 = np.random.default_rng(1).uniform(-1, 1, (h, w))
    p3 = np.broadcast_to(p2, (d, h, w)).copy()
    ob2 = (2, 4, 1, 5)
    got = O3.laplacian_closure(d, h, w, p3, 0.05, (0, d, *ob2))
    want = oracle.laplacian_closure(h, w, p2, 0.05, ob2)
    for z in range(d):
        assert np.array_equal(got[z], want)


def test_laplacian_is_symmetric_psd():
    d, h, w = 4, 5, 6
    ob = (1, 3, 2, 4, 1, 5)
    n = d * h * w
    A = np.zeros((n, n))
    for k in range(n):
        e = np.zeros(n)
        e[k] = 1.0
        A[:, k] = O3.laplacian_closure(d, h, w, e.reshape(d, h, w), 1.0, ob).ravel()
    assert np.array_equal(A, A.T)
    assert np.all(np.abs(A.sum(axis=1)) == 0)
    assert np.linalg.eigvalsh(A).min() > -1e-12
    assert set(np.unique(np.diag(A))) <= {0.0, 1.0, 2.0, 3.0, 4.0, 5.0, 6.0}


def test_pcg_and_step_agree():
    d, h, w = 12, 16, 10
    rng = np.random.default_rng(2)
    ob = (3, 6, 8, 10, 2, 7)
    b = NP3.laplacian(rng.uniform(-1, 1, (d, h, w)), 0.05, ob)     # in the range of the (singular) operator
    r = O3.pcg(d, h, w, b, 100, 1e-6, 0.05, ob)
    x, it, err = NP3.pcg(b, 100, 1e-6, lambda s: NP3.laplacian(s, 0.05, ob))
    assert abs(r.iterations - it) <= 1 and r.iterations < 100
    assert np.abs(r.x - x).max() <= 1e-8 * np.abs(x).max()
    # the whole step, three passes of a small plume
    p = O3.smoke_params(32)
    st = O3.FluidState3(**p)
    ns = dict(density=np.zeros((32,) * 3), vz=np.zeros((33, 32, 32)), vy=np.zeros((32, 33, 32)), vx=np.zeros((32, 32, 33)), pressure=None)
    for i in range(3):
        o, n = st.step(want_rhs=True), NP3.step(ns, p)
        assert np.array_equal(st.field("density"), ns["density"]), i       # advection of step i: same inputs (1e-9-close) ...
        assert abs(o["iterations"] - n["iterations"]) <= 1
        vz, vy, vx = O3.split(st.field("vel"), 32, 32, 32)
        for g, w_ in ((vz, ns["vz"]), (vy, ns["vy"]), (vx, ns["vx"])):
            assert np.abs(g - w_).max() <= 1e-6 * max(1.0, np.abs(w_).max())
        # ... so re-synchronise the numpy state on the C state (dot-product order differs)
        ns.update(density=st.field("density").copy(), vz=vz.copy(), vy=vy.copy(), vx=vx.copy())
    assert st.field("density").max() > 0.5


def test_random_shapes_property():
    """Random small shapes, velocity scales and obstacle boxes: the C checker and the numpy statement agree bit for bit on every
    element-wise pass (a cheap fuzz of the index arithmetic of both)."""
    rng = np.random.default_rng(2024)
    for trial in range(25):
        d, h, w = (int(v) for v in rng.integers(2, 12, 3))
        vmax = float(rng.choice([0.5, 30.0, 300.0, 1e6]))
        q = rng.uniform(-1, 1, (d, h, w))
        vel = rng.uniform(-vmax, vmax, O3.num_faces(d, h, w))
        src = rng.uniform(-2, 2, vel.size)
        z0, y0, x0 = (int(rng.integers(0, n)) for n in (d, h, w))
        ob = (z0, int(rng.integers(z0, d + 1)), y0, int(rng.integers(y0, h + 1)), x0, int(rng.integers(x0, w + 1)))
        vz, vy, vx = O3.split(vel, d, h, w)
        qz, qy, qx = O3.split(src, d, h, w)
        assert np.array_equal(O3.advect(d, h, w, q, 0.05, vel), NP3.advect(q, 0.05, vz, vy, vx)), (trial, d, h, w)
        for g, want in zip(O3.split(O3.advect_mac(d, h, w, src, 0.05, vel), d, h, w), NP3.advect_mac(qz, qy, qx, 0.05, vz, vy, vx)):
            assert np.array_equal(g, want), (trial, d, h, w)
        assert np.array_equal(O3.neg_divergence(d, h, w, vel, ob), NP3.neg_divergence(vz, vy, vx, ob)), (trial, ob)
        assert np.array_equal(O3.laplacian_closure(d, h, w, q, 0.05, ob), NP3.laplacian(q, 0.05, ob)), (trial, ob)
