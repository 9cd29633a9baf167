"""The per-cell expressions shared by every CUDA kernel (panopaea_b200/csrc/pano_cell_math.h),
compiled for the HOST and compared bit for bit with the oracle.  This checks the arithmetic
the GPU executes without needing a GPU; the kernels' indexing is covered by the -m gpu tests."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "hostcheck", "cell_math_host.cpp")
HDR = os.path.join(os.path.dirname(HERE), "panopaea_b200", "csrc", "pano_cell_math.h")
SO = os.path.join(HERE, "hostcheck", "libcell_math_host.so")


@pytest.fixture(scope="module")
def hc():
    if not os.path.exists(SO) or max(os.path.getmtime(SRC), os.path.getmtime(HDR)) > os.path.getmtime(SO):
        subprocess.run(["/usr/bin/g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", SRC, "-o", SO], check=True)
    return C.CDLL(SO)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


CASES = [(2, 2, 50.0), (3, 3, 30.0), (5, 5, 30.0), (17, 33, 30.0), (33, 17, 200.0), (64, 48, 200.0), (40, 40, 1e4), (9, 11, 1e11), (9, 11, 1e14),
         (12, 10, 1e25)]


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("h,w,vmax", CASES)
def test_advect_all(hc, oracle, dtype, h, w, vmax):
    rng = np.random.default_rng(11)
    q = rng.uniform(-1, 1, (h, w)).astype(dtype)
    vel = rng.uniform(-vmax, vmax, oracle.num_elem_1(h, w)).astype(dtype)
    qd, vd = np.zeros_like(q), np.zeros_like(vel)
    real = C.c_double if dtype == np.float64 else C.c_float
    getattr(hc, "hc_advect_all_" + ("f64" if dtype == np.float64 else "f32"))(h, w, _p(qd), _p(vd), _p(q), _p(vel), real(0.05))
    assert np.array_equal(qd, oracle.advect(h, w, q, 0.05, vel))
    assert np.array_equal(vd, oracle.advect_mac(h, w, vel, 0.05, vel))


@pytest.mark.parametrize("h,w,vmax", CASES + [(9, 11, 1e11), (9, 11, 1e14), (6, 5, 1e300), (31, 29, 1e-300)])
def test_advect_all_fast_forms(hc, oracle, h, w, vmax):
    """The exact fast forms of k_advect_march3 (one clamp per axis, floor by a magic add) give the oracle's bits,
    including backtraces far beyond the grid (guarded fall-back above 2^32 cells) and denormal-scale velocities."""
    rng = np.random.default_rng(13)
    q = rng.uniform(-1, 1, (h, w))
    vel = rng.uniform(-vmax, vmax, oracle.num_elem_1(h, w))
    vel[::7] = 0.0
    vel[3::11] *= 1e-3
    qd, vd = np.zeros_like(q), np.zeros_like(vel)
    hc.hc_advect_all_fast_f64(h, w, _p(qd), _p(vd), _p(q), _p(vel), C.c_double(0.05))
    assert np.array_equal(qd, oracle.advect(h, w, q, 0.05, vel))
    assert np.array_equal(vd, oracle.advect_mac(h, w, vel, 0.05, vel))


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("h,w", [(2, 2), (3, 3), (5, 7), (16, 16), (17, 33)])
def test_laplacian_and_divergence(hc, oracle, dtype, h, w):
    rng = np.random.default_rng(12)
    p = rng.uniform(-3, 3, (h, w)).astype(dtype)
    vel = rng.uniform(-5, 5, oracle.num_elem_1(h, w)).astype(dtype)
    for obstacle in [(0, 0, 0, 0), (h // 2, min(h, h // 2 + 2), w // 3, min(w, w // 3 + 3)), (0, 1, 0, 2), (h - 1, h, w - 2, w)]:
        sfx = "f64" if dtype == np.float64 else "f32"
        real = C.c_double if dtype == np.float64 else C.c_float
        z = np.zeros_like(p)
        getattr(hc, "hc_laplacian_" + sfx)(h, w, _p(z), _p(p), real(0.05), *obstacle)
        assert np.array_equal(z, oracle.laplacian_closure(h, w, p, 0.05, obstacle))
        b = np.zeros_like(p)
        getattr(hc, "hc_neg_divergence_" + sfx)(h, w, _p(b), _p(vel), *obstacle)
        e = oracle.hodge_1_dual(h, w, vel)
        ey, ex = oracle.split(e, h, w)
        y0, y1, x0, x1 = obstacle
        ey[y0:y1, x0:x1] = 0
        ex[y0:y1, x0:x1] = 0
        assert np.array_equal(b, -oracle.derivative_1_primal(h, w, e))


# ------------------------------------------------------------------ Grid3d forms against oracle/pano_oracle3.inc
CASES3 = [(2, 2, 2, 50.0), (3, 4, 5, 30.0), (9, 17, 12, 30.0), (16, 8, 33, 200.0), (6, 7, 5, 1e4), (5, 6, 7, 1e11), (4, 5, 6, 1e14), (3, 4, 5, 1e300),
          (7, 6, 5, 1e-300)]


@pytest.mark.parametrize("fast", [0, 1])
@pytest.mark.parametrize("d,h,w,vmax", CASES3)
def test_advect3_all(hc, fast, d, h, w, vmax):
    from oracle import pano_oracle3 as O3
    rng = np.random.default_rng(21)
    q = rng.uniform(-1, 1, (d, h, w))
    vel = rng.uniform(-vmax, vmax, O3.num_faces(d, h, w))
    vel[::7] = 0.0
    vel[3::11] *= 1e-3
    src = rng.uniform(-2, 2, vel.size)
    qd, vd = np.zeros_like(q), np.zeros_like(vel)
    hc.hc_advect3_all(fast, d, h, w, _p(qd), _p(vd), _p(q), _p(src), _p(vel), C.c_double(0.05))
    assert np.array_equal(qd, O3.advect(d, h, w, q, 0.05, vel))
    assert np.array_equal(vd, O3.advect_mac(d, h, w, src, 0.05, vel))


@pytest.mark.parametrize("d,h,w", [(2, 2, 2), (3, 4, 5), (8, 8, 8), (7, 12, 9)])
def test_laplacian3_and_divergence3(hc, d, h, w):
    from oracle import pano_oracle3 as O3
    rng = np.random.default_rng(22)
    p = rng.uniform(-3, 3, (d, h, w))
    vel = rng.uniform(-5, 5, O3.num_faces(d, h, w))
    for ob in [(0,) * 6, (d // 2, min(d, d // 2 + 2), h // 2, min(h, h // 2 + 2), w // 3, min(w, w // 3 + 3)), (0, 1, 0, 2, 0, 2),
               (d - 1, d, h - 2, h, w - 1, w), (0, d, 0, h, 0, w)]:
        out = np.zeros_like(p)
        hc.hc_laplacian3(d, h, w, _p(out), _p(p), C.c_double(0.05), *ob)
        assert np.array_equal(out, O3.laplacian_closure(d, h, w, p, 0.05, ob))
        b = np.zeros_like(p)
        hc.hc_neg_divergence3(d, h, w, _p(b), _p(vel), *ob)
        assert np.array_equal(b, O3.neg_divergence(d, h, w, vel, ob))
