"""-m gpu: examples/dec_fluid.rs re-typed in C++ over the C ABI (panopaea_b200/host/dec_fluid.cpp), both with the
reference's own call sequence ("composed") and with the fused step, against the oracle."""
import os
import re
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "panopaea_b200", "host")


@pytest.mark.parametrize("mode", ["composed", "fused"])
def test_dec_fluid_cpp(oracle, tmp_path, mode):
    subprocess.run(["make", "-C", HOST, "-s"], check=True)
    steps = 30
    out = tmp_path / "state.bin"
    r = subprocess.run([os.path.join(HOST, "dec_fluid"), str(steps), mode, str(out)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    its = [int(m) for m in re.findall(r"Iterations (-?\d+)", r.stdout)]
    assert len(its) == steps
    ref = oracle.FluidState(**oracle.smoke_params(128))
    want = [ref.step()["iterations"] for _ in range(steps)]
    assert all(abs(a - b) <= 2 for a, b in zip(its, want))
    if its == want:
        raw = np.fromfile(out, dtype=np.float64)
        n2, n1 = 128 * 128, 129 * 128 + 128 * 129
        d, v, p = raw[:n2], raw[n2:n2 + n1], raw[n2 + n1:]
        for got, ref_f in ((d, ref.field("density").ravel()), (v, ref.field("vel")), (p, ref.field("pressure").ravel())):
            assert np.abs(got - ref_f).max() <= 1e-5 * max(1e-300, np.abs(ref_f).max())


def test_dec_fluid_cpp_multigrid_preconditioner():
    """The same example with `pcg::Multigrid` as the Preconditioner object of the generic loop (trait seam pcg.rs:4-6):
    every solve reaches the threshold within a few iterations."""
    subprocess.run(["make", "-C", HOST, "-s"], check=True)
    r = subprocess.run([os.path.join(HOST, "dec_fluid"), "30", "multigrid"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    its = [int(m) for m in re.findall(r"Iterations (-?\d+)", r.stdout)]
    assert len(its) == 30 and all(-1 <= i <= 4 for i in its), its


def test_dec_fluid_cpp_grid3(tmp_path):
    """The C++ mirror of the Grid3d addition (host/panopaea.hpp: domain::Grid3d, DecFluid3) against its checker."""
    from oracle import pano_oracle3 as O3
    subprocess.run(["make", "-C", HOST, "-s"], check=True)
    steps = 6
    out = tmp_path / "state3.bin"
    r = subprocess.run([os.path.join(HOST, "dec_fluid"), str(steps), "grid3", str(out)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    its = [int(m) for m in re.findall(r"Iterations (-?\d+)", r.stdout)]
    ref = O3.FluidState3(64, 64, 64, inflow=(27, 32, 2, 10, 27, 32), obstacle=(25, 35, 35, 40, 25, 35))
    want = [ref.step()["iterations"] for _ in range(steps)]
    assert len(its) == steps and all(abs(a - b) <= 2 for a, b in zip(its, want)), (its, want)
    if its == want:
        raw = np.fromfile(out, dtype=np.float64)
        nc, nf = 64 ** 3, O3.num_faces(64, 64, 64)
        for got, ref_f in ((raw[:nc], ref.field("density").ravel()), (raw[nc:nc + nf], ref.field("vel")), (raw[nc + nf:], ref.field("pressure").ravel())):
            assert np.abs(got - ref_f).max() <= 1e-5 * max(1e-300, np.abs(ref_f).max())
