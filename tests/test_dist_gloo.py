"""CPU coverage of the N>1 path (world size 2 and 3, gloo): the slab partition and the halo / reduction
protocol of panopaea_b200/csrc/pano_dist.cu, emulated with the numpy restatement of the reference."""
import ctypes as C
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_slab_range_matches_library():
    from panopaea_b200 import _lib
    from panopaea_b200.dist import slab_range
    L = _lib.load()
    for h in (8, 97, 128, 1000, 8192):
        for n in (1, 2, 3, 4, 8):
            covered = 0
            for r in range(n):
                a, b = C.c_size_t(), C.c_size_t()
                assert L.pano_slab_range(h, r, n, C.byref(a), C.byref(b)) == 0
                assert (a.value, b.value) == slab_range(h, r, n)
                assert a.value == covered
                covered = b.value
            assert covered == h
    a, b = C.c_size_t(), C.c_size_t()
    assert L.pano_slab_range(10, 3, 2, C.byref(a), C.byref(b)) != 0      # rank out of range


@pytest.mark.parametrize("mode", ["exchanged", "fused"])
@pytest.mark.parametrize("world,h,w", [(2, 64, 40), (3, 75, 32)])
def test_decomposed_step_matches_global(tmp_path, world, h, w, mode):
    """exchanged: four ghost-row exchanges per step + two reductions per CG iteration (k_cg_stream path);
    fused: ONE exchange per step, local recomputation of the rest, single-reduction CG (the default path of pano_dist.cu)."""
    out = tmp_path / "result.txt"
    port = 29600 + world + (10 if mode == "fused" else 0)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "dist_gloo_worker.py"), str(out), str(h), str(w), "6", mode]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stderr[-3000:]
    text = out.read_text()
    assert text.startswith("PASS"), text
