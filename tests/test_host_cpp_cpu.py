"""The C++ mirror of the crate interface (panopaea_b200/host) builds against the C ABI and, without a
GPU, fails loudly (no CPU fallback)."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "panopaea_b200", "host")


def test_host_mirror_builds_and_refuses_without_gpu():
    from panopaea_b200 import build as pb
    pb.build()
    subprocess.run(["make", "-C", HOST, "-s"], check=True)
    r = subprocess.run([os.path.join(HOST, "dec_fluid"), "1"], capture_output=True, text=True)
    import ctypes as C
    from panopaea_b200 import _lib
    n = C.c_int()
    _lib.load().pano_device_count(C.byref(n))
    if n.value == 0:
        assert r.returncode == 1 and "no CPU fallback" in r.stderr
    else:
        assert r.returncode == 0 and "Iterations" in r.stdout
