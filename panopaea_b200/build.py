"""Builds panopaea_b200/lib/libpanopaea_b200.so from csrc/*.cu with nvcc for sm_100a.

In-tree, explicit nvcc (no JIT cache): the .so travels to the GPU box with the repo snapshot.
--fmad=false keeps a*b+c unfused, as rustc does for the reference, so element-wise results are
bit-identical to the CPU oracle.  cudart is linked statically, so the library loads (and its
symbols can be checked) on a machine without a CUDA driver.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
OBJDIR = os.path.join(HERE, "build")
SO = os.path.join(LIBDIR, "libpanopaea_b200.so")

NVCC = os.environ.get("PANO_NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
CFLAGS = ["-O3", "-std=c++17", "-lineinfo", "--fmad=false", "-ccbin", "/usr/bin/g++",
          "-Xcompiler", "-fPIC,-fvisibility=hidden,-Wall,-Wno-unused-function", "-Xptxas", "-v"]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _deps():
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(os.path.dirname(HERE), "include", "panopaea_b200.h"))
    hdrs.append(os.path.abspath(__file__))
    return hdrs


def _stamp(paths):
    m = hashlib.sha256()
    for p in sorted(paths):
        m.update(p.encode())
        with open(p, "rb") as f:
            m.update(f.read())
    m.update(" ".join(ARCH + CFLAGS).encode())
    return m.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(OBJDIR, exist_ok=True)
    srcs = sources()
    stamp_file = os.path.join(LIBDIR, "build.stamp")   # next to the .so: travels with it to the GPU box
    stamp = _stamp(srcs + _deps())
    if not force and os.path.exists(SO) and os.path.exists(stamp_file) and open(stamp_file).read() == stamp:
        return SO
    if not os.path.exists(NVCC):
        raise RuntimeError(f"nvcc not found at {NVCC}; cannot build the CUDA library")

    def compile_one(src):
        obj = os.path.join(OBJDIR, os.path.basename(src)[:-3] + ".o")
        cmd = [NVCC] + ARCH + CFLAGS + ["-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        log = os.path.join(OBJDIR, os.path.basename(src)[:-3] + ".ptxas.log")
        with open(log, "w") as f:
            f.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(compile_one, srcs))
    cmd = [NVCC] + ARCH + ["-shared", "-ccbin", "/usr/bin/g++", "-o", SO] + objs + ["-cudart", "static", "-lpthread", "-ldl"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(stamp_file, "w") as f:
        f.write(stamp)
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
