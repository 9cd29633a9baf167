"""Regenerates tests/golden/oracle_vectors.npz.

The Rust reference cannot run in this image, so these vectors come from the C
oracle (oracle/pano_oracle.c) after it has been pinned against the reference's
own golden vectors and cross-checked against oracle/np_oracle.py.  They guard
the oracle against silent edits and give the GPU tests size-independent
anchors.  Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import pano_oracle as O  # noqa: E402


def inputs(h, w, vmax, seed=0):
    rng = np.random.default_rng(seed)
    q = rng.uniform(-1.0, 1.0, (h, w))
    vel = rng.uniform(-vmax, vmax, O.num_elem_1(h, w))
    return q, vel


def main():
    O.set_threading(O.SERIAL)
    out = {}
    for (h, w, vmax) in [(5, 5, 30.0), (3, 3, 30.0), (17, 33, 30.0), (17, 33, 200.0), (33, 17, 200.0)]:
        q, vel = inputs(h, w, vmax)
        tag = f"{h}x{w}_v{int(vmax)}"
        out[f"advect_{tag}"] = O.advect(h, w, q, 0.05, vel)
        out[f"advect_mac_{tag}"] = O.advect_mac(h, w, vel, 0.05, vel)
        obstacle = (h // 2, min(h, h // 2 + 2), w // 3, min(w, w // 3 + 3))
        out[f"lap_{tag}"] = O.laplacian_closure(h, w, q, 0.05, obstacle)
    # shipped example: 128^2, 25 steps; keep iteration counts and field checksums + one row of each field
    S = O.FluidState(**O.smoke_params(128))
    its, res = [], []
    for _ in range(25):
        r = S.step()
        its.append(r["iterations"])
        res.append(r["final_residual"])
    out["dec_fluid_128_iterations"] = np.array(its)
    out["dec_fluid_128_final_residual"] = np.array(res)
    out["dec_fluid_128_density_rows"] = S.field("density")[::16].copy()
    out["dec_fluid_128_pressure_rows"] = S.field("pressure")[::16].copy()
    vy, vx = O.split(S.field("vel"), 128, 128)
    out["dec_fluid_128_vy_rows"] = vy[::16].copy()
    out["dec_fluid_128_vx_rows"] = vx[::16].copy()
    out["dec_fluid_128_sums"] = np.array([S.field("density").sum(), np.abs(S.field("vel")).sum(),
                                          np.abs(S.field("pressure")).sum()])
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_vectors.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
