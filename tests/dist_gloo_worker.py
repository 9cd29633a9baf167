"""Worker of tests/test_dist_gloo.py: one rank of the slab-decomposed step, emulated on the CPU.

Launched by torchrun with the gloo backend (world size >= 2).  Every rank keeps GLOBAL-shaped arrays
that are NaN outside the rows it is entitled to see (its slab + the ghost rows of
panopaea_b200/csrc/pano_dist.cu), exchanges exactly the rows the CUDA path exchanges (8 ghost rows of
q / vy / vx before the advection, 1 row of the new vy, 1 row of b, 1 row of s per CG iteration, 1 row
of p; dot products and max-norms by all-reduce), and computes with the numpy restatement of the
reference.  A NaN in an owned row means the decomposition needed a row it does not exchange.
Rank 0 compares the gathered result with the undecomposed numpy run and writes PASS/FAIL to argv[1].
TEST INFRASTRUCTURE ONLY (uses oracle/)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import np_oracle as NP  # noqa: E402
from panopaea_b200.dist import slab_range  # noqa: E402

GHOST = 8
# "fused" mode = the fused-halo step of pano_dist.cu (the default with the single-reduction solver): ONE exchange of GHOST_FUSED
# ghost rows per step; every rank advects EXTEND rows beyond its slab itself, computes b two rows beyond it, and the solver
# (Chronopoulos-Gear form, pano_cg_sr.cu) exchanges two rows of r and one of s per pass, ONE all-reduce of three values per pass,
# and mirrors the last row of x into the lower neighbour.
GHOST_FUSED, EXTEND = 12, 4


def main():
    out_path, H, W, steps = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    if len(sys.argv) > 5 and sys.argv[5] == "fused":
        return main_fused(out_path, H, W, steps)
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    y0, y1 = slab_range(H, rank, world)
    last = rank == world - 1
    dt, thr, max_it = 0.05, 0.1, 100
    inflow, obstacle = (3, 9, 20, 26), (H // 2 + 1, H // 2 + 5, 10, 22)

    def nan(shape):
        return np.full(shape, np.nan)

    def exchange(a, nrows):
        """my first `nrows` owned rows -> upper neighbour's ghost rows below its slab, my last -> lower neighbour."""
        reqs, bufs = [], {}
        if rank > 0:
            reqs.append(dist.isend(torch.from_numpy(np.ascontiguousarray(a[y0:y0 + nrows])), rank - 1))
            bufs["up"] = torch.empty((nrows, a.shape[1]), dtype=torch.float64)
            reqs.append(dist.irecv(bufs["up"], rank - 1))
        if not last:
            reqs.append(dist.isend(torch.from_numpy(np.ascontiguousarray(a[y1 - nrows:y1])), rank + 1))
            bufs["dn"] = torch.empty((nrows, a.shape[1]), dtype=torch.float64)
            reqs.append(dist.irecv(bufs["dn"], rank + 1))
        for r in reqs:
            r.wait()
        if "up" in bufs:
            a[y0 - nrows:y0] = bufs["up"].numpy()        # the upper neighbour's last rows
        if "dn" in bufs:
            a[y1:y1 + nrows] = bufs["dn"].numpy()        # the lower neighbour's first rows

    def allreduce(v, op):
        t = torch.tensor([v], dtype=torch.float64)
        dist.all_reduce(t, op=op)
        return float(t.item())

    yf = y1 + 1 if last else y1                          # end of my vy face rows
    density, vy, vx, pressure = nan((H, W)), nan((H + 1, W)), nan((H, W + 1)), nan((H, W))
    density[y0:y1], vy[y0:yf], vx[y0:y1], pressure[y0:y1] = 0.0, 0.0, 0.0, 0.0
    iters = []
    for _ in range(steps):
        # inflow on the rows I own
        a, b_ = max(inflow[0], y0), min(inflow[1], y1)
        if b_ > a:
            density[a:b_, inflow[2]:inflow[3]] = 1.0
            vy[a:b_, inflow[2]:inflow[3]] = 20.0
        for arr in (density, vy, vx):
            exchange(arr, GHOST)
        vy_c, vx_c = np.nan_to_num(vy), np.nan_to_num(vx)          # velocities only steer the backtrace
        d_new = NP.advect(density, dt, vy_c, vx_c)
        vy_new, vx_new = NP.advect_mac(vy, vx, dt, vy_c, vx_c)
        density, vy, vx = nan((H, W)), nan((H + 1, W)), nan((H, W + 1))
        density[y0:y1], vy[y0:yf], vx[y0:y1] = d_new[y0:y1], vy_new[y0:yf], vx_new[y0:y1]
        exchange(vy, 1)
        b = nan((H, W))
        b[y0:y1] = NP.neg_divergence(vy, vx, obstacle)[y0:y1]
        # CG (pcg.rs:14-82) with one ghost row of the search direction per iteration
        x = np.zeros((y1 - y0, W))
        bmax = allreduce(np.abs(b[y0:y1]).max(), dist.ReduceOp.MAX)
        it = -1
        if bmax >= thr:
            r = b[y0:y1].copy()
            s = nan((H, W))
            s[y0:y1] = r
            sigma = allreduce(float((r * r).sum()), dist.ReduceOp.SUM)
            it = max_it
            for i in range(max_it):
                exchange(s, 1)
                z = NP.laplacian(s, dt, obstacle)[y0:y1]
                alpha = sigma / allreduce(float((z * s[y0:y1]).sum()), dist.ReduceOp.SUM)
                x = x + alpha * s[y0:y1]
                r = r + (-alpha) * z
                if allreduce(np.abs(r).max(), dist.ReduceOp.MAX) < thr:
                    it = i
                    break
                sigma_new = allreduce(float((r * r).sum()), dist.ReduceOp.SUM)
                s[y0:y1] = r + (sigma_new / sigma) * s[y0:y1]
                sigma = sigma_new
        iters.append(it)
        pressure = nan((H, W))
        pressure[y0:y1] = x
        exchange(pressure, 1)
        gy, gx = np.zeros((H + 1, W)), np.zeros((H, W + 1))
        NP.derivative_0_dual(pressure, gy, gx)
        vy[y0:yf] = vy[y0:yf] + dt * gy[y0:yf]
        vx[y0:y1] = vx[y0:y1] + dt * gx[y0:y1]
        vx[y0:y1, 0] = 0.0
        vx[y0:y1, -1] = 0.0
        if rank == 0:
            vy[0] = 0.0
        if last:
            vy[H] = 0.0
    parts = [None] * world
    dist.all_gather_object(parts, (density[y0:y1], vy[y0:yf], vx[y0:y1], pressure[y0:y1], iters))
    ok = True
    if rank == 0:
        ref = NP.FluidState(H, W, inflow=inflow, obstacle=obstacle)
        ref_iters = [ref.step()["iterations"] for _ in range(steps)]
        got = [np.concatenate([p[k] for p in parts]) for k in range(4)]
        want = [ref.density, ref.vy, ref.vx, ref.pressure]
        msgs = []
        for name, g, w_ in zip(("density", "vy", "vx", "pressure"), got, want):
            bad = (not np.all(np.isfinite(g))) or np.abs(g - w_).max() > 1e-9 * max(1.0, np.abs(w_).max())
            msgs.append(f"{name}: finite={bool(np.all(np.isfinite(g)))} err={np.nanmax(np.abs(g - w_)):.2e}")
            ok &= not bad
        ok &= all(p[4] == parts[0][4] for p in parts) and all(abs(a_ - b2) <= 1 for a_, b2 in zip(parts[0][4], ref_iters))
        with open(out_path, "w") as f:
            f.write(("PASS" if ok else "FAIL") + "\n" + "\n".join(msgs) + f"\niters {parts[0][4]} ref {ref_iters}\n")
    dist.barrier()
    dist.destroy_process_group()


def main_fused(out_path, H, W, steps):
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    y0, y1 = slab_range(H, rank, world)
    last = rank == world - 1
    dt, thr, max_it = 0.05, 0.1, 100
    inflow, obstacle = (3, 9, 20, 26), (H // 2 + 1, H // 2 + 5, 10, 22)

    def nan(shape):
        return np.full(shape, np.nan)

    def exchange(a, nrows, up=True, down=True):
        """my first `nrows` owned rows -> upper neighbour (if up), my last `nrows` -> lower neighbour (if down)"""
        reqs, bufs = [], {}
        if rank > 0:
            if up:
                reqs.append(dist.isend(torch.from_numpy(np.ascontiguousarray(a[y0:y0 + nrows])), rank - 1))
            if down:
                bufs["up"] = torch.empty((nrows, a.shape[1]), dtype=torch.float64)
                reqs.append(dist.irecv(bufs["up"], rank - 1))
        if not last:
            if down:
                reqs.append(dist.isend(torch.from_numpy(np.ascontiguousarray(a[y1 - nrows:y1])), rank + 1))
            if up:
                bufs["dn"] = torch.empty((nrows, a.shape[1]), dtype=torch.float64)
                reqs.append(dist.irecv(bufs["dn"], rank + 1))
        for r in reqs:
            r.wait()
        if "up" in bufs:
            a[y0 - nrows:y0] = bufs["up"].numpy()
        if "dn" in bufs:
            a[y1:y1 + nrows] = bufs["dn"].numpy()

    def allreduce3(g, d_, m):
        t = torch.tensor([g, d_], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        u = torch.tensor([m], dtype=torch.float64)
        dist.all_reduce(u, op=dist.ReduceOp.MAX)
        return float(t[0]), float(t[1]), float(u[0])

    def keep(a, lo, hi):
        """a copy of `a` that is NaN outside rows [lo, hi): what this rank may know of a freshly computed array"""
        out = nan(a.shape)
        lo, hi = max(lo, 0), min(hi, a.shape[0])
        out[lo:hi] = a[lo:hi]
        return out

    yf = y1 + 1 if last else y1
    density, vy, vx, pressure = nan((H, W)), nan((H + 1, W)), nan((H, W + 1)), nan((H, W))
    density[y0:y1], vy[y0:yf], vx[y0:y1], pressure[y0:y1] = 0.0, 0.0, 0.0, 0.0
    iters = []
    for _ in range(steps):
        a, b_ = max(inflow[0], y0), min(inflow[1], y1)
        if b_ > a:
            density[a:b_, inflow[2]:inflow[3]] = 1.0
            vy[a:b_, inflow[2]:inflow[3]] = 20.0
        for arr in (density, vy, vx):
            exchange(arr, GHOST_FUSED)                               # THE exchange of the step
        vy_c, vx_c = np.nan_to_num(vy), np.nan_to_num(vx)
        A, B = max(0, y0 - EXTEND), min(H, y1 + EXTEND)              # cell rows advected here
        Bf = B + 1 if B == H else B                                  # ... and their vy face rows
        d_new = NP.advect(density, dt, vy_c, vx_c)
        vy_new, vx_new = NP.advect_mac(vy, vx, dt, vy_c, vx_c)
        density, vy, vx = keep(d_new, A, B), keep(vy_new, A, Bf), keep(vx_new, A, B)
        A2, B2 = max(0, y0 - 2), min(H, y1 + 2)
        b = keep(NP.neg_divergence(vy, vx, obstacle), A2, B2)
        assert np.all(np.isfinite(b[A2:B2])), "b needs a row this rank did not advect"
        # ---- single-reduction CG: r with two ghost rows, s with one
        lap = lambda v: NP.laplacian(v, dt, obstacle)
        x = np.zeros((H, W))
        g0, d0, bmax = allreduce3(float((b[y0:y1] ** 2).sum()), float((lap(b)[y0:y1] * b[y0:y1]).sum()), float(np.abs(b[y0:y1]).max()))
        it = -1
        if bmax >= thr:
            r = b.copy()                                             # rows [y0-2, y1+2) known
            s_old, p_old = np.zeros((H, W)), np.zeros((H, W))
            gamma, alpha, beta = g0, g0 / d0, 0.0
            it = max_it
            for i in range(max_it):
                lo1, hi1 = max(0, y0 - 1), min(H, y1 + 1)            # tile + one-cell ring
                w_i = lap(r)
                s_new = keep(w_i + beta * s_old, lo1, hi1)
                r_new = keep(r + (-alpha) * s_new, lo1, hi1)
                assert np.all(np.isfinite(r_new[lo1:hi1])), "the ring recomputation read a row that was not exchanged"
                p_new = r[y0:y1] + beta * p_old[y0:y1]
                x[y0:y1] = x[y0:y1] + alpha * p_new
                p_old = np.zeros((H, W))
                p_old[y0:y1] = p_new
                w_next = lap(r_new)[y0:y1]
                assert np.all(np.isfinite(w_next))
                gn, dn_, rmax = allreduce3(float((r_new[y0:y1] ** 2).sum()), float((w_next * r_new[y0:y1]).sum()), float(np.abs(r_new[y0:y1]).max()))
                # halo stores of the pass: two rows of r, one of s (the kernel writes them into the neighbours' ghost rows)
                r = keep(r_new, y0, y1)
                s_old = keep(s_new, y0, y1)
                exchange(r, 2)
                exchange(s_old, 1)
                if rmax < thr:
                    it = i
                    break
                beta = gn / gamma
                alpha = gn / (dn_ - beta * gn / alpha)
                gamma = gn
        iters.append(it)
        pressure = nan((H, W))
        pressure[y0:y1] = x[y0:y1]
        exchange(pressure, 1, up=False, down=True)                   # my last row -> the lower neighbour's ghost row (in-kernel mirror)
        gy, gx = np.zeros((H + 1, W)), np.zeros((H, W + 1))
        NP.derivative_0_dual(pressure, gy, gx)
        vy[y0:yf] = vy[y0:yf] + dt * gy[y0:yf]
        vx[y0:y1] = vx[y0:y1] + dt * gx[y0:y1]
        vx[y0:y1, 0] = 0.0
        vx[y0:y1, -1] = 0.0
        if rank == 0:
            vy[0] = 0.0
        if last:
            vy[H] = 0.0
        # rows beyond the slab were scratch of this step
        density, vy, vx = keep(density, y0, y1), keep(vy, y0, yf), keep(vx, y0, y1)
    parts = [None] * world
    dist.all_gather_object(parts, (density[y0:y1], vy[y0:yf], vx[y0:y1], pressure[y0:y1], iters))
    ok = True
    if rank == 0:
        ref = NP.FluidState(H, W, inflow=inflow, obstacle=obstacle)
        ref_iters = [ref.step()["iterations"] for _ in range(steps)]
        got = [np.concatenate([p[k] for p in parts]) for k in range(4)]
        want = [ref.density, ref.vy, ref.vx, ref.pressure]
        msgs = []
        for name, g, w_ in zip(("density", "vy", "vx", "pressure"), got, want):
            bad = (not np.all(np.isfinite(g))) or np.abs(g - w_).max() > 1e-9 * max(1.0, np.abs(w_).max())
            msgs.append(f"{name}: finite={bool(np.all(np.isfinite(g)))} err={np.nanmax(np.abs(g - w_)):.2e}")
            ok &= not bad
        ok &= all(p[4] == parts[0][4] for p in parts) and all(abs(a_ - b2) <= 1 for a_, b2 in zip(parts[0][4], ref_iters))
        with open(out_path, "w") as f:
            f.write(("PASS" if ok else "FAIL") + "\n" + "\n".join(msgs) + f"\niters {parts[0][4]} ref {ref_iters}\n")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
