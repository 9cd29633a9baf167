#!/usr/bin/env python
"""Per-iteration section profile of the streaming CG kernel on an H x W grid, on one GPU or slab-decomposed over the
ranks of a torchrun job (option cg_profile: clock64 totals of CTA 0, per-CTA tile-loop totals).  Used to separate the
fixed per-phase costs of a small slab from what the cross-GPU exchange adds.
usage: [torchrun --nproc-per-node N] scripts/prof_slab.py H W [key=value ...]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import panopaea_b200 as P
from panopaea_b200 import _lib, dist, fluid

rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
h, w = int(sys.argv[1]), int(sys.argv[2])
opts = dict(kv.split("=") for kv in sys.argv[3:])
ky, kx = h // 128, w // 128
prm = dict(timestep=0.05, threshold=0.1, max_iterations=100, inflow=(5 * ky, 20 * ky, 54 * kx, 64 * kx), inflow_density=1.0,
           inflow_vy=20.0, obstacle=(70 * ky, 80 * ky, 50 * kx, 70 * kx))
ctx = P.Context(local)
for k, v in opts.items():
    ctx.set_option(k, int(v))
L = _lib.load()
if world > 1:
    import torch
    import torch.distributed as dist_t
    torch.cuda.set_device(local)
    dist_t.init_process_group("nccl", device_id=torch.device("cuda", local))
    D = dist.DistFluid(ctx, h, w, rank, world, prm)
    handles = [None] * world
    dist_t.all_gather_object(handles, D.ipc_handle())
    D.connect_ipc(handles)
    dist_t.barrier()
    step = D.step
    sync = lambda: (D.sync(), dist_t.barrier())
    info = lambda: (D.step(), D.sync())[1]
else:
    ctx.set_option("cg_kernel", int(opts.get("cg_kernel", 6 if int(opts.get("cg_single_reduction", 1)) else 2)))
    sim = fluid.DecFluid(h=h, w=w, ctx=ctx, **prm)
    step = lambda: sim.step(want_info=False)
    sync = ctx.sync
    info = sim.step

for _ in range(8):
    step()
sync()
K = 10
ctx.timer_start()
for _ in range(K):
    step()
ms = ctx.timer_stop_ms() / K
sync()
ctx.set_option("cg_profile", 1)
inf = info()
out = (C.c_int64 * 8)()
_lib.check(L.pano_ctx_cg_profile(ctx.handle, out))
G = ctx.num_sms()
arr = (C.c_int64 * (2 * G))()
_lib.check(L.pano_ctx_cg_profile_ctas(ctx.handle, arr, 2 * G))
ctx.set_option("cg_profile", 0)
its = max(1, inf["applies"])
a = np.array(arr[:], dtype=np.float64).reshape(2, G) / its / 1965.0
if int(opts.get("cg_single_reduction", 1)) != 0 and (world > 1 or int(opts.get("cg_kernel", 6)) == 6):
    # k_cg_sr: 0 wait for the pass's first tile, 1 tile loop, 2 batch-unit poll, 3 CTA reduction, 4 grid (+ cross-GPU) all-reduce
    us = [v / its / 1965.0 for v in out[:6]]
    line = (f"rank {rank}/{world} grid {h}x{w} opts {opts}: step {ms:.3f} ms; passes {its}+1; per pass (CTA 0): first-tile wait {us[0]:.1f} us, "
            f"tile loop {us[1]:.1f}, unit poll {us[2]:.1f}, CTA reduce {us[3]:.1f}, thread-0 fence {us[5]:.1f}, all-reduce {us[4]:.1f}, sum {sum(us):.1f} us; "
            f"over CTAs: first-tile wait min/med/max {a[0].min():.1f}/{np.median(a[0]):.1f}/{a[0].max():.1f}, "
            f"tile loop {a[1].min():.1f}/{np.median(a[1]):.1f}/{a[1].max():.1f}")
else:
    us = [v / its / 1965.0 for v in out[:4]]
    line = (f"rank {rank}/{world} grid {h}x{w} opts {opts}: step {ms:.3f} ms; applies {its}; per iteration (CTA 0): P1 tiles {us[0]:.1f} us, "
            f"reduce1 {us[1]:.1f}, P2 tiles {us[2]:.1f}, reduce2 {us[3]:.1f}, sum {sum(us):.1f} us; "
            f"tile loops over CTAs: P1 min/med/max {a[0].min():.1f}/{np.median(a[0]):.1f}/{a[0].max():.1f}, "
            f"P2 {a[1].min():.1f}/{np.median(a[1]):.1f}/{a[1].max():.1f}")
if world > 1:
    lines = [None] * world
    dist_t.all_gather_object(lines, line)
    if rank == 0:
        print("\n".join(lines), flush=True)
    dist_t.barrier()
    dist_t.destroy_process_group()
else:
    print(line, flush=True)
