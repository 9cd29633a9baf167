#!/usr/bin/env python
"""Step and per-phase device times of an H x W grid slab-decomposed over the ranks of a torchrun job (or the whole grid through
pano_dist with one rank).  2048 x 8192 on 2 GPUs = the per-GPU slab of BASELINE configs[2] (8192^2) on 8 GPUs.
usage: [torchrun --nproc-per-node N] scripts/time_slab.py H W [key=value ...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import panopaea_b200 as P
from panopaea_b200 import dist

rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
h, w = int(sys.argv[1]), int(sys.argv[2])
opts = dict(kv.split("=") for kv in sys.argv[3:])
ky, kx = h // 128, w // 128
prm = dict(timestep=0.05, threshold=0.1, max_iterations=100, inflow=(5 * ky, 20 * ky, 54 * kx, 64 * kx), inflow_density=1.0,
           inflow_vy=20.0, obstacle=(70 * ky, 80 * ky, 50 * kx, 70 * kx))
import torch
import torch.distributed as dist_t
torch.cuda.set_device(local)
ctx = P.Context(local)
for k, v in opts.items():
    ctx.set_option(k, int(v))
D = dist.DistFluid(ctx, h, w, rank, world, prm)
if world > 1:
    dist_t.init_process_group("nccl", device_id=torch.device("cuda", local))
    handles = [None] * world
    dist_t.all_gather_object(handles, D.ipc_handle())
    D.connect_ipc(handles)
    dist_t.barrier()


def sync():
    out = D.sync()
    if world > 1:
        dist_t.barrier()
    return out


for _ in range(10):
    D.step()
sync()
K = 20
ctx.set_option("step_timing", 1)
ctx.step_times()
ctx.timer_mark()
for _ in range(K):
    D.step()
    ctx.timer_mark()
laps = sorted(ctx.timer_marks_ms())
inf = sync()
pm, ps = ctx.step_times()
line = (f"rank {rank}/{world} grid {h}x{w} opts {opts}: step median {laps[len(laps) // 2]:.4f} ms (min {laps[0]:.4f}); phases ms "
        + " ".join(f"{n}={m / max(1, ps):.4f}" for n, m in zip(("inflow+ex", "advect", "negdiv", "cg", "project"), pm))
        + f"; cg per iteration {pm[3] / max(1, ps) / max(1, inf['applies']) * 1e3:.2f} us; info {inf}")
if world > 1:
    lines = [None] * world
    dist_t.all_gather_object(lines, line)
    if rank == 0:
        print("\n".join(lines), flush=True)
    dist_t.barrier()
    dist_t.destroy_process_group()
else:
    print(line, flush=True)
