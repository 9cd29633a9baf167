"""Real multi-process run of the slab-decomposed step (one rank per GPU, CUDA IPC + NVLink):
  torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/dist_mp_check.py [n] [steps]
Rank 0 also runs the single-GPU step on the same problem and compares."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist_t

import panopaea_b200 as P
from panopaea_b200 import dist, fluid

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist_t.init_process_group("nccl", device_id=torch.device("cuda", local))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
prm = fluid.smoke_params(n)
ctx = P.Context(local)
D = dist.DistFluid(ctx, n, n, rank, world, {k: v for k, v in prm.items() if k not in ("h", "w")})
handles = [None] * world
dist_t.all_gather_object(handles, D.ipc_handle())
D.connect_ipc(handles)
dist_t.barrier()
single = fluid.DecFluid(**prm, ctx=ctx) if rank == 0 else None
ok = True
for s in range(steps):
    D.step()
    info = D.sync()
    parts = [None] * world
    dist_t.all_gather_object(parts, (D.download(dist.DENSITY), D.download(dist.VY), D.download(dist.PRESSURE)))
    if rank == 0:
        want = single.step()
        d = np.concatenate([p[0] for p in parts]); vy = np.concatenate([p[1] for p in parts]); pr = np.concatenate([p[2] for p in parts])
        svy, _ = single.vel.split()
        e_d = np.abs(d - single.density.to_host()).max()
        e_v = np.abs(vy - svy).max() / max(1.0, np.abs(svy).max())
        e_p = np.abs(pr - single.pressure.to_host()).max() / max(1.0, np.abs(single.pressure.to_host()).max())
        good = abs(info["iterations"] - want["iterations"]) <= 1 and e_d < 1e-9 and (info["iterations"] != want["iterations"] or (e_v < 1e-8 and e_p < 1e-8))
        ok &= good
        print(f"step {s}: dist {info} single {want['iterations']} err density {e_d:.1e} vy {e_v:.1e} p {e_p:.1e} {'OK' if good else 'MISMATCH'}", flush=True)
        if info["iterations"] != want["iterations"]:
            break
if rank == 0:
    print("DIST_MP_CHECK", "PASS" if ok else "FAIL", flush=True)
dist_t.barrier()
dist_t.destroy_process_group()
