#!/usr/bin/env python
"""Concurrent host<->device bandwidth of all ranks of a torchrun job, with the default CPU placement and with every rank bound to its
GPU's own NUMA node (nvmlDeviceSetCpuAffinity) before the pinned buffers are allocated.  usage: torchrun --nproc-per-node N this.py [MB]"""
import os
import sys
import time

import torch
import torch.distributed as dist

rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
mb = int(sys.argv[1]) if len(sys.argv) > 1 else 200
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.empty(mb * 1000000 // 8, dtype=torch.float64, device="cuda")


def measure(tag):
    host = torch.empty(mb * 1000000 // 8, dtype=torch.float64).pin_memory()
    host.fill_(1.0)
    out = []
    for direction in ("h2d", "d2h", "both"):
        host2 = torch.empty_like(host).pin_memory() if direction == "both" else None
        s2 = torch.cuda.Stream()
        for rep in range(3):
            dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            if direction == "h2d":
                dev.copy_(host, non_blocking=True)
            elif direction == "d2h":
                host.copy_(dev, non_blocking=True)
            else:
                dev.copy_(host, non_blocking=True)
                with torch.cuda.stream(s2):
                    host2.copy_(dev, non_blocking=True)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
        out.append(f"{direction} {mb / 1e3 / dt * (2 if direction == 'both' else 1):.1f} GB/s")
    cpus = sorted(os.sched_getaffinity(0))
    line = f"[{tag}] rank {rank} gpu {local} cpus {cpus[0]}..{cpus[-1]} ({len(cpus)}): " + ", ".join(out)
    lines = [None] * world
    dist.all_gather_object(lines, line)
    if rank == 0:
        print("\n".join(lines), flush=True)


measure("default placement")
try:
    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(local)
    pynvml.nvmlDeviceSetCpuAffinity(h)
    measure("bound to the GPU's NUMA node")
except Exception as e:   # noqa: BLE001
    if rank == 0:
        print("no NVML affinity:", e)
if rank == 0:
    os.system("nvidia-smi topo -m 2>&1 | head -30; (numactl -H 2>/dev/null || lscpu | grep -i numa) | head -12")
dist.destroy_process_group()
