"""Bare page-locked device-to-host bandwidth of this box, one process per GPU (is ~93 GB/s aggregate the
ceiling of the host side?).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 profiles/scripts/d2h_probe.py

Every rank allocates a pinned 1 GiB buffer, streams a resident 1 GiB device buffer into it for ~3 s with cudaMemcpyAsync on its
own stream (all ranks between two barriers), and rank 0 prints one JSON line with per-rank and aggregate GB/s -- no kernels,
no engine, nothing but the copies the end-to-end path issues."""
import json
import os
import time

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
nbytes = 1 << 30
dev = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
host = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
host.copy_(dev, non_blocking=True)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
n = 0
while time.perf_counter() - t0 < 3.0:
    for _ in range(4):
        host.copy_(dev, non_blocking=True)
    torch.cuda.synchronize()
    n += 4
dt = time.perf_counter() - t0
gbs = torch.tensor([n * nbytes / dt * 1e-9], dtype=torch.float64, device="cuda")
if world > 1:
    allg = [torch.zeros_like(gbs) for _ in range(world)]
    dist.all_gather(allg, gbs)
    vals = [float(x.item()) for x in allg]
else:
    vals = [float(gbs.item())]
if rank == 0:
    print(json.dumps({"probe": "pinned D2H, 1 GiB cudaMemcpyAsync per copy, all ranks concurrently", "n_gpus": world,
                      "per_rank_gbs": [round(v, 1) for v in vals], "aggregate_gbs": round(sum(vals), 1),
                      "host_cpus": os.cpu_count()}))
if world > 1:
    dist.destroy_process_group()
