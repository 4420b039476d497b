#!/usr/bin/env python
"""Host-link ceiling of the box for the end-to-end path (DESIGN.md, e2e scaling): every rank copies a pinned 42 MB buffer (the
observation block of one 131072-env step) device -> host, all ranks at once; prints per-rank and aggregate GB/s.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 tools/gpu/d2h_probe.py
"""
import os
import time

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
nbytes = 131072 * 322
src = torch.empty(nbytes, dtype=torch.uint8, device=dev)
dst = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
hsrc = torch.empty(131072 * 8, dtype=torch.uint8).pin_memory()
for _ in range(5):
    dst.copy_(src, non_blocking=True)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
reps = 40
t0 = time.perf_counter()
for _ in range(reps):
    dst.copy_(src, non_blocking=True)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
gbs = nbytes * reps / dt / 1e9
t = torch.tensor([gbs], dtype=torch.float64, device=dev)
lo = t.clone()
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f'{{"d2h_probe": {{"n_gpus": {world}, "bytes": {nbytes}, "aggregate_GBps": {t.item():.2f}, "slowest_rank_GBps": {lo.item():.2f}, '
          f'"cpus": {os.cpu_count()}}}}}', flush=True)
if world > 1:
    dist.destroy_process_group()
