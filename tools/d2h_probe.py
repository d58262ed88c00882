#!/usr/bin/env python
"""tools/d2h_probe.py: aggregate device-to-host copy rate of the box with N ranks copying at once (run under torchrun; pinned
host buffers, 256 MB per copy).  Evidence for what bounds the end-to-end leg at N GPUs: the host side of the box."""
import os, time
import torch, torch.distributed as dist
rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
n = 256 << 20
src = torch.empty(n, dtype=torch.uint8, device="cuda")
dst = [torch.empty(n, dtype=torch.uint8, pin_memory=True) for _ in range(2)]
for i in range(4):
    dst[i & 1].copy_(src, non_blocking=True)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
reps = 40
for i in range(reps):
    dst[i & 1].copy_(src, non_blocking=True)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
t = torch.tensor([dt], dtype=torch.float64, device="cuda")
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print("ranks %d: %.1f GB/s per rank, %.1f GB/s aggregate (device to pinned host, 256 MB copies)" % (
        world, reps * n / t.item() / 1e9, world * reps * n / t.item() / 1e9), flush=True)
if world > 1:
    dist.destroy_process_group()
