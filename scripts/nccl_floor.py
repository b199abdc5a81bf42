"""Latency floor of the multi-GPU step tail on the latency-bound configs (C1 / C2): one ncclAllReduce of [P + 3] = 3,024 floats
(12 KB) per step.  Run under torchrun; prints the device time per all-reduce (CUDA events over 200 back-to-back calls, max over
ranks) for a few payloads.  Usage: torchrun --nproc-per-node N scripts/nccl_floor.py"""
import os
import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
for n in (3024, 12740, 83204, 462342):
    x = torch.ones(n, device="cuda", dtype=torch.float32)
    for _ in range(20):
        dist.all_reduce(x)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 200
    e0.record()
    for _ in range(iters):
        dist.all_reduce(x)
    e1.record(); torch.cuda.synchronize()
    us = torch.tensor([e0.elapsed_time(e1) / iters * 1e3], device="cuda")
    dist.all_reduce(us, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"ranks {world}: all-reduce of {n} floats ({n * 4 / 1024:.0f} KB): {us.item():.1f} us per call", flush=True)
dist.destroy_process_group()
