"""Aggregate pinned-host copy bandwidth when every GPU of the box copies at once (the ceiling of the end-to-end
mrp_step_host numbers at N > 1):  torchrun --nproc-per-node N profiles/pcie_bw_multi.py"""
import os

import torch
import torch.distributed as dist

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
if world > 1:
    dist.init_process_group("nccl")
n = 174 * 1024 * 1024
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
hu = torch.empty(n // 3, dtype=torch.uint8).pin_memory()
du = torch.empty(n // 3, dtype=torch.uint8, device="cuda")


def timed(fn, reps=10):
    fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


ms_d2h = timed(lambda: h.copy_(d, non_blocking=True))
ms_h2d = timed(lambda: du.copy_(hu, non_blocking=True))
if rank == 0:
    print(f"{world} GPUs at once: D2H {n / 1e6:.0f} MB per GPU in {ms_d2h:.2f} ms = {n / ms_d2h / 1e6:.1f} GB/s per GPU, {world * n / ms_d2h / 1e6:.1f} GB/s aggregate; "
          f"H2D {n / 3e6:.0f} MB in {ms_h2d:.2f} ms = {n / 3 / ms_h2d / 1e6:.1f} GB/s per GPU", flush=True)
if world > 1:
    dist.destroy_process_group()
