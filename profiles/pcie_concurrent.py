"""Host <-> device copy bandwidth per GPU when 1 ... N ranks copy at the same time (pinned memory, 256 MB, both directions):
   python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 profiles/pcie_concurrent.py
Explains the end-to-end scaling of bench.py at N = 8 (every call of the plugin boundary moves host vectors)."""
import os
import time

import torch
import torch.distributed as dist

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
h = torch.empty(32 * 1024 * 1024, dtype=torch.float64).pin_memory()
d = torch.empty_like(h, device="cuda")


def bw(active, direction):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t = 0.0
    if rank < active:
        t0 = time.perf_counter()
        for _ in range(8):
            if direction == "h2d":
                d.copy_(h, non_blocking=True)
            else:
                h.copy_(d, non_blocking=True)
        torch.cuda.synchronize()
        t = time.perf_counter() - t0
    out = torch.tensor([t], device="cuda")
    if world > 1:
        dist.all_reduce(out, op=dist.ReduceOp.MAX)
    return 8 * h.numel() * 8 / out.item() / 1e9


for direction in ("h2d", "d2h"):
    bw(world, direction)  # warm-up
    for active in sorted({1, 2, 4, world} & set(range(1, world + 1))):
        g = bw(active, direction)
        if rank == 0:
            print(f"{direction}: {active} rank(s) copying at once: {g:6.1f} GB/s per GPU (slowest), {g * active:7.1f} GB/s aggregate", flush=True)
if world > 1:
    dist.destroy_process_group()
