"""Aggregate host<->device copy bandwidth with one process per GPU, all ranks copying at once (no kernels): the ceiling of
any end-to-end number with host buffers on this box.  torchrun --nproc-per-node N tools/pcie_aggregate.py"""
import os
import time
import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 1 << 30
hin = torch.empty(n, dtype=torch.uint8, pin_memory=True)
hout = torch.empty(n, dtype=torch.uint8, pin_memory=True)
din = torch.empty(n, dtype=torch.uint8, device="cuda")
dout = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
res = {}
for mode in ("h2d", "d2h", "both"):
    for rep in range(3):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        if mode in ("h2d", "both"):
            with torch.cuda.stream(s1):
                din.copy_(hin, non_blocking=True)
        if mode in ("d2h", "both"):
            with torch.cuda.stream(s2):
                hout.copy_(dout, non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res[mode] = min(res.get(mode, 1e9), float(t.item()))
if rank == 0:
    for mode, dt in res.items():
        per_dir = world * n / dt / 1e9
        print(f"n_gpus={world} {mode}: {dt * 1e3:.2f} ms for 1 GiB per GPU per direction -> {per_dir:.1f} GB/s per direction aggregate ({per_dir / world:.1f} per GPU)")
if world > 1:
    dist.destroy_process_group()
