"""torchrun --nproc-per-node N tools/slab_variants.py: slab fft2 8192^2 c128, fused peer loads vs pipelined pull (J chunks)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from impulse_b200 import dist as idist

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
dist.init_process_group("nccl")
n = 8192
lo, hi = idist.shard_rows(n, rank, world)
x = torch.view_as_complex(torch.rand((hi - lo, n, 2), device="cuda", dtype=torch.float64) - 0.5)
ref = None
for pull, ctas in ((0, 0), (1, 64), (2, 64), (4, 64), (2, 32), (4, 148)):
    op = idist.SlabFFT2P2P(hi - lo, n, torch.complex128, pull_chunks=pull, copy_ctas=ctas)
    for _ in range(5):
        y = op(x)
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        y = op(x)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / 20], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if ref is None:
        ref = y.clone()
    err = float(torch.linalg.vector_norm(y - ref) / torch.linalg.vector_norm(ref))
    if rank == 0:
        print(f"world={world} pull_chunks={op.pull_chunks} copy_ctas={ctas}: {float(t):.4f} ms  (rel diff vs fused {err:.1e})", flush=True)
    op.close()
dist.destroy_process_group()
