import os, sys, ctypes as C
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from impulse_b200 import _lib
from torch.multiprocessing.reductions import reduce_tensor
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
buf = torch.full((64, 96), float(rank + 1), dtype=torch.complex128, device=dev)
g = [None] * world
dist.all_gather_object(g, (local, reduce_tensor(buf)))
torch.cuda.synchronize()
L = _lib.lib()
for q in range(world):
    if q == rank: continue
    pdev, (fn, args) = g[q]
    v = fn(*args)
    print(rank, "opened peer", q, "tensor device", v.device, "ptr", hex(v.data_ptr()), flush=True)
    print(rank, "can access", torch.cuda.can_device_access_peer(local, pdev), flush=True)
    rc = L.impulse_fft_enable_peer_access(int(pdev))
    print(rank, "enable rc", rc, L.impulse_fft_last_error(), flush=True)
    t = torch.empty((64, 96), dtype=torch.complex128, device=dev)
    t.copy_(v); torch.cuda.synchronize()
    print(rank, "torch copy ok, value", t[0, 0].item(), flush=True)
    out = torch.zeros((64, 96), dtype=torch.complex128, device=dev)
    rc = L.impulse_fft_copy2d(1, C.c_void_p(v.data_ptr()), C.c_void_p(out.data_ptr()), 64, 96, 96, 96, 1, 0, 0, None)
    print(rank, "copy2d rc", rc, L.impulse_fft_last_error(), flush=True)
    torch.cuda.synchronize()
    print(rank, "raw kernel read ok, value", out[5, 7].item(), flush=True)
dist.barrier()
dist.destroy_process_group()
