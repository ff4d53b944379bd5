"""One slab-decomposed fft2 through the single-process multi-GPU C ABI (impulse_fft_dist_*), device shards in, column
slabs out — for ncu captures of the peer-loading column kernels: python tools/run_slab_dist.py [n] [ndev]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from impulse_b200 import _lib, dist
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
g = int(sys.argv[2]) if len(sys.argv) > 2 else torch.cuda.device_count()
devs = list(range(g))
p = dist.DistPlan("slab", _lib.C2C, _lib.F64, (n, n), (n * 16, 16), (n * 16, 16), [0, 1], True, devs)
ins = [torch.view_as_complex(torch.rand((n // g, n, 2), device=f"cuda:{d}", dtype=torch.float64) - 0.5) for d in devs]
outs = [torch.empty((n, n // g), dtype=torch.complex128, device=f"cuda:{d}") for d in devs]
import time
for _ in range(3):
    p.run_parts(ins, outs, 1.0)
t0 = time.perf_counter()
for _ in range(10):
    p.run_parts(ins, outs, 1.0)
print(f"slab fft2 {n}^2 on {g} devices (single process, synchronous per call): {(time.perf_counter() - t0) * 100:.3f} ms per call")
p.close()
