"""Time FFTFilter2D.apply on the device: python tools/time_filter.py B H W [dtype f32|f64]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import impulse_b200 as ib
from impulse_b200.filter import FFTFilter2D

b, h, w = (int(v) for v in sys.argv[1:4])
dt = torch.float64 if len(sys.argv) > 4 and sys.argv[4] == "f64" else torch.float32
x = torch.rand((b, h, w), device="cuda", dtype=dt)
k = torch.rand((31, 31), device="cuda", dtype=dt)
f = FFTFilter2D(k / k.sum(), h, w)
y = torch.empty_like(x)
for _ in range(3):
    f.apply(x, out=y)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 10
e0.record()
for _ in range(n):
    f.apply(x, out=y)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
print(f"filter {b}x{h}x{w} {sys.argv[4] if len(sys.argv) > 4 else 'f32'}: {ms:.3f} ms  ({2 * x.numel() * x.element_size() / ms * 1e-6:.0f} GB/s of image in+out)")
