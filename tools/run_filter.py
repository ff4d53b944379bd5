"""A few FFTFilter2D steps (for ncu captures): python tools/run_filter.py [images]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from impulse_b200.filter import FFTFilter2D
b = int(sys.argv[1]) if len(sys.argv) > 1 else 16
x = torch.rand((b, 4096, 4096), device="cuda", dtype=torch.float32)
k = torch.rand((31, 31), device="cuda", dtype=torch.float32)
f = FFTFilter2D(k / k.sum(), 4096, 4096)
y = torch.empty_like(x)
for _ in range(2):
    f.apply(x, out=y)
torch.cuda.synchronize()
