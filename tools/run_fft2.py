"""Run a 2-D complex transform a few times (for ncu captures): python tools/run_fft2.py [n]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import impulse_b200 as ib
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
x = torch.view_as_complex(torch.rand((n, n, 2), device="cuda", dtype=torch.float64) - 0.5)
y = torch.empty_like(x)
f = ib.FFTDesc.init(axes=[0, 1], forward=True)
for _ in range(3):
    f.apply(ib.DataDesc.init(y), ib.DataDesc.init(x))
torch.cuda.synchronize()
print(ib.last_kernel())
