"""Run one batched transform a few times (for ncu captures): python tools/run_one.py c2c f64 3000 [rows]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import impulse_b200 as ib
kind, dt, n = sys.argv[1], sys.argv[2], int(sys.argv[3])
rows = int(sys.argv[4]) if len(sys.argv) > 4 else max(1, (1 << 25) // n)
rdt = torch.float64 if dt == "f64" else torch.float32
cdt = torch.complex128 if dt == "f64" else torch.complex64
if kind == "c2c":
    x = torch.view_as_complex(torch.rand((rows, n, 2), device="cuda", dtype=rdt) - 0.5)
    y = torch.empty_like(x)
else:
    x = torch.rand((rows, n), device="cuda", dtype=rdt) - 0.5
    y = torch.empty((rows, n // 2 + 1), device="cuda", dtype=cdt)
f = ib.FFTDesc.init(axes=[1], forward=kind != "c2r")
if kind == "c2r":   # spectrum of x in, real rows out
    ib.FFTDesc.init(axes=[1], forward=True).apply(ib.DataDesc.init(y), ib.DataDesc.init(x))
    x, y = y, x
for _ in range(3):
    f.apply(ib.DataDesc.init(y), ib.DataDesc.init(x))
torch.cuda.synchronize()
print(ib.last_kernel())
