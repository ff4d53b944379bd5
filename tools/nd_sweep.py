"""Kernel selection and throughput for common N-D shapes: python tools/nd_sweep.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import impulse_b200 as ib

CASES = [("c2c", "f32", (1024, 256, 256), [1, 2]), ("c2c", "f64", (512, 256, 256), [1, 2]), ("r2c", "f32", (1024, 256, 256), [1, 2]),
         ("r2c", "f32", (256, 512, 512), [1, 2]), ("r2c", "f64", (128, 1024, 1024), [1, 2]), ("c2c", "f32", (64, 1024, 1024), [1, 2]),
         ("c2c", "f32", (64, 64, 64, 64), [1, 2, 3]), ("r2c", "f32", (32, 128, 128, 128), [1, 2, 3]), ("c2c", "f64", (4096, 4096), [0, 1]),
         ("c2c", "f64", (1000, 1000), [0, 1]), ("r2c", "f64", (2000, 3000), [0, 1]), ("c2c", "f32", (16, 480, 640), [1, 2]),
         ("c2c", "f32", (32, 2048, 2048), [1, 2]), ("c2c", "f64", (16, 2048, 2048), [1, 2]), ("r2c", "f32", (64, 2048, 2048), [1, 2]),
         ("c2c", "f64", (32, 1024, 1024), [1, 2])]
if len(sys.argv) > 1:
    CASES = [c for c in CASES if any(str(d) == sys.argv[1] for d in c[2])]
for kind, dt, shape, axes in CASES:
    rdt = torch.float64 if dt == "f64" else torch.float32
    cdt = torch.complex128 if dt == "f64" else torch.complex64
    if kind == "c2c":
        x = torch.view_as_complex(torch.rand(shape + (2,), device="cuda", dtype=rdt) - 0.5)
        y = torch.empty_like(x)
    else:
        x = torch.rand(shape, device="cuda", dtype=rdt) - 0.5
        y = torch.empty(shape[:-1] + (shape[-1] // 2 + 1,), device="cuda", dtype=cdt)
    f = ib.FFTDesc.init(axes=axes, forward=True)
    din, dout = ib.DataDesc.init(x), ib.DataDesc.init(y)
    kernels = []
    n0 = ib.launch_count()
    f.apply(dout, din)
    nl = ib.launch_count() - n0
    for _ in range(2):
        f.apply(dout, din)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        f.apply(dout, din)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    nbytes = x.numel() * x.element_size() + y.numel() * y.element_size()
    print(f"{kind} {dt} {shape} axes={axes}: {ms:.3f} ms, {nbytes / ms / 1e6:.0f} GB/s (one read + one write), {nl} launches, last {ib.last_kernel()}", flush=True)
