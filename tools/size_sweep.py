"""Throughput of batched 1-D transforms over a range of lengths (device resident, out of place):
which kernel serves each length and how far it is from the copy roofline.
Usage: python tools/size_sweep.py [--kinds c2c,r2c,c2r] [--dtypes f64,f32] [--lengths 512,1000,...]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import impulse_b200 as ib

LENGTHS = [16, 32, 64, 100, 128, 243, 256, 500, 512, 625, 1000, 1024, 1536, 2000, 2048, 2187, 3000, 4096, 4099, 5000, 6561,
           8192, 10000, 12288, 16384, 32768, 65536, 100003, 262144, 1048576]
TOTAL = 1 << 25   # complex elements per transform batch (512 MiB of complex128)


def run(kind, dt, n):
    cdt = torch.complex128 if dt == "f64" else torch.complex64
    rdt = torch.float64 if dt == "f64" else torch.float32
    rows = max(1, TOTAL // n)
    if kind == "c2c":
        x = torch.view_as_complex(torch.rand((rows, n, 2), device="cuda", dtype=rdt) - 0.5)
        y = torch.empty_like(x)
        nbytes = 2 * x.numel() * x.element_size()
    else:
        x = torch.rand((rows, n), device="cuda", dtype=rdt) - 0.5
        y = torch.empty((rows, n // 2 + 1), device="cuda", dtype=cdt)
        nbytes = x.numel() * x.element_size() + y.numel() * y.element_size()
    f = ib.FFTDesc.init(axes=[1], forward=kind != "c2r")
    din, dout = ib.DataDesc.init(x), ib.DataDesc.init(y)
    if kind == "c2r":   # half spectrum in, real rows out (the spectrum of x, so the values are representative)
        ib.FFTDesc.init(axes=[1], forward=True).apply(dout, din)
        din, dout = dout, din
    for _ in range(3):
        f.apply(dout, din)
    torch.cuda.synchronize()
    n0 = ib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record()
    for _ in range(reps):
        f.apply(dout, din)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    return nbytes / ms / 1e6, ms, (ib.launch_count() - n0) // reps, ib.last_kernel()


if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--kinds", default="c2c,r2c")
    ap.add_argument("--dtypes", default="f64,f32")
    ap.add_argument("--lengths", default="")
    a = ap.parse_args()
    lengths = [int(v) for v in a.lengths.split(",")] if a.lengths else LENGTHS
    for kind, dt in [(k, d) for k in a.kinds.split(",") for d in a.dtypes.split(",")]:
        for n in lengths:
            try:
                gbs, ms, nl, k = run(kind, dt, n)
                print(f"{kind} {dt} n={n:8d} {gbs:8.0f} GB/s {ms:8.3f} ms launches={nl} {k}", flush=True)
            except Exception as ex:  # noqa: BLE001
                print(f"{kind} {dt} n={n:8d} ERROR {ex}", flush=True)
