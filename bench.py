#!/usr/bin/env python
"""bench.py — headline benchmark of the FFT hot path (BASELINE.json metric / configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

One "step" = one pass of the hot path over one batch: batched complex128 c2c FFT, 65536 rows x
1024 points per GPU (BASELINE config 2), out of place, synthetic uniform[-0.5,0.5) input.
Per-GPU work is fixed (rows are independent: batch sharding, no collective) -> "scaling": "weak".

Prints ONE JSON line (rank 0).  `value` = algorithmic GB/s with inputs resident in HBM (CUDA
events on the launching stream, max over ranks); `e2e` = the same metric through the public
C-ABI call with pinned HOST buffers (H2D + kernel + D2H inside the timed region); `roofline`
= the dominant kernel against the measured HBM copy peak; `cpu_baseline` = the reference's own
pocketfft timed on this box's host cores.  `--impl reference` times only that CPU path.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (kind, rows, n, dtype)  — bytes/row are the SURVEY 8(d) algorithmic figures
    "c2c_65536x1024_c128": ("c2c", 65536, 1024, "f64"),
    "r2c_1024x4096_f64": ("r2c", 1024, 4096, "f64"),
    "r2c_16384x1000_f64": ("r2c", 16384, 1000, "f64"),
    "r2c_16384x3888_f64": ("r2c", 16384, 3888, "f64"),
    "r2c_16384x4099_f64": ("r2c", 16384, 4099, "f64"),
    "c2r_16384x1000_f64": ("c2r", 16384, 1000, "f64"),
    "c2r_16384x3888_f64": ("c2r", 16384, 3888, "f64"),
    "c2r_16384x4099_f64": ("c2r", 16384, 4099, "f64"),
    "c2c_16384x4096_c128": ("c2c", 16384, 4096, "f64"),
    "c2c_8192x8192_c128": ("c2c", 8192, 8192, "f64"),
    "c2c_131072x1024_c64": ("c2c", 131072, 1024, "f32"),
    "c2c_131072x512_c128": ("c2c", 131072, 512, "f64"),
    "c2c_262144x256_c128": ("c2c", 262144, 256, "f64"),
    "fft2_8192x8192_c128": ("fft2", 8192, 8192, "f64"),
    "filter2d_64x4096x4096_f32": ("filter2d", 64, 4096, "f32"),
    "fftconvolve_4096x16384_k257_f64": ("fftconv", 4096, 16384, "f64"),   # rows (*) 257-tap FIR, full
}
FIR_TAPS = 257
DEFAULT_WORKLOAD = "c2c_65536x1024_c128"


def algorithmic_bytes(kind, rows, n, dtype):
    r = 8 if dtype == "f64" else 4
    if kind == "c2c":
        return rows * n * 2 * r * 2          # read + write one complex element each
    if kind == "fft2":
        return 2 * rows * n * 2 * r * 2      # two passes, each one read + one write
    if kind == "fftconv":                    # signal read once, full convolution written once
        return rows * (n + n + FIR_TAPS - 1) * r
    if kind == "filter2d":                   # rows = images, n = P: SURVEY 8(d) config 5 (circular, P x P)
        return rows * (2 * r * n * n + 8 * n * (n // 2 + 1) * r)
    return rows * (n * r + (n // 2 + 1) * 2 * r)  # real side + half-spectrum side


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def committed_traffic(workload):
    """dram bytes per launch of the dominant kernel from the committed ncu --set full capture."""
    p = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    try:
        return json.load(open(p)).get(workload)
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled while the timed region runs (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.t = [], None, None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.06)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        rows = [r for ts, r in self.rows if t0 - 0.05 <= ts <= t1 + 0.05] or [r for _, r in self.rows]
        sm, mx, reasons = [], None, set()
        for r in rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def cpu_reference(kind, rows, n, dtype, budget_s=12.0):
    """Time the reference's own CPU implementation (oracle/_ref, else the port) on a bounded
    sample of the workload with all host threads.  Returns (GB/s, info dict)."""
    from oracle import oracle
    chk = oracle.load()
    cores = max(1, chk.hardware_threads() if chk.kind == "reference" else 1)
    rng = np.random.default_rng(1234)
    srows = rows
    # bound the sample: ~7.3 us per 1024-pt row per core measured in the survey; keep each rep <= ~2 s
    est = rows * n * np.log2(max(n, 2)) * 0.75e-9 / cores * (1 if chk.kind == "reference" else 60)
    while est > 2.0 and srows > 64:
        srows //= 2
        est /= 2
    if dtype != "f64":
        raise SystemExit("cpu baseline implemented for the float64 workloads")
    if kind == "fft2":
        srows = rows
        x = rng.uniform(-0.5, 0.5, (rows, n)) + 1j * rng.uniform(-0.5, 0.5, (rows, n))
        y = np.empty_like(x)
        run = lambda: chk.c2c(x, [0, 1], True, 1.0, out=y, nthreads=0)
    elif kind == "c2c":
        x = rng.uniform(-0.5, 0.5, (srows, n)) + 1j * rng.uniform(-0.5, 0.5, (srows, n))
        run = lambda: chk.cfft_rows(x, True, 1.0, nthreads=cores)
    elif kind == "r2c":
        x = rng.uniform(-0.5, 0.5, (srows, n))
        run = lambda: chk.rfft_rows(x, True, 1.0, nthreads=cores)
    else:
        x = rng.uniform(-0.5, 0.5, (srows, n))
        run = lambda: chk.rfft_rows(x, False, 1.0 / n, nthreads=cores)
    run()  # warm-up (page faults, plan)
    best, reps, t_all = float("inf"), 0, time.perf_counter()
    while reps < 5 or (time.perf_counter() - t_all < budget_s and reps < 50):
        t = time.perf_counter()
        run()
        best = min(best, time.perf_counter() - t)
        reps += 1
        if time.perf_counter() - t_all > budget_s:
            break
    gbs = algorithmic_bytes(kind, srows, n, dtype) / best / 1e9
    info = {"value": round(gbs, 3), "unit": "GB/s", "cores": cores, "kind": chk.kind,
            "sample": f"{srows} of {rows} rows x {n} ({kind} {dtype}), in place, one shared plan, best of {reps}",
            "ms": round(best * 1e3, 3), "elements_per_s": round(srows * n / best, 1)}
    return gbs, info


def accuracy_report(ib, torch, kind, x, y, n, dtype):
    """max over (sampled) rows of the rel-L2 error against the oracle with its pass bound, and the
    forward->backward round trip over ALL rows on the device.  x = the step's input, y = its output (already
    computed by the timed steps).  The oracle is the checker here, never the thing measured."""
    from oracle import oracle
    chk = oracle.load()
    rows = x.shape[0]
    sel = sorted({0, 1, rows // 3, rows // 2, rows - 1})
    xs, ys = x[sel].cpu().numpy().copy(), y[sel].cpu().numpy()
    if kind == "c2c":
        want = chk.c2c(xs, [1], True, 1.0)
    elif kind == "r2c":
        want = chk.r2c(xs, [1], True, 1.0)
    elif kind == "c2r":
        want = chk.c2r(xs, ys.shape, [1], False, 1.0)
    else:
        return {"unavailable": f"not reported for {kind}"}
    bound = (1e-12 if dtype == "f64" else 1e-5) * max(1.0, float(np.log2(max(n, 2))))
    err = float(oracle.max_row_rel_l2(ys, want))
    out = {"max_rel_l2_vs_oracle": err, "bound": bound, "rows_checked": len(sel), "oracle": chk.kind, "pass": bool(err <= bound)}
    if kind in ("c2c", "r2c"):   # round trip: backward transform of the output with 1/N, against the input
        back = torch.empty_like(x)
        ib.FFTDesc.init(axes=[1], forward=False, scalingFactor=1.0 / n).apply(ib.DataDesc.init(back), ib.DataDesc.init(y))
        num = torch.linalg.vector_norm(back - x, dim=1)
        den = torch.linalg.vector_norm(x, dim=1)
        out["round_trip_max_rel_l2"] = float((num / den).max())
    return out


def run_reference(args, kind, rows, n, dtype, rank, world):
    if rank != 0:
        return
    steps = max(1, args.steps)
    gbs, info = cpu_reference(kind, rows, n, dtype, budget_s=min(60.0, 2.0 * (steps + args.warmup)))
    line = {"impl": "reference", "metric": "batched fp64 FFT throughput (algorithmic GB/s)", "value": info["value"],
            "unit": "GB/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": info["ms"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": dtype, "data": "synthetic", "config": {"workload": args.workload, "sample": info["sample"]},
            "cpu_baseline": info,
            "e2e": {"value": info["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    kind, rows, n, dtype = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, kind, rows, n, dtype, rank, world)
        return

    import torch
    import torch.distributed as dist

    import impulse_b200 as ib

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: impulse_b200 has no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG", "WARN")  # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=dev)
    warmup = max(3, args.warmup)
    steps = max(1, args.steps)

    rdt = torch.float64 if dtype == "f64" else torch.float32
    cdt = torch.complex128 if dtype == "f64" else torch.complex64
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    slab = kind == "fft2" and world > 1
    filt = None
    conv = None
    if kind == "fftconv":
        from impulse_b200 import signal as isig
        x = torch.rand((rows, n), generator=g, device=dev, dtype=rdt) - 0.5
        taps = torch.rand((FIR_TAPS,), generator=g, device=dev, dtype=rdt) - 0.5
        y = None
        conv = lambda: isig.fftconvolve(x, taps, "full")  # noqa: E731
    elif kind == "filter2d":
        from impulse_b200.filter import FFTFilter2D
        from impulse_b200 import dist as idist
        lo, hi = idist.shard_rows(rows, rank, world) if world > 1 else (0, rows)
        x = torch.rand((hi - lo, n, n), generator=g, device=dev, dtype=rdt)
        ker = torch.rand((31, 31), generator=g, device=dev, dtype=rdt)
        filt = FFTFilter2D(ker / ker.sum(), n, n)
        y = torch.empty_like(x)
    elif slab:
        # ONE 2-D transform split by row slabs: total work fixed ("strong"), one NCCL all-to-all per step
        from impulse_b200 import dist as idist
        lo, hi = idist.shard_rows(rows, rank, world)
        x = torch.view_as_complex(torch.rand((hi - lo, n, 2), generator=g, device=dev, dtype=rdt) - 0.5)
        y = None
        # default: exchange fused into the column kernels over peer memory; IMPULSE_FFT_SLAB=nccl selects
        # the pack + NCCL all-to-all variant
        slab_op = None if os.environ.get("IMPULSE_FFT_SLAB", "p2p") == "nccl" else idist.SlabFFT2P2P(hi - lo, n, cdt)
    elif kind in ("c2c", "fft2"):
        x = torch.view_as_complex(torch.rand((rows, n, 2), generator=g, device=dev, dtype=rdt) - 0.5)
        y = torch.empty_like(x)
    elif kind == "r2c":
        x = torch.rand((rows, n), generator=g, device=dev, dtype=rdt) - 0.5
        y = torch.empty((rows, n // 2 + 1), device=dev, dtype=cdt)
    else:
        x = torch.view_as_complex(torch.rand((rows, n // 2 + 1, 2), generator=g, device=dev, dtype=rdt) - 0.5)
        y = torch.empty((rows, n), device=dev, dtype=rdt)
    fdesc = ib.FFTDesc.init(axes=[0, 1] if kind == "fft2" else [1], forward=(kind != "c2r"), scalingFactor=1.0)
    if not slab and filt is None and conv is None:
        din, dout = ib.DataDesc.init(x), ib.DataDesc.init(y)

    def step():
        if conv is not None:
            conv()
        elif filt is not None:
            filt.apply(x, out=y)
        elif slab:
            if slab_op is not None:
                slab_op(x, True, 1.0)
            else:
                idist.fft2_slab(x, True, 1.0)
        else:
            fdesc.apply(dout, din)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        step()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    stream = torch.cuda.current_stream()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n0 = ib.launch_count()
    t_host0 = time.time()
    e0.record(stream)
    for _ in range(steps):
        step()
    e1.record(stream)
    barrier()
    t_host1 = time.time()
    launches = ib.launch_count() - n0
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / steps
    bytes_per_gpu = algorithmic_bytes(kind, rows, n, dtype)
    strong = slab or (filt is not None and world > 1)   # total work fixed, split over ranks
    value = (1 if strong else world) * bytes_per_gpu / (ms_per_step * 1e-3) / 1e9

    # ---- e2e: public API with pinned HOST buffers, copies inside the timed region
    e2e = None
    if not args.no_e2e and not slab and filt is None and conv is None:
        hx = torch.empty(x.shape, dtype=x.dtype, pin_memory=True)
        hy = torch.empty(y.shape, dtype=y.dtype, pin_memory=True)
        hx.copy_(x)
        hxn, hyn = hx.numpy(), hy.numpy()
        hin, hout = ib.DataDesc.init(hxn), ib.DataDesc.init(hyn)
        e_steps = max(2, min(steps, 5))
        for _ in range(2):
            fdesc.apply(hout, hin)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e_steps):
            fdesc.apply(hout, hin)   # synchronous for host pointers: H2D + kernel + D2H
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / e_steps
        if world > 1:
            t = torch.tensor([dt], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        chk = float(np.abs(hyn[0, :4]).sum())  # touch the result on the host
        e2e = {"value": round(world * bytes_per_gpu / dt / 1e9, 3), "unit": "GB/s",
               "h2d_bytes_per_step": int(hxn.nbytes), "d2h_bytes_per_step": int(hyn.nbytes),
               "ms_per_step": round(dt * 1e3, 3), "steps": e_steps, "host_memory": "pinned", "check": chk}
    clocks = sampler.stop(t_host0, time.time()) if sampler else None

    if rank == 0:
        peak, peak_src = peaks()
        # dominant (only) kernel of the step: one launch per step on this stream
        # (multi-launch steps — fft2, filter2d, fftconvolve: the step's algorithmic bytes over the whole step)
        per_step = max(1, round(launches / steps))
        k_ms = e0.elapsed_time(e1) / steps if per_step > 1 else e0.elapsed_time(e1) / max(1, launches)
        achieved = bytes_per_gpu / (k_ms * 1e-3) / 1e9
        line = {
            "metric": "batched fp64 FFT throughput (algorithmic GB/s)" if dtype == "f64" else "batched fp32 FFT throughput (algorithmic GB/s)",
            "value": round(value, 2), "unit": "GB/s", "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": round(ms_per_step, 5), "higher_is_better": True, "scaling": "strong" if strong else "weak",
            "vs_baseline": None, "dtype": dtype, "data": "synthetic",
            "config": {"workload": args.workload, "rows_per_gpu": (rows // world) if slab else rows, "length": n, "kind": kind,
                       "placement": "out of place, device resident", "l2": "input+output per step exceed the 126 MB L2"
                       if bytes_per_gpu > 2 * 126e6 else "working set fits L2: reported as is, see DESIGN.md",
                       "parallelism": (f"row-slab x{world}, " + ("column kernels load peers' row slabs over NVLink (CUDA IPC), one 1-element all-reduce as barrier"
                                                                if slab_op is not None else "pack + NCCL all-to-all") + ", result left in column slabs" if slab
                                       else f"batch-shard x{world}, no collective"),
                       "elements_per_s": round((1 if slab else world) * rows * n / (ms_per_step * 1e-3), 1),
                       "frac_of_8TBps_nominal": round(value / world / 8000.0, 4)},
            "roofline": {"bound": "hbm", "achieved": round(achieved, 2), "peak": peak, "unit": "GB/s",
                         "frac": round(achieved / peak, 4), "traffic": committed_traffic(args.workload),
                         "peak_source": peak_src, "kernel": ib.last_kernel() + (f" (last of {per_step} launches per step)" if per_step > 1 else ""),
                         "algorithmic_bytes_per_launch": bytes_per_gpu},
            "gpu_launches": int(launches), "clocks": clocks,
        }
        if e2e:
            line["e2e"] = e2e
        if not args.no_cpu and world == 1 and filt is None and conv is None:
            try:
                _, info = cpu_reference(kind, rows, n, dtype)
                line["cpu_baseline"] = info
            except Exception as ex:  # the oracle is optional for the measurement itself
                line["cpu_baseline"] = {"unavailable": str(ex)}
            try:
                line["accuracy"] = accuracy_report(ib, torch, kind, x, y, n, dtype)
            except Exception as ex:  # noqa: BLE001 — a reporting extra never costs the bench line
                line["accuracy"] = {"unavailable": str(ex)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
