#!/usr/bin/env python
"""bench.py — headline benchmark of the FFT hot path (BASELINE.json metric / configs[1]) plus every other
BASELINE config in the same JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME] [--no-configs]

One "step" = one pass of the hot path over one batch: batched complex128 c2c FFT, 65536 rows x
1024 points per GPU (BASELINE config 2), out of place, synthetic uniform[-0.5,0.5) input.
Per-GPU work is fixed (rows are independent: batch sharding, no collective) -> "scaling": "weak".

Prints ONE JSON line (rank 0).  `value` = algorithmic GB/s with inputs resident in HBM (CUDA
events on the launching stream, max over ranks); `e2e` = the same metric through the public
C-ABI call with pinned HOST buffers (H2D + kernel + D2H inside the timed region); `roofline`
= the dominant kernel against the measured HBM copy peak; `cpu_baseline` = the reference's own
pocketfft timed on this box's host cores; `accuracy` = EVERY row of the step's output against the oracle.
`configs` = the same measurements for BASELINE configs 1, 3a/3b/3c (both directions), 4 and 5, and at
N > 1 the partitioned variants: config 2 strong-scaled (65536 rows split over the ranks) and config 4 as a
slab-decomposed fft2 (both exchange variants) with its accuracy against the oracle.
`--impl reference` times only the CPU path.  The oracle is the checker / the CPU arm, never the thing measured.
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
L2_BYTES = 126e6
# the BASELINE configs reported in the `configs` block of the default run (config 2 is the main line)
CONFIG_BLOCK = [
    ("1_r2c_1024x4096", "r2c_1024x4096_f64"),
    ("3a_r2c_16384x1000", "r2c_16384x1000_f64"), ("3a_c2r_16384x1000", "c2r_16384x1000_f64"),
    ("3b_r2c_16384x3888", "r2c_16384x3888_f64"), ("3b_c2r_16384x3888", "c2r_16384x3888_f64"),
    ("3c_r2c_16384x4099", "r2c_16384x4099_f64"), ("3c_c2r_16384x4099", "c2r_16384x4099_f64"),
    ("4_fft2_8192x8192", "fft2_8192x8192_c128"),
    ("5_filter2d_64x4096x4096", "filter2d_64x4096x4096_f32"),
]


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


def committed_traffic(workload, kernel):
    """dram bytes per launch of the dominant kernel from the committed ncu --set full capture — only while the kernel
    that serves the workload is still the one the capture was taken on (null otherwise: no stale figure)."""
    p = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    try:
        ent = json.load(open(p)).get(workload)
        if isinstance(ent, dict) and ent.get("kernel") == kernel:
            return ent["bytes"]
    except Exception:
        pass
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


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference's own pocketfft (oracle/_ref), all host threads, bounded sample
# ------------------------------------------------------------------------------------------------
def filter_kernel_spectrum_cpu(chk, ker, n):
    pad = np.zeros((n, n), ker.dtype)
    ii = (np.arange(ker.shape[0]) - ker.shape[0] // 2) % n
    jj = (np.arange(ker.shape[1]) - ker.shape[1] // 2) % n
    pad[np.ix_(ii, jj)] = ker
    return chk.r2c(pad, [0, 1], True, 1.0, nthreads=0)


def cpu_reference(kind, rows, n, dtype, budget_s=12.0):
    """Time the reference's own CPU implementation (oracle/_ref, else the port) on a bounded
    sample of the workload with all host threads.  Returns (GB/s, info dict)."""
    from oracle import oracle
    chk = oracle.load()
    cores = max(1, chk.hardware_threads() if chk.kind == "reference" else 1)
    rng = np.random.default_rng(1234)
    rdt = np.float64 if dtype == "f64" else np.float32
    srows = rows
    # bound the sample: ~7.3 us per 1024-pt row per core measured in the survey; keep each rep <= ~2 s
    est = rows * n * np.log2(max(n, 2)) * 0.75e-9 / cores * (1 if chk.kind == "reference" else 60)
    if kind in ("r2c", "c2r") and n == 4099:
        est *= 7          # Bluestein rows cost ~7x per point
    if kind in ("c2c", "r2c", "c2r"):
        while est > 2.0 and srows > 64:
            srows //= 2
            est /= 2
    note = "in place, one shared plan (C engine)"
    if kind == "fft2":
        x = (rng.uniform(-0.5, 0.5, (rows, n)) + 1j * rng.uniform(-0.5, 0.5, (rows, n))).astype(np.complex128 if dtype == "f64" else np.complex64)
        y = np.empty_like(x)
        run = lambda: chk.c2c(x, [0, 1], True, 1.0, out=y, nthreads=0)  # noqa: E731
        note = "pocketfft::c2c axes=[0,1], nthreads=0"
    elif kind == "filter2d":
        srows = min(rows, 4)
        img = rng.uniform(0, 1, (srows, n, n)).astype(rdt)
        ker = rng.uniform(0, 1, (31, 31)).astype(rdt)
        kspec = filter_kernel_spectrum_cpu(chk, ker / ker.sum(), n)[None]
        spec = np.empty((srows, n, n // 2 + 1), np.complex128 if dtype == "f64" else np.complex64)
        out = np.empty_like(img)

        def run():
            chk.r2c(img, [1, 2], True, 1.0, out=spec, nthreads=0)
            np.multiply(spec, kspec, out=spec)
            chk.c2r(spec, img.shape, [1, 2], False, 1.0 / (n * n), out=out, nthreads=0)
        note = "pocketfft::r2c axes=[1,2] -> numpy multiply -> pocketfft::c2r, nthreads=0, kernel spectrum precomputed"
    elif kind == "c2c" and dtype == "f64":
        x = rng.uniform(-0.5, 0.5, (srows, n)) + 1j * rng.uniform(-0.5, 0.5, (srows, n))
        run = lambda: chk.cfft_rows(x, True, 1.0, nthreads=cores)  # noqa: E731
    elif kind == "c2c":
        x = (rng.uniform(-0.5, 0.5, (srows, n)) + 1j * rng.uniform(-0.5, 0.5, (srows, n))).astype(np.complex64)
        y = np.empty_like(x)
        run = lambda: chk.c2c(x, [1], True, 1.0, out=y, nthreads=0)  # noqa: E731
        note = "pocketfft::c2c<float> axes=[1], nthreads=0"
    elif kind in ("r2c", "c2r") and dtype == "f64":
        x = rng.uniform(-0.5, 0.5, (srows, n))
        fwd = kind == "r2c"
        run = lambda: chk.rfft_rows(x, fwd, 1.0 if fwd else 1.0 / n, nthreads=cores)  # noqa: E731
    elif kind == "r2c":
        x = rng.uniform(-0.5, 0.5, (srows, n)).astype(np.float32)
        y = np.empty((srows, n // 2 + 1), np.complex64)
        run = lambda: chk.r2c(x, [1], True, 1.0, out=y, nthreads=0)  # noqa: E731
        note = "pocketfft::r2c<float> axes=[1], nthreads=0"
    elif kind == "c2r":
        x = (rng.uniform(-0.5, 0.5, (srows, n // 2 + 1)) + 1j * rng.uniform(-0.5, 0.5, (srows, n // 2 + 1))).astype(np.complex64)
        y = np.empty((srows, n), np.float32)
        run = lambda: chk.c2r(x, y.shape, [1], False, 1.0 / n, out=y, nthreads=0)  # noqa: E731
        note = "pocketfft::c2r<float> axes=[1], nthreads=0"
    else:
        raise SystemExit(f"no cpu baseline for {kind}")
    run()  # warm-up (page faults, plan)
    best, reps, t_all = float("inf"), 0, time.perf_counter()
    while reps < 3 or (time.perf_counter() - t_all < budget_s and reps < 50):
        t = time.perf_counter()
        run()
        best = min(best, time.perf_counter() - t)
        reps += 1
        if time.perf_counter() - t_all > budget_s:
            break
    gbs = algorithmic_bytes(kind, srows, n, dtype) / best / 1e9
    unit = "images" if kind == "filter2d" else "rows"
    info = {"value": round(gbs, 3), "unit": "GB/s", "cores": cores, "kind": chk.kind,
            "sample": f"{srows} of {rows} {unit} x {n} ({kind} {dtype}), {note}, best of {reps}",
            "ms": round(best * 1e3, 3), "ms_full_workload": round(best * 1e3 * rows / srows, 3),
            "elements_per_s": round(srows * n * (n if kind == "filter2d" else 1) / best, 1)}
    return gbs, info


def run_reference(args, kind, rows, n, dtype, rank, world):
    if rank != 0:
        return
    steps = max(1, args.steps)
    gbs, info = cpu_reference(kind, rows, n, dtype, budget_s=min(60.0, 2.0 * (steps + args.warmup)))
    line = {"impl": "reference", "metric": f"batched {'fp64' if dtype == 'f64' else 'fp32'} FFT throughput (algorithmic GB/s)", "value": info["value"],
            "unit": "GB/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": info["ms"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": dtype, "data": "synthetic", "config": {"workload": args.workload, "sample": info["sample"]},
            "cpu_baseline": info,
            "e2e": {"value": info["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
_AFFINITY0 = os.sched_getaffinity(0) if hasattr(os, "sched_getaffinity") else None


class all_cpus:
    """The GPU side runs bound to the NUMA node of its device (pinned buffers local to the GPU); the CPU arm and the
    oracle get every host core back for the time they run."""

    def __enter__(self):
        self.prev = os.sched_getaffinity(0) if _AFFINITY0 is not None else None
        if _AFFINITY0 is not None:
            os.sched_setaffinity(0, _AFFINITY0)

    def __exit__(self, *a):
        if self.prev is not None:
            os.sched_setaffinity(0, self.prev)


class Ctx:
    """What every measurement needs: torch, the package, device, rank/world and the barrier."""

    def __init__(self, torch, dist, ib, dev, rank, world):
        self.torch, self.dist, self.ib, self.dev, self.rank, self.world = torch, dist, ib, dev, rank, world

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, v):
        if self.world == 1:
            return float(v)
        t = self.torch.tensor([float(v)], device=self.dev, dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())


class Workload:
    """Inputs, outputs and the step of one named workload.  Working sets smaller than twice the L2 rotate over
    enough buffer sets that no step finds its input or output in cache (timing rule: inputs larger than L2)."""

    def __init__(self, cx: Ctx, name, seed=1234, shard=None):
        torch, ib = cx.torch, cx.ib
        self.cx, self.name = cx, name
        kind, rows, n, dtype = WORKLOADS[name]
        self.kind, self.n, self.dtype, self.rows_total = kind, n, dtype, rows
        if shard is not None:       # (lo, hi) of the rows / images this rank holds: total work fixed
            rows = shard[1] - shard[0]
        self.rows = rows
        self.bytes = algorithmic_bytes(kind, rows, n, dtype)
        rdt = torch.float64 if dtype == "f64" else torch.float32
        cdt = torch.complex128 if dtype == "f64" else torch.complex64
        self.rdt, self.cdt = rdt, cdt
        g = torch.Generator(device=cx.dev).manual_seed(seed + cx.rank)
        nset = 1
        if kind in ("c2c", "r2c", "c2r") and self.bytes < 2.2 * L2_BYTES:
            nset = int(np.ceil(2.2 * L2_BYTES / self.bytes))
        self.nset = nset
        self.sets = []
        self.filt = self.conv = None
        dev = cx.dev
        for _ in range(nset):
            if kind == "fftconv":
                from impulse_b200 import signal as isig
                x = torch.rand((rows, n), generator=g, device=dev, dtype=rdt) - 0.5
                self.taps = torch.rand((FIR_TAPS,), generator=g, device=dev, dtype=rdt) - 0.5
                self.conv = isig
                y = None
            elif kind == "filter2d":
                from impulse_b200.filter import FFTFilter2D
                x = torch.rand((rows, n, n), generator=g, device=dev, dtype=rdt)
                self.ker = torch.rand((31, 31), generator=g, device=dev, dtype=rdt)
                self.ker = self.ker / self.ker.sum()
                self.filt = FFTFilter2D(self.ker, n, n)
                y = torch.empty_like(x)
            elif kind in ("c2c", "fft2"):
                x = torch.view_as_complex(torch.rand((rows, n, 2), generator=g, device=dev, dtype=rdt) - 0.5)
                y = torch.empty_like(x)
            elif kind == "r2c":
                x = torch.rand((rows, n), generator=g, device=dev, dtype=rdt) - 0.5
                y = torch.empty((rows, n // 2 + 1), device=dev, dtype=cdt)
            else:
                x = torch.view_as_complex(torch.rand((rows, n // 2 + 1, 2), generator=g, device=dev, dtype=rdt) - 0.5)
                x[:, 0].imag.zero_()           # a Hermitian half spectrum: bins 0 and N/2 are real
                if n % 2 == 0:
                    x[:, n // 2].imag.zero_()
                y = torch.empty((rows, n), device=dev, dtype=rdt)
            descs = None
            if self.filt is None and self.conv is None:
                descs = (ib.DataDesc.init(x), ib.DataDesc.init(y))
            self.sets.append((x, y, descs))
        self.x, self.y = self.sets[0][0], self.sets[0][1]
        self.fdesc = ib.FFTDesc.init(axes=[0, 1] if kind == "fft2" else [1], forward=(kind != "c2r"),
                                     scalingFactor=(1.0 / n if kind == "c2r" else 1.0))
        self.i = 0

    def step(self):
        x, y, descs = self.sets[self.i % self.nset]
        self.i += 1
        if self.conv is not None:
            self.conv.fftconvolve(x, self.taps, "full")
        elif self.filt is not None:
            self.filt.apply(x, out=y)
        else:
            self.fdesc.apply(descs[1], descs[0])

    def l2_note(self):
        if self.nset > 1:
            return f"rotating over {self.nset} input/output buffer sets ({self.nset * self.bytes / 1e6:.0f} MB > 2x the 126 MB L2)"
        return "input+output per step exceed the 126 MB L2"


def time_steps(cx: Ctx, step, steps, warmup):
    """W untimed steps, then K steps between two CUDA events on the launching stream, barrier + synchronize on both
    sides, max over ranks.  Returns (ms per step, launches in the timed region, this rank's own ms per step)."""
    torch, ib = cx.torch, cx.ib
    for _ in range(warmup):
        step()
    cx.barrier()
    stream = torch.cuda.current_stream()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n0 = ib.launch_count()
    e0.record(stream)
    for _ in range(steps):
        step()
    e1.record(stream)
    cx.barrier()
    launches = ib.launch_count() - n0
    own = e0.elapsed_time(e1) / steps
    return cx.max_over_ranks(own), launches, own


def time_graph(cx: Ctx, step, iters=200, replays=3):
    """>= 200 back-to-back iterations captured in ONE CUDA graph, so that the host-side
    launch cost of a 10-20 us kernel is not what is measured.  Returns ms per iteration (best replay) or None."""
    torch = cx.torch
    try:
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(3):
                step()
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=s):
                for _ in range(iters):
                    step()
            g.replay()
            torch.cuda.synchronize()
            best = float("inf")
            for _ in range(replays):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(s)
                g.replay()
                e1.record(s)
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1) / iters)
        torch.cuda.current_stream().wait_stream(s)
        return best
    except Exception:   # noqa: BLE001 — capture is an extra; the eager number stands
        torch.cuda.synchronize()
        return None


def accuracy_all_rows(cx: Ctx, w: Workload):
    """max over ALL rows of the rel-L2 error against the oracle with its pass bound (the oracle runs
    the whole batch on all host threads), plus the forward->backward round trip on the device.
    The oracle is the checker here, never the thing measured."""
    from oracle import oracle
    torch, ib = cx.torch, cx.ib
    chk = oracle.load()
    kind, n, dtype = w.kind, w.n, w.dtype
    nt = max(1, chk.hardware_threads())
    bound = (1e-12 if dtype == "f64" else 1e-5) * max(1.0, float(np.log2(max(n, 2))))
    x, y = w.x, w.y
    if kind == "filter2d":
        sel = [0, w.rows - 1] if w.rows > 1 else [0]     # two whole images, every pixel
        imgs = x[sel].cpu().numpy()
        kspec = filter_kernel_spectrum_cpu(chk, w.ker.cpu().numpy(), n)
        spec = chk.r2c(imgs, [1, 2], True, 1.0, nthreads=0) * kspec[None]
        want = chk.c2r(np.ascontiguousarray(spec), imgs.shape, [1, 2], False, 1.0 / (n * n), nthreads=0)
        got = y[sel].cpu().numpy()
        err = max(float(oracle.rel_l2(got[i], want[i])) for i in range(len(sel)))
        return {"max_rel_l2_vs_oracle": err, "bound": bound, "checked": f"{len(sel)} whole images ({len(sel) * n * n} pixels)",
                "max_abs_err": float(np.abs(got - want).max()), "oracle": chk.kind, "pass": bool(err <= bound)}
    if kind not in ("c2c", "r2c", "c2r", "fft2"):
        return {"unavailable": f"not reported for {kind}"}
    w.i = 0
    y.fill_(float("nan"))      # a row the kernel never writes stays NaN and fails the comparison
    w.step()
    torch.cuda.synchronize()
    xs, ys = x.cpu().numpy(), y.cpu().numpy()
    if kind == "c2c":
        want = chk.c2c(xs, [1], True, 1.0, nthreads=nt)
    elif kind == "fft2":
        want = chk.c2c(xs, [0, 1], True, 1.0, nthreads=0)
    elif kind == "r2c":
        want = chk.r2c(xs, [1], True, 1.0, nthreads=nt)
    else:
        want = chk.c2r(xs, ys.shape, [1], False, 1.0 / n, nthreads=nt)
    err = float(oracle.max_row_rel_l2(ys, want))
    out = {"max_rel_l2_vs_oracle": err, "bound": bound, "rows_checked": int(ys.shape[0]), "rows_total": int(ys.shape[0]),
           "oracle": chk.kind, "pass": bool(err <= bound)}
    del want, xs, ys
    if kind in ("c2c", "r2c", "fft2"):   # round trip: backward transform of the output with 1/N, against the input
        back = torch.empty_like(x)
        axes = [0, 1] if kind == "fft2" else [1]
        fct = 1.0 / (n * n) if kind == "fft2" else 1.0 / n
        ib.FFTDesc.init(axes=axes, forward=False, scalingFactor=fct).apply(ib.DataDesc.init(back), ib.DataDesc.init(y))
        if kind == "fft2":
            out["round_trip_rel_l2"] = float(torch.linalg.vector_norm(back - x) / torch.linalg.vector_norm(x))
        else:
            num = torch.linalg.vector_norm(back - x, dim=1)
            den = torch.linalg.vector_norm(x, dim=1)
            out["round_trip_max_rel_l2"] = float((num / den).max())
        del back
    return out


def e2e_host(cx: Ctx, w: Workload, e_steps):
    """The public API with pinned HOST buffers: H2D + kernels + D2H inside the timed region, every step."""
    torch, ib = cx.torch, cx.ib
    if w.conv is not None:
        return None
    x, y = w.x, w.y
    hx = torch.empty(x.shape, dtype=x.dtype, pin_memory=True)
    hy = torch.empty(y.shape, dtype=y.dtype, pin_memory=True)
    hx.copy_(x)
    if w.filt is not None:
        # FFTFilter2D takes CUDA tensors: the end-to-end call is copy in -> apply -> copy out, in chunks of images on
        # two streams so that both PCIe directions and the kernels overlap
        nchunk = 8 if w.rows >= 8 else 1
        per = (w.rows + nchunk - 1) // nchunk
        streams = [torch.cuda.Stream(), torch.cuda.Stream()]
        dxs = [torch.empty((per,) + tuple(x.shape[1:]), dtype=x.dtype, device=cx.dev) for _ in range(2)]
        dys = [torch.empty_like(d) for d in dxs]
        filts = [w.filt, type(w.filt)(w.ker, w.n, w.n)]   # one spectrum scratch buffer per stream

        def call():
            for c in range(nchunk):
                lo, hi = c * per, min(w.rows, (c + 1) * per)
                s = streams[c & 1]
                with torch.cuda.stream(s):
                    dxs[c & 1][: hi - lo].copy_(hx[lo:hi], non_blocking=True)
                    filts[c & 1].apply(dxs[c & 1][: hi - lo], out=dys[c & 1][: hi - lo])
                    hy[lo:hi].copy_(dys[c & 1][: hi - lo], non_blocking=True)
            for s in streams:
                s.synchronize()
        how = f"pinned host tensors -> {nchunk} chunks on two streams: H2D, FFTFilter2D.apply, D2H"
    else:
        hin, hout = ib.DataDesc.init(hx.numpy()), ib.DataDesc.init(hy.numpy())
        call = lambda: w.fdesc.apply(hout, hin)   # noqa: E731 — synchronous for host pointers: H2D + kernel + D2H
        how = "FFTDesc.apply on pinned numpy arrays (C ABI with host pointers)"
    for _ in range(2):
        call()
    cx.barrier()
    t0 = time.perf_counter()
    for _ in range(e_steps):
        call()
    torch.cuda.synchronize()
    dt = cx.max_over_ranks((time.perf_counter() - t0) / e_steps)
    hyn = hy.numpy()
    chk = float(np.abs(hyn.reshape(-1)[:4]).sum())  # touch the result on the host
    return {"value": round(cx.world * w.bytes / dt / 1e9, 3), "unit": "GB/s",
            "h2d_bytes_per_step": int(hx.numel() * hx.element_size()), "d2h_bytes_per_step": int(hy.numel() * hy.element_size()),
            "ms_per_step": round(dt * 1e3, 3), "steps": e_steps, "host_memory": "pinned", "how": how, "check": chk}


def config_entry(cx: Ctx, name, steps, warmup, peak, with_cpu=True, with_e2e=True):
    """One BASELINE config, measured like the main line (device-resident timing, roofline, every-row accuracy, e2e,
    CPU arm) and returned as an object for the `configs` block."""
    torch, ib = cx.torch, cx.ib
    w = Workload(cx, name)
    ms, launches, _ = time_steps(cx, w.step, steps, warmup)
    gbs = w.bytes / (ms * 1e-3) / 1e9
    per_step = max(1, round(launches / steps))
    ent = {"workload": name, "ms_per_step": round(ms, 5), "GB/s": round(gbs, 1), "frac_measured_peak": round(gbs / peak, 4),
           "frac_8TBps": round(gbs / 8000.0, 4), "launches_per_step": per_step,
           "kernel": ib.last_kernel() + (f" (last of {per_step})" if per_step > 1 else ""),
           "algorithmic_bytes": int(w.bytes), "l2": w.l2_note(),
           "elements_per_s": round(w.rows * w.n * (w.n if w.kind == "filter2d" else 1) / (ms * 1e-3), 1)}
    if w.kind == "fft2":
        ent["GB/s_one_read_one_write_ideal"] = round(gbs / 2, 1)
    if ms < 0.1 and per_step == 1:
        gms = time_graph(cx, w.step)
        if gms:
            ent["cuda_graph_200_iters"] = {"ms_per_step": round(gms, 5), "GB/s": round(w.bytes / (gms * 1e-3) / 1e9, 1),
                                           "frac_8TBps": round(w.bytes / (gms * 1e-3) / 1e9 / 8000.0, 4)}
    if cx.rank == 0:
        try:
            with all_cpus():
                ent["accuracy"] = accuracy_all_rows(cx, w)
        except Exception as ex:  # noqa: BLE001
            ent["accuracy"] = {"unavailable": repr(ex)}
    if with_e2e:
        try:
            ent["e2e"] = e2e_host(cx, w, 3)
        except Exception as ex:  # noqa: BLE001
            ent["e2e"] = {"unavailable": repr(ex)}
    if with_cpu and cx.rank == 0:
        try:
            with all_cpus():
                _, info = cpu_reference(w.kind, w.rows_total, w.n, w.dtype, budget_s=2.5)
            ent["cpu_baseline"] = info
            if isinstance(ent.get("e2e"), dict) and "value" in ent["e2e"]:
                ent["e2e_vs_cpu"] = round(ent["e2e"]["value"] / info["value"], 3)
        except Exception as ex:  # noqa: BLE001
            ent["cpu_baseline"] = {"unavailable": repr(ex)}
    del w
    torch.cuda.empty_cache()
    return ent


def slab_entries(cx: Ctx, steps, warmup):
    """BASELINE config 4 at N > 1: ONE 8192 x 8192 transform split by row slabs (total work fixed: strong scaling), both
    exchange variants, with the assembled result compared with the oracle on rank 0 and the single-GPU time of the
    same transform measured in the same run."""
    torch, dist, ib = cx.torch, cx.dist, cx.ib
    from impulse_b200 import dist as idist
    kind, rows, n, dtype = WORKLOADS["fft2_8192x8192_c128"]
    world, rank = cx.world, cx.rank
    lo, hi = idist.shard_rows(rows, rank, world)
    g = torch.Generator(device=cx.dev).manual_seed(4321 + rank)
    x = torch.view_as_complex(torch.rand((hi - lo, n, 2), generator=g, device=cx.dev, dtype=torch.float64) - 0.5)
    total_bytes = algorithmic_bytes(kind, rows, n, dtype)
    out = {}
    # single-GPU time of the whole transform (every rank runs it on its own device; max over ranks)
    full = Workload(cx, "fft2_8192x8192_c128")
    ms1, _, _ = time_steps(cx, full.step, max(3, steps // 2), 3)
    del full
    torch.cuda.empty_cache()
    out["single_gpu_ms"] = round(ms1, 5)
    # the oracle's answer, once, on rank 0: gather the row slabs
    want = None
    xr = torch.view_as_real(x)           # NCCL has no complex types: gather the (re, im) view
    xs = [torch.empty_like(xr) for _ in range(world)] if rank == 0 else None
    dist.gather(xr, xs, dst=0)
    if rank == 0:
        from oracle import oracle
        chk = oracle.load()
        xfull = torch.view_as_complex(torch.cat(xs, dim=0)).cpu().numpy()
        with all_cpus():
            want = chk.c2c(xfull, [0, 1], True, 1.0, nthreads=0)
        del xfull
    del xs
    op = idist.SlabFFT2P2P(hi - lo, n, torch.complex128, pull_chunks=0)
    # (2 chunks / 64 copy CTAs measured best on 2 GPUs, profiles/r02_ab_slab_pull_n2.txt — and still behind the fused peer loads)
    op_pull = idist.SlabFFT2P2P(hi - lo, n, torch.complex128, pull_chunks=int(os.environ.get("IMPULSE_FFT_SLAB_PULL_BENCH", "2")),
                                copy_ctas=int(os.environ.get("IMPULSE_FFT_SLAB_COPY_CTAS", "64")))
    variants = (("fft2_slab_p2p", lambda: op(x, True, 1.0), "row FFTs -> 1-element all-reduce as barrier -> column kernels load "
                 "the peers' row slabs over NVLink (CUDA IPC); no pack, no all-to-all buffer"),
                ("fft2_slab_pull", lambda: op_pull(x, True, 1.0), f"row FFTs -> barrier -> {op_pull.pull_chunks} column chunks: a small gather "
                 "kernel pulls chunk j+1 out of the peers' row slabs over NVLink (side stream) while the column transform of chunk j "
                 "runs on local memory"),
                ("fft2_slab_nccl", lambda: idist.fft2_slab(x, True, 1.0), "row FFTs -> pack -> NCCL all_to_all_single -> column FFTs"))
    for key, fn, how in variants:
        ms, launches, _ = time_steps(cx, fn, steps, warmup)
        cols = fn()                                   # [rows, n / world]: this rank's column slab of the result
        torch.cuda.synchronize()
        cr = torch.view_as_real(cols.contiguous())
        parts = [torch.empty_like(cr) for _ in range(world)] if rank == 0 else None
        dist.gather(cr, parts, dst=0)
        ent = {"ms_per_step": round(ms, 5), "GB/s": round(total_bytes / (ms * 1e-3) / 1e9, 1), "scaling": "strong",
               "speedup_vs_single_gpu": round(ms1 / ms, 3), "launches_per_step_per_rank": round(launches / steps, 2),
               "result_layout": "column slabs", "how": how}
        if rank == 0:
            from oracle import oracle
            got = torch.view_as_complex(torch.cat(parts, dim=1)).cpu().numpy()
            err = float(oracle.rel_l2(got, want))
            rowerr = float(oracle.max_row_rel_l2(got, want))
            bound = 1e-12 * 13
            ent["accuracy"] = {"rel_l2_vs_oracle": err, "max_row_rel_l2_vs_oracle": rowerr, "bound": bound,
                               "elements_checked": int(got.size), "oracle": "reference", "pass": bool(rowerr <= bound)}
            del got
        del parts, cols
        out[key] = ent
    op.close()
    op_pull.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the `configs` block (the other BASELINE configs)")
    ap.add_argument("--configs", action="store_true", help="add the `configs` block even for a non-default --workload")
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
    try:   # host buffers (pinned staging, e2e) on the NUMA node of this rank's GPU: 8 ranks must not share node 0
        numa = ib.bind_host_to_device(local)
    except Exception:  # noqa: BLE001
        numa = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG", "WARN")  # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=dev)
    cx = Ctx(torch, dist, ib, dev, rank, world)
    warmup = max(3, args.warmup)
    steps = max(1, args.steps)
    peak, peak_src = peaks()

    slab = kind == "fft2" and world > 1
    filt_shard = kind == "filter2d" and world > 1
    if slab:
        # ONE 2-D transform split by row slabs: total work fixed ("strong"), exchange fused into the column kernels
        from impulse_b200 import dist as idist
        lo, hi = idist.shard_rows(rows, rank, world)
        g = torch.Generator(device=dev).manual_seed(1234 + rank)
        x = torch.view_as_complex(torch.rand((hi - lo, n, 2), generator=g, device=dev, dtype=torch.float64) - 0.5)
        slab_op = None if os.environ.get("IMPULSE_FFT_SLAB", "p2p") == "nccl" else idist.SlabFFT2P2P(hi - lo, n, torch.complex128)
        step = (lambda: slab_op(x, True, 1.0)) if slab_op is not None else (lambda: idist.fft2_slab(x, True, 1.0))
        w = None
        bytes_per_gpu = algorithmic_bytes(kind, rows, n, dtype)
        l2_note = "input+output per step exceed the 126 MB L2"
    else:
        shard = None
        if filt_shard:
            from impulse_b200 import dist as idist
            shard = idist.shard_rows(rows, rank, world)
        w = Workload(cx, args.workload, shard=shard)
        step = w.step
        bytes_per_gpu = algorithmic_bytes(kind, rows, n, dtype)
        l2_note = w.l2_note()

    sampler = ClockSampler(local) if rank == 0 else None
    t_host0 = time.time()
    ms_per_step, launches, own_ms = time_steps(cx, step, steps, warmup)
    t_host1 = time.time()
    strong = slab or filt_shard   # total work fixed, split over ranks
    value = (1 if strong else world) * bytes_per_gpu / (ms_per_step * 1e-3) / 1e9
    last_kernel = ib.last_kernel()

    # ---- e2e: public API with pinned HOST buffers, copies inside the timed region
    e2e = None
    if not args.no_e2e and w is not None and w.conv is None:
        e2e = e2e_host(cx, w, max(2, min(steps, 5)))
    clocks = sampler.stop(t_host0, t_host1) if sampler else None

    line = None
    if rank == 0:
        # dominant (only) kernel of the step: one launch per step on this stream
        # (multi-launch steps — fft2, filter2d, fftconvolve: the step's algorithmic bytes over the whole step)
        per_step = max(1, round(launches / steps))
        k_ms = own_ms
        per_rank_bytes = bytes_per_gpu / (world if strong else 1)
        achieved = per_rank_bytes / (k_ms * 1e-3) / 1e9
        line = {
            "metric": "batched fp64 FFT throughput (algorithmic GB/s)" if dtype == "f64" else "batched fp32 FFT throughput (algorithmic GB/s)",
            "value": round(value, 2), "unit": "GB/s", "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": round(ms_per_step, 5), "higher_is_better": True, "scaling": "strong" if strong else "weak",
            "vs_baseline": None, "dtype": dtype, "data": "synthetic",
            "config": {"workload": args.workload, "rows_per_gpu": (rows // world) if strong else rows, "length": n, "kind": kind,
                       "placement": "out of place, device resident", "l2": l2_note,
                       "parallelism": (f"row-slab x{world}, " + ("column kernels load peers' row slabs over NVLink (CUDA IPC), one 1-element all-reduce as barrier"
                                                                if os.environ.get("IMPULSE_FFT_SLAB", "p2p") != "nccl" else "pack + NCCL all-to-all") + ", result left in column slabs" if slab
                                       else f"batch-shard x{world}, no collective"),
                       "elements_per_s": round((1 if strong else world) * rows * n * (n if kind == "filter2d" else 1) / (ms_per_step * 1e-3), 1),
                       "frac_of_8TBps_nominal": round(value / world / 8000.0, 4), "host_numa_node": numa},
            "roofline": {"bound": "hbm", "achieved": round(achieved, 2), "peak": peak, "unit": "GB/s",
                         "frac": round(achieved / peak, 4), "traffic": committed_traffic(args.workload, last_kernel),
                         "peak_source": peak_src, "kernel": last_kernel + (f" (last of {per_step} launches per step)" if per_step > 1 else ""),
                         "algorithmic_bytes_per_launch": int(per_rank_bytes)},
            "gpu_launches": int(launches), "clocks": clocks,
        }
        if e2e:
            line["e2e"] = e2e
        if not args.no_cpu and w is not None and w.conv is None:
            if world == 1:
                try:
                    with all_cpus():
                        _, info = cpu_reference(kind, rows, n, dtype)
                    line["cpu_baseline"] = info
                except Exception as ex:  # the oracle is optional for the measurement itself
                    line["cpu_baseline"] = {"unavailable": str(ex)}
            try:
                with all_cpus():
                    line["accuracy"] = accuracy_all_rows(cx, w)
            except Exception as ex:  # noqa: BLE001 — a reporting extra never costs the bench line
                line["accuracy"] = {"unavailable": repr(ex)}
    del w
    torch.cuda.empty_cache()

    # ---- every other BASELINE config, same measurements (all ranks run them: timing is max over ranks)
    want_configs = (args.workload == DEFAULT_WORKLOAD and not args.no_configs) or args.configs
    if want_configs:
        configs = {}
        c_steps = max(5, min(steps, 20))
        for key, name in CONFIG_BLOCK:
            try:
                ent = config_entry(cx, name, c_steps, 3, peak, with_cpu=(world == 1 and not args.no_cpu),
                                   with_e2e=(world == 1 and not args.no_e2e))
            except Exception as ex:  # noqa: BLE001
                ent = {"unavailable": repr(ex)}
                torch.cuda.synchronize()
            if world > 1 and isinstance(ent, dict) and "GB/s" in ent:
                ent["note"] = f"each of the {world} ranks runs the full config (batch replicas, no collective); ms = max over ranks"
            configs[key] = ent
        if world > 1:
            # config 2, strong: the 65536 rows split over the ranks (what BASELINE config 2 calls "sharded by batch")
            try:
                from impulse_b200 import dist as idist
                sw = Workload(cx, DEFAULT_WORKLOAD, shard=idist.shard_rows(rows, rank, world))
                ms, _, _ = time_steps(cx, sw.step, steps, warmup)
                tot = algorithmic_bytes(*WORKLOADS[DEFAULT_WORKLOAD])
                configs["2_c2c_65536x1024_strong"] = {
                    "ms_per_step": round(ms, 5), "GB/s": round(tot / (ms * 1e-3) / 1e9, 1), "scaling": "strong",
                    "rows_per_gpu": sw.rows, "speedup_vs_weak_single_gpu_step": round(ms_per_step / ms, 3), "l2": sw.l2_note()}
                del sw
                torch.cuda.empty_cache()
            except Exception as ex:  # noqa: BLE001
                configs["2_c2c_65536x1024_strong"] = {"unavailable": repr(ex)}
            try:
                configs["4_fft2_8192x8192_slab"] = slab_entries(cx, c_steps, 3)
            except Exception as ex:  # noqa: BLE001
                configs["4_fft2_8192x8192_slab"] = {"unavailable": repr(ex)}
        if rank == 0:
            line["configs"] = configs
            acc = [v.get("accuracy", {}).get("pass") for v in configs.values() if isinstance(v, dict) and "accuracy" in v]
            for v in configs.values():
                if isinstance(v, dict):
                    acc += [s["accuracy"].get("pass") for s in v.values() if isinstance(s, dict) and "accuracy" in s]
            line["configs_accuracy_all_pass"] = bool(acc) and all(a is True for a in acc)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
