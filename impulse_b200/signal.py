"""FFT-backed signal primitives: the steps either side of the FFT path in `impulse/signal.nim`
(SURVEY 8(f) rank 3).

The reference computes `upfirdn` as zero-insertion upsampling followed by arraymancer's DIRECT
`convolve(mode = full, down = down)` (impulse/signal.nim:594; arraymancer >= 0.7.30 is an external
dependency that is not vendored in the reference).  Here the convolution runs on the GPU in the
frequency domain — real rows: r2c -> pointwise multiply -> c2r; complex rows: `impulse_fft_convolve_axis`
— and `fftconvolve` is exposed in its own right, batched over rows.

Host-side pieces restated from the reference (filter design is O(order) scalar work, it stays on the
host like planning does):
  kaiser                      impulse/signal.nim:135-166   (I0 via numpy's Clenshaw/Cephes `i0`, :72-133)
  firls                       impulse/signal.nim:229-543
  reduce_resampling_rates     impulse/signal.nim:596-611
  generate_resampling_filter  impulse/signal.nim:613-649
  adjust_filter_position      impulse/signal.nim:651-688
  upfirdn                     impulse/signal.nim:545-594
  resample                    impulse/signal.nim:690-744 (explicit filter), :746-790 (designed filter)

The arithmetic goes through an *engine* (`CudaConvEngine`, the C ABI); there is no CPU fallback — the
engine argument exists so that the host logic can be exercised on CPU by tests with their own checker.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from . import _lib


# ---- filter design (host) ---------------------------------------------------------------------------
def kaiser(size: int = 10, beta: float = 5.0) -> np.ndarray:
    """Kaiser window, maximum normalised to 1 (impulse/signal.nim:135-166)."""
    if size < 0:
        raise ValueError("The size of the Kaiser window must be non-negative")
    if size == 1:
        return np.ones(1)
    n = np.arange(size, dtype=np.float64)
    alpha = (size - 1.0) / 2.0
    return np.i0(beta * np.sqrt(1.0 - ((n - alpha) / alpha) ** 2)) / np.i0(beta)


def firls(fir_order: int, bands, desired, weights=None, symmetric: bool = True, fs: float = 2.0) -> np.ndarray:
    """Least-squares linear-phase FIR design (impulse/signal.nim:229-543): minimises the weighted integral of
    the squared error between the piecewise-linear `desired` gains on `bands` and the filter's response.
    Types I/II (symmetric) and III/IV (anti-symmetric); the normal equations are only solved when the
    weights differ or the bands leave gaps, otherwise the coefficients are the closed-form integrals."""
    if fs <= 0:
        raise ValueError(f"Sampling frequency fs must be positive but got {fs}")
    f = np.asarray(bands, dtype=np.float64).reshape(-1).copy()
    d = np.asarray(desired, dtype=np.float64).reshape(-1)
    if f.max() > fs / 2.0 or f.min() < 0.0:
        raise ValueError("Some frequency values are outside of the valid range")
    if f.size % 2:
        raise ValueError("Frequency band tensor length is not even")
    if f.size != d.size:
        raise ValueError("Frequency and desired gain tensors must have the same length")
    even = fir_order % 2 == 0
    if not even and f[-1] == 1.0 and d[-1] != 0.0:
        raise ValueError("Filter order must be even when the last frequency is 1.0 and its desired gain is not 0.0")
    w = np.ones(f.size // 2) if weights is None or len(weights) == 0 else np.asarray(weights, dtype=np.float64)
    if f.size != 2 * w.size:
        raise ValueError("Weight tensor length must be half the length of the frequency band tensor")
    df = np.diff(f)
    if np.any(df < 0):
        raise ValueError("Frequency band tensor values must increasing monotonically")
    f /= fs
    solve = not (np.all(w == w[0]) and (df.size == 1 or np.all(df[1::2] == 0.0)))
    half = fir_order / 2.0
    if symmetric:
        k = np.arange(math.floor(half) + 1.0)
        if not even:
            k = k + 0.5
    else:
        k = np.arange(1.0, math.floor(half) + 1.0) if even else np.arange(math.floor(half) + 1.0) + 0.5
    if solve:
        sp = 2.0 * (k[:, None] + k[None, :])
        sn = 2.0 * (k[:, None] - k[None, :])
        q = np.zeros_like(sp)
    if symmetric and even:
        k = k[1:]
    tb = 2.0 * np.pi * k
    b0 = 0.0
    b = np.zeros_like(k)
    sgn = 1.0 if symmetric else -1.0
    for i in range(0, f.size, 2):
        wp = abs(w[(i + 1) // 2])
        f0, f1 = f[i], f[i + 1]
        if solve:
            q += wp * 0.5 * f1 * (np.sinc(sp * f1) + sgn * np.sinc(sn * f1)) - \
                 wp * 0.5 * f0 * (np.sinc(sp * f0) + sgn * np.sinc(sn * f0))
        m = (d[i + 1] - d[i]) / (f1 - f0)
        c = d[i] - m * f0
        if symmetric:
            if even:
                b0 += wp * (c * (f1 - f0) + m / 2.0 * (f1 * f1 - f0 * f0))
            b += wp * (m / (4.0 * np.pi ** 2) * (np.cos(tb * f1) - np.cos(tb * f0)) / (k * k))
            b += wp * f1 * (m * f1 + c) * np.sinc(2.0 * k * f1) - wp * f0 * (m * f0 + c) * np.sinc(2.0 * k * f0)
        else:
            b += wp * (m / (4.0 * np.pi ** 2) * (np.sin(tb * f1) - np.sin(tb * f0)) / (k * k))
            b += (wp * (m * f0 + c) * np.cos(tb * f0) - wp * (m * f1 + c) * np.cos(tb * f1)) / tb
    if symmetric:
        if even:
            b = np.concatenate([[b0], b])
        if solve:
            a = np.linalg.solve(q, b)
        else:
            a = 4.0 * w[0] * b
            if even:
                a[0] /= 2.0
        if even:
            h = int(half)
            return np.concatenate([a[h:0:-1] * 0.5, [a[0]], a[1:h + 1] * 0.5])
        return 0.5 * np.concatenate([a[::-1], a])
    a = np.linalg.solve(q, b) if solve else -4.0 * w[0] * b
    if even:
        return 0.5 * np.concatenate([a[::-1], [0.0], -a])
    return 0.5 * np.concatenate([a[::-1], -a])


def reduce_resampling_rates(up: int, down: int):
    """impulse/signal.nim:596-611."""
    g = math.gcd(up, down)
    return up // g, down // g


def generate_resampling_filter(up: int, down: int, fir_order_factor: int = 10, beta: float = 5.0) -> np.ndarray:
    """Kaiser-windowed least-squares low-pass at 1/max(up, down), gain `up` (impulse/signal.nim:613-649)."""
    if fir_order_factor == 0:
        return np.ones(up)
    r = max(up, down)
    fc = 1.0 / r
    order = 2 * fir_order_factor * r
    h = firls(order, [0.0, fc, fc, 1.0], [1.0, 1.0, 0.0, 0.0])
    h = h * kaiser(order + 1, beta)
    return h * (up / h.sum())


def adjust_filter_position(h: np.ndarray, input_len: int, result_len: int, up: int, down: int):
    """Leading zeros so that decimation samples the filter's centre, trailing zeros so that enough output
    exists; returns the padded filter and the post-decimation delay (impulse/signal.nim:651-688)."""
    mid = (len(h) - 1.0) / 2.0
    lead = int(math.floor(down - (mid % down)))
    h = np.concatenate([np.zeros(lead, dtype=h.dtype), h])
    mid += lead
    delay = int(math.floor(math.ceil(mid) / down))
    filtered_len = (input_len - 1) * up + len(h)
    trail = 0
    while delay + result_len >= math.ceil((filtered_len + trail) / down):
        trail += 1
    return np.concatenate([h, np.zeros(trail, dtype=h.dtype)]), delay


# ---- convolution engine -----------------------------------------------------------------------------
def next_fast_len(n: int, even: bool = False) -> int:
    """Smallest 2^a 3^b 5^c >= n: every factor has an in-line butterfly in the device engine.  `even=True`
    (real transforms: an even length runs as a half-length complex transform) doubles the fast length of n/2."""
    if even:
        return 2 * next_fast_len((n + 1) // 2)
    best = 1 << max(0, (n - 1).bit_length())
    p5 = 1
    while p5 < best:
        p35 = p5
        while p35 < best:
            q = -(-n // p35)
            p2 = 1 << max(0, (q - 1).bit_length())
            cand = p35 * p2
            if n <= cand < best:
                best = cand
            p35 *= 3
        p5 *= 5
    return best


# transform lengths served by the register kernels (fast_kernels.cu): real rows run as a half-length complex
# transform, so the real lengths are twice the complex ones
_FAST_REAL = {"f64": (4096, 8192, 16384), "f32": (4096, 8192)}
_FAST_CPLX = {"f64": (1024, 2048, 4096, 8192), "f32": (1024, 2048, 4096)}
_ONE_SHOT_MAX = 16384      # longest single-launch transform worth using when no register kernel fits


def plan_blocks(n: int, m: int, real: bool, prec: str):
    """Overlap-save blocking of an n-sample row convolved with m taps: returns (P, V, nb) — nb blocks of P
    samples advancing by V = P - (m - 1), each yielding V valid outputs.  Picks the transform length that
    moves the least data, counting lengths without a register kernel 2.5x."""
    out_len = n + m - 1
    best = None
    fast = (_FAST_REAL if real else _FAST_CPLX)[prec]
    one = next_fast_len(out_len + m - 1, even=real)       # a single block: V = P - (m-1) >= out_len
    cands = [(one, 1.0 if one in fast else 2.5)] if one <= _ONE_SHOT_MAX else []
    cands += [(p, 1.0) for p in fast]
    for p, penalty in cands:
        v = p - (m - 1)
        if v < 1:
            continue
        nb = -(-out_len // v)
        cost = nb * p * penalty
        if best is None or cost < best[0]:
            best = (cost, p, v, nb)
    if best is None:
        raise _lib.FFTError(-3, f"a filter of {m} taps needs a transform longer than this engine runs in one launch")
    return best[1], best[2], best[3]


class CudaConvEngine:
    """Full linear convolution of the rows of `x` ([B, n], CUDA tensor) with one filter `h` ([m]) through
    libimpulse_fft_b200, by overlap-save: the padded rows are viewed as overlapping blocks of P samples (a
    strided view, nothing is gathered), each block is transformed, multiplied by the filter's spectrum and
    transformed back (three launches for the whole batch), and the valid V = P - (m-1) samples of every
    block are the output.  Short rows are the one-block case."""

    def full(self, x, h):
        import torch
        from .desc import DataDesc, FFTDesc
        L = _lib.lib()
        b, n = x.shape
        real = not x.is_complex()
        true_len = n + h.shape[0] - 1
        if real and h.shape[0] % 2 == 0:       # an odd tap count keeps the block stride even (paired real loads)
            h = torch.cat([h, torch.zeros(1, dtype=h.dtype, device=h.device)])
        m = h.shape[0]
        prec = "f64" if x.dtype in (torch.float64, torch.complex128) else "f32"
        code = _lib.F64 if prec == "f64" else _lib.F32
        p, v, nb = plan_blocks(n, m, real, prec)
        stream = C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
        # rows laid end to end, each as [m-1 zeros | signal | zeros up to nb*v]; the zeros that the last block of
        # a row reads past its end are the leading zeros of the next row (one more run closes the buffer)
        flat = torch.zeros(b * nb * v + (m - 1) + p, dtype=x.dtype, device=x.device)
        flat[:b * nb * v].view(b, nb * v)[:, m - 1:m - 1 + n] = x
        blocks = flat.as_strided((b * nb, p), (v, 1))
        cdt = x.dtype if not real else (torch.complex128 if prec == "f64" else torch.complex64)
        pc = p // 2 + 1 if real else p
        hp = torch.zeros((p,), dtype=x.dtype, device=x.device)
        hp[:m] = h
        hf = torch.empty((pc,), dtype=cdt, device=x.device)
        FFTDesc.init(axes=[0], forward=True).apply(DataDesc.init(hf), DataDesc.init(hp))
        spec = torch.empty((b * nb, pc), dtype=cdt, device=x.device)
        FFTDesc.init(axes=[1], forward=True).apply(DataDesc.init(spec), DataDesc.init(blocks))
        _lib.check(L.impulse_fft_cmul(code, spec.data_ptr(), hf.data_ptr(), spec.data_ptr(), pc, b * nb, 1.0, stream))
        res = torch.empty((b * nb, p), dtype=x.dtype, device=x.device)
        FFTDesc.init(axes=[1], forward=False, scalingFactor=1.0 / p).apply(DataDesc.init(res), DataDesc.init(spec))
        return res[:, m - 1:].reshape(b, nb * v)[:, :true_len]


def _to_device(a, engine=None):
    """numpy / torch (any device) -> CUDA tensor; returns (tensor, restore) where restore maps a CUDA result
    back to the caller's kind of array.  An engine that declares `host = True` (tests) keeps CPU tensors."""
    import torch
    if getattr(engine, "host", False):
        if isinstance(a, np.ndarray):
            return torch.from_numpy(np.ascontiguousarray(a)), (lambda r: r.numpy())
        return a, (lambda r: r)
    if isinstance(a, np.ndarray):
        if not torch.cuda.is_available():
            raise _lib.FFTError(-6, "no usable CUDA device (there is no CPU fallback)")
        return torch.from_numpy(np.ascontiguousarray(a)).cuda(), (lambda r: r.cpu().numpy())
    if a.is_cuda:
        return a, (lambda r: r)
    if not torch.cuda.is_available():
        raise _lib.FFTError(-6, "no usable CUDA device (there is no CPU fallback)")
    return a.cuda(), (lambda r: r.cpu())


def _common(t, h):
    import torch
    if t.is_complex() or h.is_complex():
        dt = torch.complex128 if torch.float64 in (t.real.dtype if t.is_complex() else t.dtype,
                                                   h.real.dtype if h.is_complex() else h.dtype) else torch.complex64
    else:
        if not (t.is_floating_point() and h.is_floating_point()):
            t = t.to(torch.float64) if not t.is_floating_point() else t
            h = h.to(torch.float64) if not h.is_floating_point() else h
        dt = torch.float64 if torch.float64 in (t.dtype, h.dtype) else torch.float32
    return t.to(dt), h.to(dt)


def fftconvolve(t, h, mode: str = "full", engine=None):
    """Linear convolution of `t` (rank 1, or rank 2 = independent rows) with the rank-1 filter `h`, computed
    in the frequency domain on the GPU.  mode: "full" (n+m-1 samples, arraymancer's ConvolveMode.full),
    "same" (centred, n samples) or "valid" (n-m+1)."""
    engine = engine or CudaConvEngine()
    td, restore = _to_device(t, engine)
    hd, _ = _to_device(h, engine)
    td, hd = _common(td, hd)
    if hd.ndim != 1 or td.ndim not in (1, 2):
        raise ValueError("fftconvolve takes a rank-1 or rank-2 signal and a rank-1 filter")
    rows = td if td.ndim == 2 else td[None, :]
    n, m = rows.shape[1], hd.shape[0]
    if n == 0 or m == 0:
        raise ValueError("empty input")
    full = engine.full(rows.contiguous(), hd.contiguous())
    if mode == "full":
        res = full
    elif mode == "same":
        lo = (m - 1) // 2
        res = full[:, lo:lo + n]
    elif mode == "valid":
        if m > n:
            raise ValueError("the filter is longer than the signal")
        res = full[:, m - 1:n]
    else:
        raise ValueError("mode must be 'full', 'same' or 'valid'")
    res = res if td.ndim == 2 else res[0]
    return restore(res.contiguous())


def upfirdn(t, h, up: int = 1, down: int = 1, engine=None):
    """Upsample by zero insertion, FIR-filter, downsample (impulse/signal.nim:545-594; Matlab argument order:
    signal first).  Rank-1 signal, or rank 2 for independent rows."""
    if up < 1 or down < 1:
        raise ValueError("up and down must be positive")
    engine = engine or CudaConvEngine()
    import torch
    td, restore = _to_device(t, engine)
    hd, _ = _to_device(h, engine)
    integer = not (td.is_floating_point() or td.is_complex() or hd.is_floating_point() or hd.is_complex())
    int_dtype = td.dtype
    td, hd = _common(td, hd)
    rows = td if td.ndim == 2 else td[None, :]
    b, n = rows.shape
    if up > 1:                                   # [1, 2, 3] -> [1, 0, 0, 2, 0, 0, 3]
        ups = torch.zeros((b, (n - 1) * up + 1), dtype=rows.dtype, device=rows.device)
        ups[:, ::up] = rows
    else:
        ups = rows.contiguous()
    full = engine.full(ups, hd.contiguous())
    res = full[:, ::down]
    if integer:
        res = torch.round(res).to(int_dtype)
    res = res if td.ndim == 2 else res[0]
    return restore(res.contiguous())


def resample(t, h=None, up: int = 1, down: int = 1, fir_order_factor: int = 10, beta: float = 5.0, engine=None):
    """Resample at `up / down` times the sampling rate, keeping the ceil(len * up / down) samples aligned with
    the input (impulse/signal.nim:690-790).  With `h` the rates are used as given; without it they are
    reduced first and an anti-aliasing filter is designed (generate_resampling_filter)."""
    n = t.shape[-1]
    if h is None:
        if up == down:
            return t.copy() if isinstance(t, np.ndarray) else t.clone()
        up, down = reduce_resampling_rates(up, down)
        h = generate_resampling_filter(up, down, fir_order_factor, beta)
    elif up == 1 and down == 1:
        return t.copy() if isinstance(t, np.ndarray) else t.clone()
    hh = np.asarray(h.cpu() if hasattr(h, "cpu") else h).squeeze()
    if hh.ndim != 1:
        raise ValueError(f"Squeezed filter rank ({hh.ndim}) must be 1")
    result_len = int(math.ceil(n * up / down))
    if str(t.dtype) in ("float32", "complex64", "torch.float32", "torch.complex64"):
        hh = hh.astype(np.float32)   # the result has the precision of the signal
    hh, delay = adjust_filter_position(hh, n, result_len, up, down)
    res = upfirdn(t, hh, up=up, down=down, engine=engine)
    return res[..., delay:delay + result_len]
