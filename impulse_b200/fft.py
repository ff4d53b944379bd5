"""Mirror of the reference's C-backend FFT API (fft / ifft / rfft / rfft_packed / unpackFFT /
symmetrize / Normalize) from /root/reference/impulse/fft/c_pocketfft/pocketfft.nim (NC:) and
pocketfft_arraymancer.nim (NA:), running on the B200 engine through the C ABI.

Containers: numpy arrays stand in for array/seq/Tensor held in host memory (staged through the
GPU by the library); torch CUDA tensors are device-resident Tensors.  As an extension
 every function transforms along the LAST axis of an N-D input, batching the
leading axes, where the reference flattens (NA:37).

Kept from the reference: names, argument order and defaults, float64-only precision on this API,
the FFTPACK halfcomplex packing (NC:228-238), the default nkBackward scaling (NC:199-204) and the
data-dependent length guess of symmetrize on complex input (NC:173-180).
Deliberate deviation: `normalize`/`normValue` are honoured on every overload, as
the Tensor overloads do (NA:37,46,58) and tests/test_fft2.nim:13-14 require; the array/seq
overloads of the reference silently drop them (NC:305-310).
"""
from __future__ import annotations

import ctypes as C
import math
import collections
import threading

import numpy as np

from . import _buffers as B
from . import _lib

# NormalizeKind (NC:189-197)
nkBackward, nkOrtho, nkForward, nkCustom = "nkBackward", "nkOrtho", "nkForward", "nkCustom"


def isOdd(i: int) -> bool:  # NC:116
    return (i & 1) == 1


def initNormalize(kind: str, forward: bool, value: float, length: int) -> float:
    """NC:199-204 — the factor handed to the engine as `fct`."""
    if kind == nkBackward:
        return 1.0 if forward else 1.0 / float(length)
    if kind == nkForward:
        return 1.0 / float(length) if forward else 1.0
    if kind == nkOrtho:
        return 1.0 / math.sqrt(float(length))
    if kind == nkCustom:
        return float(value)
    raise ValueError(f"unknown NormalizeKind {kind!r}")


def _fct(kind: str, forward: bool, value: float, length: int) -> float:
    """Scale factor actually applied by the C backend: its length-1 plans return before scaling
    (pocketfft.c:875 pass_all, pocketfft.c:1704 rfftp_forward), so fct is ignored for N == 1."""
    return 1.0 if length == 1 else initNormalize(kind, forward, value, length)


# ---- plan cache (the reference rebuilds a plan on every call, NC:285,300; on a GPU the tables
# ---- must be cached) ---------------------------------------------------------
class _Plan:
    """A cached plan handle with a use count: a plan evicted while another thread is inside impulse_fft_execute is
    destroyed by that thread when it returns, never under it (include/impulse_fft_b200.h allows concurrent executes)."""
    __slots__ = ("h", "users", "evicted")

    def __init__(self, h):
        self.h, self.users, self.evicted = h, 0, False


_PLAN_CAP = 256
_plans: "collections.OrderedDict" = collections.OrderedDict()   # least recently used first
_plans_lock = threading.Lock()


def _acquire_plan(kind, layout, code, shape, sin, sout, axes, forward) -> _Plan:
    key = (kind, layout, code, tuple(shape), tuple(sin), tuple(sout), tuple(axes), bool(forward), _device_key())
    with _plans_lock:
        p = _plans.get(key)
        if p is not None:
            _plans.move_to_end(key)
            p.users += 1
            return p
    L = _lib.lib()
    d = _lib.Desc()
    d.kind, d.dtype, d.real_layout, d.forward = kind, code, layout, int(forward)
    d.ndim, d.naxes = len(shape), len(axes)
    for i, (s, a, b) in enumerate(zip(shape, sin, sout)):
        d.shape[i], d.stride_in[i], d.stride_out[i] = s, a, b
    for i, a in enumerate(axes):
        d.axes[i] = a
    h = C.c_void_p()
    _lib.check(L.impulse_fft_plan_create(C.byref(h), C.byref(d)))
    doomed = []
    with _plans_lock:
        other = _plans.get(key)
        if other is not None:               # another thread built the same plan meanwhile: keep theirs
            doomed.append(h)
            _plans.move_to_end(key)
            other.users += 1
            p = other
        else:
            p = _Plan(h)
            p.users = 1
            _plans[key] = p
            while len(_plans) > _PLAN_CAP:  # evict ONE least recently used entry at a time
                _, old = _plans.popitem(last=False)
                old.evicted = True
                if old.users == 0:
                    doomed.append(old.h)
    for dh in doomed:
        L.impulse_fft_plan_destroy(dh)
    return p


def _release_plan(p: _Plan) -> None:
    with _plans_lock:
        p.users -= 1
        destroy = p.evicted and p.users == 0
    if destroy:
        _lib.lib().impulse_fft_plan_destroy(p.h)


def _device_key():
    try:
        import torch
        return torch.cuda.current_device() if torch.cuda.is_available() else -1
    except ImportError:
        return -1


def _execute(kind, layout, x, out, real_shape, forward, fct):
    """Transform along the last axis of `x` into `out` (shapes already consistent)."""
    if len(real_shape) > _lib.MAX_DIMS:
        raise _lib.FFTError(-1, "too many dimensions")
    if 0 in real_shape:
        return out
    code = B.dtype_code(x)
    p = _acquire_plan(kind, layout, code, real_shape, B.byte_strides(x), B.byte_strides(out), [len(real_shape) - 1], forward)
    try:
        _lib.check(_lib.lib().impulse_fft_execute(p.h, B.ptr(x), B.ptr(out), float(fct), B.stream_of(x, out)))
    finally:
        _release_plan(p)
    return out


def _as_f64(data, complex_ok=True):
    """float -> float64 / Complex64 -> complex128 views (the C backend is float64 only, NC:18-20)."""
    if B.is_torch(data):
        import torch
        if data.dtype in (torch.float64, torch.complex128):
            return data
        return data.to(torch.complex128 if data.is_complex() else torch.float64)
    a = np.asarray(data)
    if np.iscomplexobj(a):
        return a if a.dtype == np.complex128 else a.astype(np.complex128)
    return a if a.dtype == np.float64 else a.astype(np.float64)


def _check_len(n: int):
    if n == 0:
        # the reference dereferences a NULL plan here; we raise instead
        raise _lib.FFTError(-1, "zero-length transform")


# ---- packing helpers (host index shuffles of the Nim wrapper; vectorised over leading axes) ----
def unpackFFT(data):
    """NC:126-158 — packed halfcomplex reals [..., N] -> [..., N/2+1 | (N+1)/2] complex."""
    if B.is_torch(data):
        import torch
        n = data.shape[-1]
        k = (n - 1) // 2
        out = torch.zeros(data.shape[:-1] + (n // 2 + 1,), dtype=torch.complex128, device=data.device)
        out[..., 0] = data[..., 0]
        out[..., 1:k + 1] = torch.complex(data[..., 1:2 * k:2], data[..., 2:2 * k + 1:2])
        if n % 2 == 0:
            out[..., n // 2] = data[..., n - 1]
        return out
    p = np.asarray(data, dtype=np.float64)
    n = p.shape[-1]
    k = (n - 1) // 2
    out = np.zeros(p.shape[:-1] + (n // 2 + 1,), dtype=np.complex128)
    out[..., 0] = p[..., 0]
    out[..., 1:k + 1] = p[..., 1:2 * k:2] + 1j * p[..., 2:2 * k + 1:2]
    if n % 2 == 0:
        out[..., n // 2] = p[..., n - 1]
    return out


def symmTargetSize(data) -> int:
    """NC:173-183 — for complex input the original parity is guessed from the imaginary part of the
    last bin being exactly 0.0 (taken from the first row when batched)."""
    n = data.shape[-1]
    if B.is_complex(data):
        last = data.reshape(-1, n)[0, n - 1]
        imag = float(last.imag)
        return n * 2 - (2 if imag == 0.0 else 1)
    return n


def symmetrize(data):
    """NC:160-187 — recover bins N/2+1..N-1 as Hermitian conjugates; float input is unpacked first."""
    out_len = symmTargetSize(data)
    half = data if B.is_complex(data) else unpackFFT(data)
    m = half.shape[-1]
    if B.is_torch(half):
        import torch
        res = torch.zeros(half.shape[:-1] + (out_len,), dtype=torch.complex128, device=half.device)
        res[..., :m] = half
        cnt = -(-out_len // 2) - 1  # i in 1 ..< ceilDiv(outLen, 2)
        if cnt > 0:
            res[..., out_len - cnt:] = torch.conj(half[..., 1:cnt + 1]).flip(-1)
        return res
    res = np.zeros(half.shape[:-1] + (out_len,), dtype=np.complex128)
    res[..., :m] = half
    cnt = -(-out_len // 2) - 1
    if cnt > 0:
        res[..., out_len - cnt:] = np.conj(half[..., 1:cnt + 1])[..., ::-1]
    return res


# ---- transforms -----------------------------------------------------------------------------
def fft_inplace(data, forward: bool = True, normalize: str = nkBackward, normValue: float = math.inf):
    """`fft(data: var openArray|Tensor)` (NC:305-310, NA:33-37): in place; complex data -> c2c,
    float data -> FFTPACK halfcomplex packing (forward) / its inverse (backward)."""
    n = data.shape[-1]
    _check_len(n)
    if B.dtype_code(data) != _lib.F64:
        raise TypeError("the C-backend API is float64 only (use DataDesc/FFTDesc for float32)")
    fct = _fct(normalize, forward, normValue, n)
    shape = list(data.shape)
    if B.is_complex(data):
        return _execute(_lib.C2C, _lib.HERMITIAN, data, data, shape, forward, fct)
    return _execute(_lib.R2C if forward else _lib.C2R, _lib.HALFCOMPLEX, data, data, shape, forward, fct)


def rfft_packed(data, forward: bool = True, normalize: str = nkBackward, normValue: float = math.inf):
    """NC:312-319 — real transform returned in maximally packed (halfcomplex) float form."""
    x = _as_f64(data)
    n = x.shape[-1]
    _check_len(n)
    fct = _fct(normalize, forward, normValue, n)
    out = B.empty_like_kind(x, x.shape, False, _lib.F64)
    return _execute(_lib.R2C if forward else _lib.C2R, _lib.HALFCOMPLEX, x, out, list(x.shape), forward, fct)


def rfft(data, forward: bool = True, normalize: str = nkBackward, normValue: float = math.inf):
    """NC:321-332 — the non-redundant N/2+1 (even) / (N+1)/2 (odd) complex bins of a real transform."""
    x = _as_f64(data)
    n = x.shape[-1]
    _check_len(n)
    if not forward:
        return unpackFFT(rfft_packed(x, forward, normalize, normValue))
    fct = _fct(normalize, forward, normValue, n)
    out = B.empty_like_kind(x, tuple(x.shape[:-1]) + (n // 2 + 1,), True, _lib.F64)
    return _execute(_lib.R2C, _lib.HERMITIAN, x, out, list(x.shape), True, fct)


def fft(data, forward: bool = True, normalize: str = nkBackward, normValue: float = math.inf):
    """NC:334-348 / NA:68-81 — out of place; always returns the full complex spectrum (for real
    input the Hermitian half is mirrored on the GPU, fusing `symmetrize`)."""
    x = _as_f64(data)
    n = x.shape[-1]
    _check_len(n)
    fct = _fct(normalize, forward, normValue, n)
    if B.is_complex(x):
        out = B.empty_like_kind(x, x.shape, True, _lib.F64)
        return _execute(_lib.C2C, _lib.HERMITIAN, x, out, list(x.shape), forward, fct)
    if not forward:  # reference semantics: symmetrize(rfft_packed(data, forward=false))  (NC:345)
        return symmetrize(rfft_packed(x, forward, normalize, normValue))
    out = B.empty_like_kind(x, x.shape, True, _lib.F64)
    return _execute(_lib.R2C, _lib.FULLSYM, x, out, list(x.shape), True, fct)


def ifft(data, backward: bool = True, normalize: str = nkBackward, normValue: float = math.inf):
    """NC:350-360."""
    return fft(data, not backward, normalize, normValue)
