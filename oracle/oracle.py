"""TEST INFRASTRUCTURE — NOT PRODUCT CODE.

ctypes loaders for the two CPU checkers:

* ``Ref``  — the reference's own engines compiled from /root/reference into
  ``oracle/_ref/`` by ``oracle/Makefile`` (kind "reference").
* ``Port`` — this repo's restatement ``oracle/pocketfft_port.c`` (kind "port").

Only tests/, ``__graft_entry__.smoke()`` and bench.py's cpu_baseline /
``--impl reference`` legs may import this module.  Nothing here reads
/root/reference at run time: the GPU box only has the prebuilt ``.so`` files.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

_size_p = C.POINTER(C.c_size_t)
_ssize_p = C.POINTER(C.c_ssize_t)


def build(verbose: bool = False) -> None:
    """Compile the port (always) and oracle/_ref (when /root/reference is present)."""
    out = subprocess.run(["make", "-C", HERE, "all"], capture_output=True, text=True)
    if verbose or out.returncode:
        print(out.stdout, out.stderr)
    if out.returncode:
        raise RuntimeError("oracle build failed")


def _has_avx512() -> bool:
    try:
        with open("/proc/cpuinfo") as f:
            return "avx512f" in f.read()
    except OSError:
        return False


def _arr(vals, ctype):
    return (ctype * len(vals))(*vals)


class Ref:
    """The compiled reference (pocketfft C + pocketfft_hdronly C++)."""

    kind = "reference"

    def __init__(self, path: str | None = None):
        if path is None:
            cand = []
            if _has_avx512():
                cand.append(os.path.join(HERE, "_ref", "libpocketfft_ref_avx512.so"))
            cand.append(os.path.join(HERE, "_ref", "libpocketfft_ref.so"))
            path = next((p for p in cand if os.path.exists(p)), None)
            if path is None:
                raise FileNotFoundError("oracle/_ref not built (run `make -C oracle`)")
        self.path = path
        L = self.lib = C.CDLL(path)
        for name in ("ref_c2c", "ref_r2c", "ref_c2r"):
            f = getattr(L, name)
            f.restype = C.c_int
            f.argtypes = [C.c_int, C.c_size_t, _size_p, _ssize_p, _ssize_p, C.c_size_t, _size_p,
                          C.c_int, C.c_void_p, C.c_void_p, C.c_double, C.c_size_t]
        for name in ("ref_c_cfft_rows", "ref_c_rfft_rows"):
            f = getattr(L, name)
            f.restype = C.c_int
            f.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_int, C.c_double, C.c_int, C.c_int]
        L.ref_last_error.restype = C.c_char_p
        L.ref_hardware_threads.restype = C.c_uint
        if hasattr(L, "ref_r2r_real"):
            L.ref_r2r_real.restype = C.c_int
            L.ref_r2r_real.argtypes = [C.c_int, C.c_int, C.c_size_t, _size_p, _ssize_p, _ssize_p, C.c_size_t, _size_p, C.c_int,
                                       C.c_int, C.c_void_p, C.c_void_p, C.c_double, C.c_size_t]
        if hasattr(L, "ref_r2r"):
            L.ref_r2r.restype = C.c_int
            L.ref_r2r.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_size_t, _size_p, _ssize_p, _ssize_p, C.c_size_t,
                                  _size_p, C.c_void_p, C.c_void_p, C.c_double, C.c_size_t]
        # the ten C symbols (c_pocketfft/pocketfft.h:18-32)
        L.make_cfft_plan.restype = C.c_void_p
        L.make_cfft_plan.argtypes = [C.c_size_t]
        L.make_rfft_plan.restype = C.c_void_p
        L.make_rfft_plan.argtypes = [C.c_size_t]
        for n in ("destroy_cfft_plan", "destroy_rfft_plan"):
            getattr(L, n).restype = None
            getattr(L, n).argtypes = [C.c_void_p]
        for n in ("cfft_forward", "cfft_backward", "rfft_forward", "rfft_backward"):
            getattr(L, n).restype = C.c_int
            getattr(L, n).argtypes = [C.c_void_p, C.c_void_p, C.c_double]
        for n in ("cfft_length", "rfft_length"):
            getattr(L, n).restype = C.c_size_t
            getattr(L, n).argtypes = [C.c_void_p]

    # ---- C engine, rows in place -------------------------------------
    def hardware_threads(self) -> int:
        return int(self.lib.ref_hardware_threads())

    def cfft_rows(self, a: np.ndarray, forward=True, fct=1.0, nthreads=1, plan_per_row=False):
        """a: complex128 [rows, n] C-contiguous, transformed in place."""
        assert a.dtype == np.complex128 and a.flags.c_contiguous and a.ndim == 2
        rc = self.lib.ref_c_cfft_rows(a.ctypes.data, a.shape[0], a.shape[1], int(forward), fct,
                                      nthreads, int(plan_per_row))
        if rc:
            raise RuntimeError("reference cfft failed")
        return a

    def rfft_rows(self, a: np.ndarray, forward=True, fct=1.0, nthreads=1, plan_per_row=False):
        """a: float64 [rows, n]; in place, FFTPACK halfcomplex packing."""
        assert a.dtype == np.float64 and a.flags.c_contiguous and a.ndim == 2
        rc = self.lib.ref_c_rfft_rows(a.ctypes.data, a.shape[0], a.shape[1], int(forward), fct,
                                      nthreads, int(plan_per_row))
        if rc:
            raise RuntimeError("reference rfft failed")
        return a

    # ---- C++ engine, N-D strided ---------------------------------------
    def _nd(self, fn, a_in, a_out, real_shape, axes, forward, fct, nthreads):
        dt = {np.dtype(np.float32): 0, np.dtype(np.complex64): 0,
              np.dtype(np.float64): 1, np.dtype(np.complex128): 1}[a_in.dtype]
        nd = len(real_shape)
        rc = fn(dt, nd, _arr(list(real_shape), C.c_size_t), _arr(list(a_in.strides), C.c_ssize_t),
                _arr(list(a_out.strides), C.c_ssize_t), len(axes), _arr(list(axes), C.c_size_t),
                int(forward), a_in.ctypes.data, a_out.ctypes.data, float(fct), nthreads)
        if rc:
            raise RuntimeError(self.lib.ref_last_error().decode())
        return a_out

    def c2c(self, a, axes, forward=True, fct=1.0, out=None, nthreads=1):
        out = np.empty_like(a) if out is None else out
        return self._nd(self.lib.ref_c2c, a, out, a.shape, axes, forward, fct, nthreads)

    def r2c(self, a, axes, forward=True, fct=1.0, out=None, nthreads=1):
        if out is None:
            shp = list(a.shape)
            shp[axes[-1]] = shp[axes[-1]] // 2 + 1
            out = np.zeros(shp, dtype=np.complex64 if a.dtype == np.float32 else np.complex128)
        return self._nd(self.lib.ref_r2c, a, out, a.shape, axes, forward, fct, nthreads)

    def c2r(self, a, real_shape, axes, forward=False, fct=1.0, out=None, nthreads=1):
        if out is None:
            out = np.empty(real_shape, dtype=np.float32 if a.dtype == np.complex64 else np.float64)
        return self._nd(self.lib.ref_c2r, a, out, real_shape, axes, forward, fct, nthreads)

    def r2r(self, cosine, type_, a, axes, fct=1.0, ortho=False, out=None, nthreads=1):
        """pocketfft::dct (cosine=True) / dst of type 1..4."""
        out = np.empty_like(a) if out is None else out
        nd = a.ndim
        rc = self.lib.ref_r2r(int(cosine), type_, int(ortho), 1 if a.dtype == np.float64 else 0, nd,
                              _arr(list(a.shape), C.c_size_t), _arr(list(a.strides), C.c_ssize_t),
                              _arr(list(out.strides), C.c_ssize_t), len(axes), _arr(list(axes), C.c_size_t),
                              a.ctypes.data, out.ctypes.data, float(fct), nthreads)
        if rc:
            raise RuntimeError(self.lib.ref_last_error().decode())
        return out


    def r2r_real(self, which, a, axes, real2hermitian=True, forward=True, fct=1.0, out=None, nthreads=1):
        """pocketfft::r2r_fftpack ('fftpack') / r2r_separable_hartley / r2r_genuine_hartley."""
        w = {"fftpack": 0, "separable_hartley": 1, "genuine_hartley": 2}[which]
        out = np.empty_like(a) if out is None else out
        nd = a.ndim
        rc = self.lib.ref_r2r_real(w, 1 if a.dtype == np.float64 else 0, nd, _arr(list(a.shape), C.c_size_t),
                                   _arr(list(a.strides), C.c_ssize_t), _arr(list(out.strides), C.c_ssize_t), len(axes),
                                   _arr(list(axes), C.c_size_t), int(real2hermitian), int(forward), a.ctypes.data,
                                   out.ctypes.data, float(fct), nthreads)
        if rc:
            raise RuntimeError(self.lib.ref_last_error().decode())
        return out


class Port:
    """This repo's restatement (oracle/pocketfft_port.c)."""

    kind = "port"

    def __init__(self, path: str | None = None):
        path = path or os.path.join(HERE, "libpocketfft_port.so")
        if not os.path.exists(path):
            build()
        self.path = path
        L = self.lib = C.CDLL(path)
        for sfx in ("f64", "f32"):
            for n in ("port_cfft_rows_", "port_rfft_rows_"):
                f = getattr(L, n + sfx)
                f.restype = C.c_int
                f.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_int, C.c_double]
            f = getattr(L, "port_good_size_" + sfx)
            f.restype = C.c_size_t
            f.argtypes = [C.c_size_t]
            f = getattr(L, "port_uses_bluestein_" + sfx)
            f.restype = C.c_int
            f.argtypes = [C.c_size_t, C.c_int]
            f = getattr(L, "port_factors_" + sfx)
            f.restype = C.c_int
            f.argtypes = [C.c_size_t, _size_p]

    def hardware_threads(self) -> int:
        return 1

    @staticmethod
    def _sfx(a):
        return "f32" if a.dtype in (np.float32, np.complex64) else "f64"

    def cfft_rows(self, a, forward=True, fct=1.0, nthreads=1, plan_per_row=False):
        assert a.dtype in (np.complex128, np.complex64) and a.flags.c_contiguous and a.ndim == 2
        rc = getattr(self.lib, "port_cfft_rows_" + self._sfx(a))(
            a.ctypes.data, a.shape[0], a.shape[1], int(forward), fct)
        if rc:
            raise RuntimeError("port cfft failed")
        return a

    def rfft_rows(self, a, forward=True, fct=1.0, nthreads=1, plan_per_row=False):
        assert a.dtype in (np.float64, np.float32) and a.flags.c_contiguous and a.ndim == 2
        rc = getattr(self.lib, "port_rfft_rows_" + self._sfx(a))(
            a.ctypes.data, a.shape[0], a.shape[1], int(forward), fct)
        if rc:
            raise RuntimeError("port rfft failed")
        return a

    def factors(self, n):
        buf = (C.c_size_t * 32)()
        k = self.lib.port_factors_f64(n, buf)
        return list(buf[:k])

    def uses_bluestein(self, n, real_input=False):
        return bool(self.lib.port_uses_bluestein_f64(n, int(real_input)))

    def good_size(self, n):
        return int(self.lib.port_good_size_f64(n))

    # ---- N-D composition mirroring pocketfft_hdronly.h:3272-3390 --------
    def _lines(self, a, axis, fn):
        """apply fn to contiguous [lines, n] copies of the lines along `axis`."""
        m = np.moveaxis(a, axis, -1)
        flat = np.ascontiguousarray(m).reshape(-1, m.shape[-1])
        res = fn(flat)
        return np.moveaxis(res.reshape(m.shape[:-1] + (res.shape[-1],)), -1, axis)

    def c2c(self, a, axes, forward=True, fct=1.0, out=None, nthreads=1):
        res = np.array(a, copy=True)
        for n, ax in enumerate(axes):
            f = fct if n == 0 else 1.0  # fct on the first axis only: hdronly.h:3048
            if res.shape[ax] == 1:  # the C++ engine scales length-1 lines (hdronly.h:1311), the C engine does not
                res = (res * f).astype(res.dtype)
                continue
            res = self._lines(res, ax, lambda r: self.cfft_rows(r, forward, f))
        if out is not None:
            out[...] = res
            return out
        return np.ascontiguousarray(res)

    @staticmethod
    def _unpack_rows(p):
        """halfcomplex rows -> N/2+1 complex (hdronly.h:3146-3158)."""
        n = p.shape[1]
        out = np.zeros((p.shape[0], n // 2 + 1), dtype=np.complex64 if p.dtype == np.float32 else np.complex128)
        out[:, 0] = p[:, 0]
        k = (n - 1) // 2
        out[:, 1:k + 1] = p[:, 1:2 * k:2] + 1j * p[:, 2:2 * k + 1:2]
        if n % 2 == 0:
            out[:, n // 2] = p[:, n - 1]
        return out

    @staticmethod
    def _pack_rows(c, n):
        """N/2+1 complex rows -> halfcomplex (hdronly.h:3196-3215)."""
        p = np.empty((c.shape[0], n), dtype=np.float32 if c.dtype == np.complex64 else np.float64)
        p[:, 0] = c[:, 0].real
        k = (n - 1) // 2
        p[:, 1:2 * k:2] = c[:, 1:k + 1].real
        p[:, 2:2 * k + 1:2] = c[:, 1:k + 1].imag
        if n % 2 == 0:
            p[:, n - 1] = c[:, n // 2].real
        return p

    def r2c(self, a, axes, forward=True, fct=1.0, out=None, nthreads=1):
        def one(r):
            if r.shape[1] == 1:
                return (r * fct).astype(np.complex64 if r.dtype == np.float32 else np.complex128)
            r = self.rfft_rows(np.array(r, copy=True), True, fct)
            c = self._unpack_rows(r)
            return c if forward else np.conj(c)  # hdronly.h:3152-3155
        res = self._lines(np.asarray(a), axes[-1], one)
        if len(axes) > 1:
            res = self.c2c(np.ascontiguousarray(res), list(axes[:-1]), forward, 1.0)
        if out is not None:
            out[...] = res
            return out
        return np.ascontiguousarray(res)

    def c2r(self, a, real_shape, axes, forward=False, fct=1.0, out=None, nthreads=1):
        a = np.asarray(a)
        if len(axes) > 1:
            a = self.c2c(a, list(axes[:-1]), forward, 1.0)
        n = real_shape[axes[-1]]

        def one(c):
            if n == 1:
                return (c.real * fct).astype(np.float32 if c.dtype == np.complex64 else np.float64)
            c = np.conj(c) if forward else c  # hdronly.h:3202-3208
            return self.rfft_rows(self._pack_rows(c, n), False, fct)
        res = self._lines(a, axes[-1], one)
        if out is not None:
            out[...] = res
            return out
        return np.ascontiguousarray(res)


def r2r_direct(cosine, type_, x, fct=1.0, ortho=False):
    """O(N^2) restatement of the DCT/DST definitions pocketfft implements (FFTW's REDFT/RODFT kinds,
    README_pocketfft.md:220-241 for `ortho`), along the last axis.  Checker of last resort for small N and
    the thing the compiled reference's conventions are pinned against in tests/test_oracle.py."""
    x = np.array(x, dtype=np.float64, copy=True)
    n = x.shape[-1]
    j = np.arange(n)[:, None]
    k = np.arange(n)[None, :]
    r2 = np.sqrt(2.0)
    if ortho and type_ == 1 and cosine:
        x[..., 0] *= r2
        x[..., -1] *= r2
    if ortho and type_ == 3:
        x[..., 0] *= r2
    if cosine:
        if type_ == 1:
            m = 2 * np.cos(np.pi * j * k / (n - 1))
            m[0, :] = 1.0
            m[n - 1, :] = (-1.0) ** np.arange(n)
        elif type_ == 2:
            m = 2 * np.cos(np.pi * (j + 0.5) * k / n)
        elif type_ == 3:
            m = 2 * np.cos(np.pi * j * (k + 0.5) / n)
            m[0, :] = 1.0
        else:
            m = 2 * np.cos(np.pi * (j + 0.5) * (k + 0.5) / n)
    else:
        if type_ == 1:
            m = 2 * np.sin(np.pi * (j + 1) * (k + 1) / (n + 1))
        elif type_ == 2:
            m = 2 * np.sin(np.pi * (j + 0.5) * (k + 1) / n)
        elif type_ == 3:
            m = 2 * np.sin(np.pi * (j + 1) * (k + 0.5) / n)
            m[n - 1, :] = (-1.0) ** np.arange(n)
        else:
            m = 2 * np.sin(np.pi * (j + 0.5) * (k + 0.5) / n)
    y = (x @ m) * fct
    if ortho and type_ == 1 and cosine:
        y[..., 0] /= r2
        y[..., -1] /= r2
    if ortho and type_ == 2:
        y[..., 0] /= r2
    return y


def load(prefer_ref: bool = True):
    """Best available checker: the compiled reference, else the port."""
    if prefer_ref:
        try:
            return Ref()
        except (FileNotFoundError, OSError):
            pass
    return Port()


def rel_l2(a, b) -> float:
    """sqrt(sum|a-b|^2 / sum|b|^2) — the metric of tests/test_fft.nim:14-22."""
    a = np.asarray(a)
    b = np.asarray(b)
    den = float(np.sum(np.abs(b.astype(np.complex128)) ** 2))
    num = float(np.sum(np.abs(a.astype(np.complex128) - b.astype(np.complex128)) ** 2))
    if den == 0.0:
        return 0.0 if num == 0.0 else float("inf")
    return (num / den) ** 0.5


def max_row_rel_l2(a, b) -> float:
    a = np.asarray(a).reshape(-1, np.asarray(a).shape[-1]).astype(np.complex128)
    b = np.asarray(b).reshape(-1, np.asarray(b).shape[-1]).astype(np.complex128)
    num = np.sum(np.abs(a - b) ** 2, axis=1)
    den = np.sum(np.abs(b) ** 2, axis=1)
    den = np.where(den == 0, 1.0, den)
    return float(np.sqrt(np.max(num / den)))


# ---- numpy restatements of the real-to-real FFTPACK / Hartley entry points -------------------------
def _halfcomplex_pack(spec, n):
    """N/2+1 complex bins (last axis) -> FFTPACK halfcomplex reals [R0, R1, I1, ..., (R_{N/2})]."""
    out = np.empty(spec.shape[:-1] + (n,), dtype=np.float64)
    out[..., 0] = spec[..., 0].real
    m = (n - 1) // 2
    out[..., 1:2 * m + 1:2] = spec[..., 1:m + 1].real
    out[..., 2:2 * m + 2:2] = spec[..., 1:m + 1].imag
    if n % 2 == 0:
        out[..., n - 1] = spec[..., n // 2].real
    return out


def _halfcomplex_unpack(x):
    n = x.shape[-1]
    spec = np.zeros(x.shape[:-1] + (n // 2 + 1,), dtype=np.complex128)
    spec[..., 0] = x[..., 0]
    m = (n - 1) // 2
    spec[..., 1:m + 1] = x[..., 1:2 * m + 1:2] + 1j * x[..., 2:2 * m + 2:2]
    if n % 2 == 0:
        spec[..., n // 2] = x[..., n - 1]
    return spec


def fftpack_numpy(a, axes, real2hermitian=True, forward=True, fct=1.0):
    """r2r_fftpack as the VENDORED header computes it (ExecR2R, pocketfft_hdronly.h:3123-3143): the 1-D plan
    is executed with r2hc = `forward`; (not real2hermitian) and forward negates elements 2,4,.. of the
    input, real2hermitian and (not forward) negates them on the output.  Axes in the given order, fct once."""
    x = np.array(a, dtype=np.float64, copy=True)
    for i, ax in enumerate(axes):
        x = np.moveaxis(x, ax, -1).copy()
        n = x.shape[-1]
        if (not real2hermitian) and forward:
            x[..., 2::2] *= -1.0
        if forward:
            x = _halfcomplex_pack(np.fft.rfft(x, axis=-1), n)
        else:
            x = np.fft.irfft(_halfcomplex_unpack(x), n=n, axis=-1) * n
        if real2hermitian and not forward:
            x[..., 2::2] *= -1.0
        if i == 0:
            x = x * fct
        x = np.moveaxis(x, -1, ax)
    return np.ascontiguousarray(x)


def hartley_numpy(a, axes, genuine=False, fct=1.0):
    """r2r_separable_hartley / r2r_genuine_hartley (pocketfft_hdronly.h:3066-3103, 3405-3445): Re + Im of
    the forward transform, per axis (separable) or of the N-D transform (genuine)."""
    x = np.array(a, dtype=np.float64, copy=True)
    if genuine:
        f = np.fft.fftn(x, axes=axes)
        return (f.real + f.imag) * fct
    for ax in axes:
        f = np.fft.fft(x, axis=ax)
        x = f.real + f.imag
    return x * fct
