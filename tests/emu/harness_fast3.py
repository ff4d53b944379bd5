"""TEST INFRASTRUCTURE: builds and drives the host emulation of fast3_kernel (emu_fast3.cpp)."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SO = os.path.join(HERE, "libimpulse_fast3_emu.so")
SRC = os.path.join(HERE, "emu_fast3.cpp")
DEPS = [SRC] + [os.path.join(ROOT, "impulse_b200", "csrc", f) for f in
                ("fast3_device.cuh", "fft_device.cuh", "fft_types.h", "trig_tables.h", "tma_device.cuh")]
KIND = {"c2c": 0, "r2c": 1, "c2r": 2}


def build():
    if os.path.exists(SO) and all(os.path.getmtime(SO) >= os.path.getmtime(d) for d in DEPS):
        return
    cmd = ["g++", "-std=c++17", "-O1", "-Wno-unknown-pragmas", "-fPIC", "-shared", "-pthread", "-o", SO, SRC]
    out = subprocess.run(cmd, capture_output=True, text=True)
    if out.returncode:
        raise RuntimeError(out.stderr[-4000:])


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(SO)
        _lib.emu_fast3.restype = C.c_int
        _lib.emu_fast3.argtypes = [C.c_int] * 5 + [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int64, C.c_int64, C.c_double, C.c_uint]
    return _lib


def run(shape, kind, x, forward=True, fct=1.0, pair=False, ctas=2, prefetch=False, double_buffer=False, staged=False):
    """shape = (R1, R2, R3, E); x = [rows, n] array (real for r2c, complex otherwise).  Returns the transform the
    kernel would write (rows are `forward` transforms; backward = the reference's forward=False semantics)."""
    r1, r2, r3, e = shape
    n = r1 * r2 * r3
    f64 = x.dtype in (np.float64, np.complex128)
    rdt, cdt = (np.float64, np.complex128) if f64 else (np.float32, np.complex64)
    x = np.ascontiguousarray(x)
    rows = x.shape[0]
    if kind == "c2c":
        assert x.shape[1] == n
        oshape, odt = (rows, n), cdt
    elif kind == "r2c":
        assert x.shape[1] == 2 * n
        oshape, odt = (rows, n + 1), cdt
    else:
        assert x.shape[1] == n + 1
        oshape, odt = (rows, 2 * n), rdt
    # guard zones around the output and the input: an out-of-bounds store / a stray load shows up as a changed canary
    # / a NaN in the result
    pad = 64
    flat = np.full(oshape[0] * oshape[1] + 2 * pad, np.nan, odt)
    out = flat[pad:-pad].reshape(oshape)
    xin = np.full(x.size + 2 * pad, np.nan, x.dtype)
    xin[pad:-pad] = x.ravel()
    x = xin[pad:-pad].reshape(x.shape)
    # the kernel's BWD flag: c2c / r2c = backward transform; c2r = "conjugate the input" = forward=True (F_CONJ_IN)
    bwd = (1 if forward else 0) if kind == "c2r" else (0 if forward else 1)
    rc = lib().emu_fast3(r1 * 1000000 + r2 * 10000 + r3 * 100 + e, 1 if f64 else 0, KIND[kind], bwd,
                         (1 if pair else 0) | (2 if prefetch else 0) | (4 if double_buffer else 0) | (8 if staged else 0), x.ctypes.data, out.ctypes.data, rows, x.shape[1], out.shape[1], fct, ctas)
    if rc:
        raise RuntimeError(f"emu_fast3 rc={rc}")
    if not (np.isnan(flat[:pad]).all() and np.isnan(flat[-pad:]).all()):
        raise AssertionError("store outside the output rows")
    return out.copy()
