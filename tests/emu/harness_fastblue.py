"""TEST INFRASTRUCTURE: builds and drives the host emulation of fastblue_kernel (emu_fastblue.cpp)."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SO = os.path.join(HERE, "libimpulse_fastblue_emu.so")
CSRC = os.path.join(ROOT, "impulse_b200", "csrc")
SRC = [os.path.join(HERE, "emu_fastblue.cpp"), os.path.join(CSRC, "planner.cpp")]
DEPS = SRC + [os.path.join(CSRC, f) for f in ("fastblue_device.cuh", "fast3_device.cuh", "fft_device.cuh", "fft_types.h",
                                                "trig_tables.h", "planner.h", "tmem_device.cuh")]
KIND = {"c2c": 0, "r2c": 1, "c2r": 2}


def build():
    if os.path.exists(SO) and all(os.path.getmtime(SO) >= os.path.getmtime(d) for d in DEPS):
        return
    cmd = ["g++", "-std=c++17", "-O1", "-Wno-unknown-pragmas", "-fPIC", "-shared", "-pthread", "-o", SO] + SRC
    out = subprocess.run(cmd, capture_output=True, text=True)
    if out.returncode:
        raise RuntimeError(out.stderr[-4000:])


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(SO)
        _lib.emu_fastblue.restype = C.c_int
        _lib.emu_fastblue.argtypes = [C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int64,
                                      C.c_int64, C.c_double, C.c_uint]
        _lib.emu_fast4_8192.restype = C.c_int
        _lib.emu_fast4_8192.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int64, C.c_int64, C.c_double, C.c_uint]
    return _lib


def run_fast4(x, forward=True, fct=1.0, ctas=2):
    """c2c rows of 8192 complex128 points through fast4_8192_kernel."""
    x = np.ascontiguousarray(x, dtype=np.complex128)
    assert x.shape[1] == 8192
    pad = 64
    flat = np.full(x.size + 2 * pad, np.nan, np.complex128)
    out = flat[pad:-pad].reshape(x.shape)
    rc = lib().emu_fast4_8192(0 if forward else 1, x.ctypes.data, out.ctypes.data, x.shape[0], 8192, 8192, fct, ctas)
    if rc:
        raise RuntimeError(f"emu_fast4_8192 rc={rc}")
    if not (np.isnan(flat[:pad]).all() and np.isnan(flat[-pad:]).all()):
        raise AssertionError("store outside the output rows")
    return out.copy()


def run(kind, x, length, forward=True, fct=1.0, bk_smem=False, bf_early=False, ctas=2, four_pass=False, bf_tmem=False):
    """x = [rows, L] complex (c2c), [rows, L] real (r2c) or [rows, L//2+1] complex (c2r); float64 or float32.
    `forward` has the reference's meaning (pocketfft_hdronly.h:3125-3250)."""
    x = np.ascontiguousarray(x)
    rows = x.shape[0]
    f32 = x.dtype in (np.float32, np.complex64)
    cdt, rdt = (np.complex64, np.float32) if f32 else (np.complex128, np.float64)
    if kind == "c2c":
        oshape, odt, bwd = (rows, length), cdt, not forward
    elif kind == "r2c":
        oshape, odt, bwd = (rows, length // 2 + 1), cdt, not forward
    else:
        oshape, odt, bwd = (rows, length), rdt, forward      # c2r: BWD = "conjugate the input" = forward=True
    pad = 64
    flat = np.full(oshape[0] * oshape[1] + 2 * pad, np.nan, odt)
    out = flat[pad:-pad].reshape(oshape)
    rc = lib().emu_fastblue(KIND[kind], 1 if bwd else 0, (1 if bk_smem else 0) | (2 if bf_early else 0) | (4 if four_pass else 0) | (8 if f32 else 0) | (16 if bf_tmem else 0), length, x.ctypes.data,
                            out.ctypes.data, rows, x.shape[1], out.shape[1], fct, ctas)
    if rc:
        raise RuntimeError(f"emu_fastblue rc={rc}")
    if not (np.isnan(flat[:pad]).all() and np.isnan(flat[-pad:]).all()):
        raise AssertionError("store outside the output rows")
    return out.copy()
