"""TEST INFRASTRUCTURE: builds and drives the host emulation of the CUDA engine (emu.cpp)."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SO = os.path.join(HERE, "libimpulse_fft_emu.so")
SRC = [os.path.join(HERE, "emu.cpp"), os.path.join(HERE, "emu_col.cpp"), os.path.join(ROOT, "impulse_b200", "csrc", "planner.cpp")]
DEPS = SRC + [os.path.join(ROOT, "impulse_b200", "csrc", f) for f in
              ("fft_device.cuh", "fft_types.h", "planner.h", "trig_tables.h", "col_device.cuh", "colconvw_device.cuh", "fast3_device.cuh")]

KIND = {"c2c": 0, "r2c": 1, "c2r": 2}
LAYOUT = {"hermitian": 0, "halfcomplex": 1, "fullsym": 2}


def build():
    if os.path.exists(SO) and all(os.path.getmtime(SO) >= os.path.getmtime(d) for d in DEPS):
        return
    cmd = ["g++", "-std=c++17", "-O2", "-Wno-unknown-pragmas", "-fPIC", "-shared", "-pthread", "-o", SO] + SRC
    out = subprocess.run(cmd, capture_output=True, text=True)
    if out.returncode:
        raise RuntimeError(out.stderr)


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(SO)
        _lib.emu_nd.restype = C.c_int
        _lib.emu_nd.argtypes = [C.c_int, C.c_int, C.c_int, C.c_size_t, C.POINTER(C.c_size_t),
                                C.POINTER(C.c_ssize_t), C.POINTER(C.c_ssize_t), C.c_size_t,
                                C.POINTER(C.c_size_t), C.c_int, C.c_void_p, C.c_void_p, C.c_double]
        _lib.emu_last_error.restype = C.c_char_p
        _lib.emu_nd_steps.restype = C.c_int
        _lib.emu_nd_steps.argtypes = _lib.emu_nd.argtypes[:10]
        _lib.emu_plan_info.restype = C.c_int
        _lib.emu_plan_info.argtypes = [C.c_uint32, C.c_int, C.POINTER(C.c_uint32), C.POINTER(C.c_int),
                                       C.POINTER(C.c_uint32), C.c_int]
    return _lib


class EmuError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"emu rc={code}: {msg}")
        self.code = code


def nd(kind, a_in, a_out, shape, axes, forward=True, fct=1.0, layout="hermitian"):
    L = lib()
    dt = 1 if a_in.dtype in (np.float64, np.complex128) else 0
    n = len(shape)
    rc = L.emu_nd(KIND[kind], dt, LAYOUT[layout], n, (C.c_size_t * n)(*shape),
                  (C.c_ssize_t * n)(*a_in.strides), (C.c_ssize_t * n)(*a_out.strides), len(axes),
                  (C.c_size_t * len(axes))(*axes), int(forward), a_in.ctypes.data, a_out.ctypes.data, fct)
    if rc:
        raise EmuError(rc, L.emu_last_error().decode())
    return a_out


def c2c_mul(a_in, a_out, axes, mul, forward=True, fct=1.0):
    """Mirror of impulse_fft_c2c_mul: a_out = c2c(a_in) * mul[flat offset % mul.size]."""
    L = lib()
    L.emu_c2c_mul.restype = C.c_int
    L.emu_c2c_mul.argtypes = [C.c_int, C.c_size_t, C.POINTER(C.c_size_t), C.POINTER(C.c_ssize_t), C.POINTER(C.c_ssize_t),
                              C.c_size_t, C.POINTER(C.c_size_t), C.c_int, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p,
                              C.c_size_t]
    dt = 1 if a_in.dtype == np.complex128 else 0
    n = a_in.ndim
    rc = L.emu_c2c_mul(dt, n, (C.c_size_t * n)(*a_in.shape), (C.c_ssize_t * n)(*a_in.strides),
                       (C.c_ssize_t * n)(*a_out.strides), len(axes), (C.c_size_t * len(axes))(*axes), int(forward),
                       a_in.ctypes.data, a_out.ctypes.data, fct, mul.ctypes.data, mul.size)
    if rc:
        raise EmuError(rc, L.emu_last_error().decode())
    return a_out


def convolve_axis(a_in, a_out, axis, mul, fct=1.0):
    L = lib()
    L.emu_convolve_axis.restype = C.c_int
    L.emu_convolve_axis.argtypes = [C.c_int, C.c_size_t, C.POINTER(C.c_size_t), C.POINTER(C.c_ssize_t), C.POINTER(C.c_ssize_t),
                                    C.c_size_t, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_size_t]
    dt = 1 if a_in.dtype == np.complex128 else 0
    n = a_in.ndim
    rc = L.emu_convolve_axis(dt, n, (C.c_size_t * n)(*a_in.shape), (C.c_ssize_t * n)(*a_in.strides),
                             (C.c_ssize_t * n)(*a_out.strides), axis, a_in.ctypes.data, a_out.ctypes.data, fct, mul.ctypes.data,
                             mul.size)
    if rc:
        raise EmuError(rc, L.emu_last_error().decode())
    return a_out


def r2r_real(which, a_in, a_out, axes, real2hermitian=True, forward=True, fct=1.0):
    """which: 'fftpack' | 'separable_hartley' | 'genuine_hartley' (mirrors the impulse_fft_r2r_* entry points)."""
    L = lib()
    L.emu_r2r_real.restype = C.c_int
    L.emu_r2r_real.argtypes = [C.c_int, C.c_int, C.c_size_t, C.POINTER(C.c_size_t), C.POINTER(C.c_ssize_t), C.POINTER(C.c_ssize_t),
                               C.c_size_t, C.POINTER(C.c_size_t), C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_double]
    w = {"fftpack": 0, "separable_hartley": 1, "genuine_hartley": 2}[which]
    dt = 1 if a_in.dtype == np.float64 else 0
    n = a_in.ndim
    rc = L.emu_r2r_real(w, dt, n, (C.c_size_t * n)(*a_in.shape), (C.c_ssize_t * n)(*a_in.strides),
                        (C.c_ssize_t * n)(*a_out.strides), len(axes), (C.c_size_t * len(axes))(*axes), int(real2hermitian),
                        int(forward), a_in.ctypes.data, a_out.ctypes.data, fct)
    if rc:
        raise EmuError(rc, L.emu_last_error().decode())
    return a_out


def r2r(cosine, type_, a_in, a_out, axes, fct=1.0, ortho=False):
    L = lib()
    if not hasattr(L, "_r2r_bound"):
        L.emu_r2r.restype = C.c_int
        L.emu_r2r.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_size_t, C.POINTER(C.c_size_t), C.POINTER(C.c_ssize_t),
                              C.POINTER(C.c_ssize_t), C.c_size_t, C.POINTER(C.c_size_t), C.c_void_p, C.c_void_p, C.c_double]
        L._r2r_bound = True
    dt = 1 if a_in.dtype == np.float64 else 0
    n = a_in.ndim
    rc = L.emu_r2r(int(cosine), type_, int(ortho), dt, n, (C.c_size_t * n)(*a_in.shape), (C.c_ssize_t * n)(*a_in.strides),
                   (C.c_ssize_t * n)(*a_out.strides), len(axes), (C.c_size_t * len(axes))(*axes), a_in.ctypes.data,
                   a_out.ctypes.data, fct)
    if rc:
        raise EmuError(rc, L.emu_last_error().decode())
    return a_out


def nd_steps(kind, a_in, a_out, shape, axes, forward=True, layout="hermitian"):
    L = lib()
    dt = 1 if a_in.dtype in (np.float64, np.complex128) else 0
    n = len(shape)
    rc = L.emu_nd_steps(KIND[kind], dt, LAYOUT[layout], n, (C.c_size_t * n)(*shape), (C.c_ssize_t * n)(*a_in.strides),
                        (C.c_ssize_t * n)(*a_out.strides), len(axes), (C.c_size_t * len(axes))(*axes), int(forward))
    if rc < 0:
        raise EmuError(rc, L.emu_last_error().decode())
    return rc


def nd_fast_ids(kind, a_in, a_out, shape, axes, forward=True, layout="hermitian"):
    """FastId of every step the planner emits (0 = generic engine, 0xffffffff = elementwise / fold pass)."""
    L = lib()
    L.emu_nd_fast_ids.restype = C.c_int
    dt = 1 if a_in.dtype in (np.float64, np.complex128) else 0
    n = len(shape)
    ids = (C.c_uint32 * 16)()
    rc = L.emu_nd_fast_ids(KIND[kind], dt, LAYOUT[layout], C.c_size_t(n), (C.c_size_t * n)(*shape), (C.c_ssize_t * n)(*a_in.strides),
                           (C.c_ssize_t * n)(*a_out.strides), C.c_size_t(len(axes)), (C.c_size_t * len(axes))(*axes), int(forward), ids, 16)
    if rc < 0:
        raise EmuError(rc, L.emu_last_error().decode())
    return list(ids[:rc])


def nd_fuse_flags(kind, a_in, a_out, shape, axes, forward=True, layout="hermitian"):
    """1 for every step that is the first launch of a fusable pair (two launches of a split column transform)."""
    L = lib()
    L.emu_nd_fuse_flags.restype = C.c_int
    dt = 1 if a_in.dtype in (np.float64, np.complex128) else 0
    n = len(shape)
    fl = (C.c_uint32 * 16)()
    rc = L.emu_nd_fuse_flags(KIND[kind], dt, LAYOUT[layout], C.c_size_t(n), (C.c_size_t * n)(*shape), (C.c_ssize_t * n)(*a_in.strides),
                             (C.c_ssize_t * n)(*a_out.strides), C.c_size_t(len(axes)), (C.c_size_t * len(axes))(*axes), int(forward), fl, 16)
    if rc < 0:
        raise EmuError(rc, L.emu_last_error().decode())
    return list(fl[:rc])


def set_fast_cols(pipe_groups=None):
    """None: every job on the generic engine's phase emulation.  An integer g >= 0: column-kernel jobs (strided axes,
    the four-step split, the fused convolution pass) on the thread-level emulation, colpipe2 with g groups per CTA
    (0 or 1 = the plain colfast2 kernel; the product default is 2)."""
    L = lib()
    L.emu_set_fast_cols(0 if pipe_groups is None else int(pipe_groups) + 1)


def col_job_count():
    """Jobs that ran on the thread-level column-kernel emulation since the last set_fast_cols()."""
    return int(lib().emu_col_job_count())


def plan_info(L_, dtype=1):
    L = lib()
    n_fft = C.c_uint32()
    blue = C.c_int()
    rad = (C.c_uint32 * 32)()
    k = L.emu_plan_info(L_, dtype, C.byref(n_fft), C.byref(blue), rad, 32)
    if k < 0:
        raise EmuError(k, L.emu_last_error().decode())
    return n_fft.value, bool(blue.value), list(rad[:k])
