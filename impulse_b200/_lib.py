"""ctypes binding of libimpulse_fft_b200.so (include/impulse_fft_b200.h, include/pocketfft.h).

The library is built in-tree by ``__graft_entry__.build()`` / ``make -C impulse_b200/csrc``.
Loading fails loudly when it is missing; calls fail loudly when there is no B200 — there is no
CPU transform path in this package.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(HERE, "libimpulse_fft_b200.so")

MAX_DIMS = 8
C2C, R2C, C2R = 0, 1, 2
F32, F64 = 0, 1
HERMITIAN, HALFCOMPLEX, FULLSYM = 0, 1, 2

ERR_NAMES = {0: "OK", -1: "INVALID", -2: "STRIDE", -3: "UNSUPPORTED", -4: "NOMEM", -5: "CUDA", -6: "NO_DEVICE"}


class FFTError(RuntimeError):
    """Raised for a non-zero status of the C ABI (the Nim wrapper raises Exception on rc != 0,
    c_pocketfft/pocketfft.nim:206-214; the C++ backend throws, pocketfft_hdronly.h:446-476)."""

    def __init__(self, code: int, msg: str):
        super().__init__(f"impulse_fft_b200: {ERR_NAMES.get(code, code)}: {msg}")
        self.code = code


class Desc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("dtype", C.c_int32), ("real_layout", C.c_int32), ("forward", C.c_int32),
                ("ndim", C.c_uint32), ("naxes", C.c_uint32),
                ("shape", C.c_size_t * MAX_DIMS), ("stride_in", C.c_ssize_t * MAX_DIMS),
                ("stride_out", C.c_ssize_t * MAX_DIMS), ("axes", C.c_size_t * MAX_DIMS)]


class PlanInfo(C.Structure):
    _fields_ = [("n_steps", C.c_uint32), ("n_fft", C.c_uint32), ("bluestein", C.c_uint32),
                ("lines_per_cta", C.c_uint32), ("threads", C.c_uint32), ("smem_bytes", C.c_uint32),
                ("tmp_bytes", C.c_uint64), ("n_radices", C.c_uint32), ("radices", C.c_uint32 * 32)]


EXPORTS = [
    # include/impulse_fft_b200.h
    "impulse_fft_plan_create", "impulse_fft_plan_destroy", "impulse_fft_execute", "impulse_fft_c2c",
    "impulse_fft_r2c", "impulse_fft_c2r", "impulse_fft_dct", "impulse_fft_dst", "impulse_fft_c2c_mul", "impulse_fft_convolve_axis",
    "impulse_fft_r2r_fftpack", "impulse_fft_r2r_separable_hartley", "impulse_fft_r2r_genuine_hartley", "impulse_fft_cfft_rows", "impulse_fft_rfft_rows",
    "impulse_fft_plan_get_info", "impulse_fft_launch_count", "impulse_fft_last_error", "impulse_fft_version",
    "impulse_fft_last_kernel",
    "impulse_fft_cmul", "impulse_fft_transpose", "impulse_fft_copy2d", "impulse_fft_cols_from_parts", "impulse_fft_gather_parts", "impulse_fft_enable_peer_access", "impulse_fft_ipc_alloc", "impulse_fft_ipc_free",
    "impulse_fft_ipc_open", "impulse_fft_ipc_close", "impulse_fft_bind_host_to_device",
    "impulse_fft_dist_create", "impulse_fft_dist_execute", "impulse_fft_dist_execute_parts", "impulse_fft_dist_shard", "impulse_fft_dist_destroy",
    # include/pocketfft.h
    "make_cfft_plan", "destroy_cfft_plan", "cfft_backward", "cfft_forward", "cfft_length",
    "make_rfft_plan", "destroy_rfft_plan", "rfft_backward", "rfft_forward", "rfft_length",
]

_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise ImportError(f"{SO_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "or `make -C impulse_b200/csrc` (there is no fallback implementation)")
    L = C.CDLL(SO_PATH)
    sp, ssp, vp = C.POINTER(C.c_size_t), C.POINTER(C.c_ssize_t), C.c_void_p
    L.impulse_fft_plan_create.restype = C.c_int
    L.impulse_fft_plan_create.argtypes = [C.POINTER(vp), C.POINTER(Desc)]
    L.impulse_fft_plan_destroy.restype = C.c_int
    L.impulse_fft_plan_destroy.argtypes = [vp]
    L.impulse_fft_execute.restype = C.c_int
    L.impulse_fft_execute.argtypes = [vp, vp, vp, C.c_double, vp]
    for n in ("impulse_fft_c2c", "impulse_fft_r2c", "impulse_fft_c2r"):
        f = getattr(L, n)
        f.restype = C.c_int
        f.argtypes = [C.c_int, C.c_size_t, sp, ssp, ssp, C.c_size_t, sp, C.c_int, vp, vp, C.c_double, C.c_size_t, vp]
    L.impulse_fft_c2c_mul.restype = C.c_int
    L.impulse_fft_c2c_mul.argtypes = [C.c_int, C.c_size_t, sp, ssp, ssp, C.c_size_t, sp, C.c_int, vp, vp, C.c_double, vp, C.c_size_t, vp]
    L.impulse_fft_convolve_axis.restype = C.c_int
    L.impulse_fft_convolve_axis.argtypes = [C.c_int, C.c_size_t, sp, ssp, ssp, C.c_size_t, vp, vp, C.c_double, vp, C.c_size_t, vp]
    L.impulse_fft_r2r_fftpack.restype = C.c_int
    L.impulse_fft_r2r_fftpack.argtypes = [C.c_int, C.c_size_t, sp, ssp, ssp, C.c_size_t, sp, C.c_int, C.c_int, vp, vp, C.c_double,
                                          C.c_size_t, vp]
    for n in ("impulse_fft_r2r_separable_hartley", "impulse_fft_r2r_genuine_hartley"):
        f = getattr(L, n)
        f.restype = C.c_int
        f.argtypes = [C.c_int, C.c_size_t, sp, ssp, ssp, C.c_size_t, sp, vp, vp, C.c_double, C.c_size_t, vp]
    for n in ("impulse_fft_dct", "impulse_fft_dst"):
        f = getattr(L, n)
        f.restype = C.c_int
        f.argtypes = [C.c_int, C.c_size_t, sp, ssp, ssp, C.c_size_t, sp, C.c_int, vp, vp, C.c_double, C.c_int, C.c_size_t, vp]
    for n in ("impulse_fft_cfft_rows", "impulse_fft_rfft_rows"):
        f = getattr(L, n)
        f.restype = C.c_int
        f.argtypes = [vp, C.c_size_t, C.c_size_t, C.c_int, C.c_double, vp]
    L.impulse_fft_plan_get_info.restype = C.c_int
    L.impulse_fft_plan_get_info.argtypes = [vp, C.POINTER(PlanInfo)]
    L.impulse_fft_launch_count.restype = C.c_uint64
    L.impulse_fft_last_error.restype = C.c_char_p
    L.impulse_fft_version.restype = C.c_char_p
    L.impulse_fft_last_kernel.restype = C.c_char_p
    L.impulse_fft_cmul.restype = C.c_int
    L.impulse_fft_cmul.argtypes = [C.c_int, vp, vp, vp, C.c_size_t, C.c_size_t, C.c_double, vp]
    L.impulse_fft_transpose.restype = C.c_int
    L.impulse_fft_transpose.argtypes = [C.c_int, vp, vp, C.c_size_t, C.c_size_t, C.c_size_t, C.c_size_t, C.c_size_t, vp]
    L.impulse_fft_copy2d.restype = C.c_int
    L.impulse_fft_copy2d.argtypes = [C.c_int, vp, vp] + [C.c_size_t] * 7 + [vp]
    L.impulse_fft_ipc_alloc.restype = C.c_int
    L.impulse_fft_ipc_alloc.argtypes = [C.c_size_t, C.POINTER(vp), vp]
    L.impulse_fft_ipc_open.restype = C.c_int
    L.impulse_fft_ipc_open.argtypes = [vp, C.POINTER(vp)]
    for n in ("impulse_fft_ipc_free", "impulse_fft_ipc_close"):
        getattr(L, n).restype = C.c_int
        getattr(L, n).argtypes = [vp]
    L.impulse_fft_enable_peer_access.restype = C.c_int
    L.impulse_fft_enable_peer_access.argtypes = [C.c_int]
    L.impulse_fft_cols_from_parts.restype = C.c_int
    L.impulse_fft_cols_from_parts.argtypes = [C.c_int, C.c_size_t, C.POINTER(vp), C.c_size_t, C.c_size_t, C.c_size_t, C.c_size_t, vp,
                                              C.c_size_t, C.c_int, C.c_double, vp]
    L.impulse_fft_gather_parts.restype = C.c_int
    L.impulse_fft_gather_parts.argtypes = [C.c_int, C.c_size_t, C.POINTER(vp), C.c_size_t, C.c_size_t, C.c_size_t, C.c_size_t, vp,
                                           C.c_size_t, C.c_int, vp]
    L.impulse_fft_bind_host_to_device.restype = C.c_int
    L.impulse_fft_bind_host_to_device.argtypes = [C.c_int, C.POINTER(C.c_int)]
    L.impulse_fft_dist_create.restype = C.c_int
    L.impulse_fft_dist_create.argtypes = [C.POINTER(vp), C.c_int, C.POINTER(Desc), C.c_int, C.POINTER(C.c_int)]
    L.impulse_fft_dist_execute.restype = C.c_int
    L.impulse_fft_dist_execute.argtypes = [vp, vp, vp, C.c_double]
    L.impulse_fft_dist_execute_parts.restype = C.c_int
    L.impulse_fft_dist_execute_parts.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.c_double]
    L.impulse_fft_dist_shard.restype = C.c_int
    L.impulse_fft_dist_shard.argtypes = [vp, C.c_int, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
    L.impulse_fft_dist_destroy.restype = C.c_int
    L.impulse_fft_dist_destroy.argtypes = [vp]
    L.make_cfft_plan.restype = vp
    L.make_cfft_plan.argtypes = [C.c_size_t]
    L.make_rfft_plan.restype = vp
    L.make_rfft_plan.argtypes = [C.c_size_t]
    for n in ("destroy_cfft_plan", "destroy_rfft_plan"):
        getattr(L, n).restype = None
        getattr(L, n).argtypes = [vp]
    for n in ("cfft_forward", "cfft_backward", "rfft_forward", "rfft_backward"):
        getattr(L, n).restype = C.c_int
        getattr(L, n).argtypes = [vp, vp, C.c_double]
    for n in ("cfft_length", "rfft_length"):
        getattr(L, n).restype = C.c_size_t
        getattr(L, n).argtypes = [vp]
    _lib = L
    return L


def check(rc: int) -> None:
    if rc != 0:
        raise FFTError(rc, lib().impulse_fft_last_error().decode(errors="replace"))


def last_kernel() -> str:
    return lib().impulse_fft_last_kernel().decode()


def launch_count() -> int:
    return int(lib().impulse_fft_launch_count())


def bind_host_to_device(device: int):
    """Bind the calling thread to the CPUs of the NUMA node `device` hangs off (pinned buffers allocated afterwards are
    local to that GPU).  Returns the node, or None when the box exposes no NUMA information."""
    node = C.c_int(-1)
    check(lib().impulse_fft_bind_host_to_device(int(device), C.byref(node)))
    return None if node.value < 0 else int(node.value)
