"""Buffer adapters: numpy arrays (host memory) and torch CUDA tensors (device memory).

PyTorch is plumbing only (device memory + streams); it is imported lazily so the host-array API
works without it.
"""
from __future__ import annotations

import numpy as np

from . import _lib

_NP_DT = {np.dtype(np.float32): _lib.F32, np.dtype(np.complex64): _lib.F32,
          np.dtype(np.float64): _lib.F64, np.dtype(np.complex128): _lib.F64}


def _torch():
    import torch
    return torch


def is_torch(x) -> bool:
    return type(x).__module__.split(".")[0] == "torch"


def is_complex(x) -> bool:
    if is_torch(x):
        return x.is_complex()
    return np.iscomplexobj(x)


def dtype_code(x) -> int:
    if is_torch(x):
        t = _torch()
        if x.dtype in (t.float32, t.complex64):
            return _lib.F32
        if x.dtype in (t.float64, t.complex128):
            return _lib.F64
        raise TypeError(f"unsupported dtype {x.dtype}")
    try:
        return _NP_DT[np.dtype(x.dtype)]
    except KeyError:
        raise TypeError(f"unsupported dtype {x.dtype}") from None


def ptr(x) -> int:
    return int(x.data_ptr()) if is_torch(x) else int(x.ctypes.data)


def byte_strides(x):
    if is_torch(x):
        es = x.element_size()
        return [int(s) * es for s in x.stride()]
    return [int(s) for s in x.strides]


def stream_of(*xs):
    """cudaStream_t of torch's current stream when any buffer is a CUDA tensor, else NULL."""
    for x in xs:
        if is_torch(x):
            if not x.is_cuda:
                raise TypeError("torch tensors must live on a CUDA device (use numpy arrays for host data)")
            return int(_torch().cuda.current_stream(x.device).cuda_stream)
    return None


def empty_like_kind(x, shape, complex_out: bool, code: int):
    """New buffer of the same family (numpy/torch, same device) as x."""
    if is_torch(x):
        t = _torch()
        dt = {(_lib.F32, False): t.float32, (_lib.F32, True): t.complex64,
              (_lib.F64, False): t.float64, (_lib.F64, True): t.complex128}[(code, complex_out)]
        return t.empty(tuple(shape), dtype=dt, device=x.device)
    dt = {(_lib.F32, False): np.float32, (_lib.F32, True): np.complex64,
          (_lib.F64, False): np.float64, (_lib.F64, True): np.complex128}[(code, complex_out)]
    return np.empty(tuple(shape), dtype=dt)
