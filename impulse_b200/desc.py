"""Mirror of the reference's C++-backend API: ``DataDesc``, ``FFTDesc`` and ``apply``
(/root/reference/impulse/fft/cpp_pocketfft/pocketfft.nim:137-149, 158-215, 235-277), running on
the B200 engine through the C ABI (``impulse_fft_c2c/r2c/c2r``, include/impulse_fft_b200.h).

Same names, argument meaning and error behaviour:
  * ``DataDesc.init(buffer, shape[, stride])`` — stride in ELEMENTS as in Nim (pocketfft.nim:158-176),
    converted to bytes; without stride the data is C-contiguous (pocketfft.nim:178-199).
  * ``FFTDesc.init(axes, forward, scalingFactor=1, nthreads=1)`` (pocketfft.nim:201-215).
  * ``fft.apply(descOut, descIn)`` dispatches on the complexness of In/Out (pocketfft.nim:235-277):
    complex->complex = c2c, real->complex = r2c, complex->real = c2r; real->real raises.
Deliberate fixes of reference quirks: the c2r branch passes the REAL (output)
shape as pocketfft requires, and ``assert shape.len == stride.len`` is a comparison.
Errors surface as ``FFTError`` where the reference throws from pocketfft::util::sanity_check.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Sequence

from . import _buffers as B
from . import _lib


@dataclass
class DataDesc:
    buf: object
    shape: list
    stride: list  # bytes, like the Nim object (pocketfft.nim:174,191-194)
    complex: bool = False
    dtype: int = _lib.F64

    @classmethod
    def init(cls, buffer, shape: Sequence[int] | None = None, stride: Sequence[int] | None = None) -> "DataDesc":
        if buffer is None:
            raise AssertionError("buffer must not be nil")  # pocketfft.nim:167
        cplx = B.is_complex(buffer)
        code = B.dtype_code(buffer)
        esz = (8 if code == _lib.F64 else 4) * (2 if cplx else 1)
        if shape is None:
            shape = list(buffer.shape)
            stride_b = B.byte_strides(buffer)
            if stride is not None:
                stride_b = [int(s) * esz for s in stride]
        elif stride is not None:
            assert len(shape) == len(stride)
            stride_b = [int(s) * esz for s in stride]
        else:
            stride_b, acc = [0] * len(shape), esz
            for i in range(len(shape) - 1, -1, -1):
                stride_b[i] = acc
                acc *= int(shape[i])
        return cls(buffer, [int(s) for s in shape], stride_b, cplx, code)


@dataclass
class FFTDesc:
    axes: list = field(default_factory=list)
    scalingFactor: float = 1.0
    nthreads: int = 1  # accepted for signature parity; the GPU grid replaces pocketfft's thread pool
    forward: bool = True

    @classmethod
    def init(cls, axes: Sequence[int], forward: bool, scalingFactor: float = 1.0, nthreads: int = 1) -> "FFTDesc":
        return cls([int(a) for a in axes], float(scalingFactor), int(nthreads), bool(forward))

    def apply(self, descOut: DataDesc, descIn: DataDesc) -> None:
        apply(self, descOut, descIn)


@dataclass
class DCTDesc:
    """Descriptor of a discrete cosine transform (cpp_pocketfft/pocketfft.nim:151-156, 217-233).
    `sine=True` selects the DST family (pocketfft::dst, imported at pocketfft.nim:120-132 but not
    reachable through the reference's DCTDesc)."""
    axes: list = field(default_factory=list)
    dctType: int = 2
    scalingFactor: float = 1.0
    nthreads: int = 1
    ortho: bool = False
    sine: bool = False

    @classmethod
    def init(cls, axes: Sequence[int], dctType: int = 2, ortho: bool = False, scalingFactor: float = 1.0,
             nthreads: int = 1, sine: bool = False) -> "DCTDesc":
        if not 1 <= int(dctType) <= 4:
            raise ValueError("dctType must be in 1..4")  # range[1'i32..4'i32] in the Nim type
        return cls([int(a) for a in axes], int(dctType), float(scalingFactor), int(nthreads), bool(ortho), bool(sine))

    def apply(self, descOut: DataDesc, descIn: DataDesc) -> None:
        """pocketfft.nim:279-295."""
        L = _lib.lib()
        if descIn.complex or descOut.complex:
            raise TypeError("DCT/DST are real-to-real transforms")
        if descIn.dtype != descOut.dtype:
            raise TypeError("input and output precision differ")
        nd, na = len(descIn.shape), len(self.axes)
        if len(descIn.stride) != nd or len(descOut.stride) != nd:
            raise _lib.FFTError(-2, "stride dimension mismatch")
        if nd > _lib.MAX_DIMS or na > _lib.MAX_DIMS or any(a < 0 for a in self.axes):
            raise _lib.FFTError(-1, "bad axis number")
        fn = L.impulse_fft_dst if self.sine else L.impulse_fft_dct
        rc = fn(descIn.dtype, nd, (C.c_size_t * nd)(*descIn.shape), (C.c_ssize_t * nd)(*descIn.stride),
                (C.c_ssize_t * nd)(*descOut.stride), na, (C.c_size_t * na)(*self.axes), self.dctType,
                B.ptr(descIn.buf), B.ptr(descOut.buf), float(self.scalingFactor), int(self.ortho), int(self.nthreads),
                B.stream_of(descIn.buf, descOut.buf))
        _lib.check(rc)


def _r2r_real(fn_name, descOut: DataDesc, descIn: DataDesc, axes, fct, nthreads, extra=()):
    L = _lib.lib()
    if descIn.complex or descOut.complex:
        raise TypeError("real-to-real transform")
    if descIn.dtype != descOut.dtype:
        raise TypeError("input and output precision differ")
    nd, na = len(descIn.shape), len(axes)
    if len(descIn.stride) != nd or len(descOut.stride) != nd:
        raise _lib.FFTError(-2, "stride dimension mismatch")
    if nd > _lib.MAX_DIMS or na > _lib.MAX_DIMS or any(int(a) < 0 for a in axes):
        raise _lib.FFTError(-1, "bad axis number")
    rc = getattr(L, fn_name)(descIn.dtype, nd, (C.c_size_t * nd)(*descIn.shape), (C.c_ssize_t * nd)(*descIn.stride),
                             (C.c_ssize_t * nd)(*descOut.stride), na, (C.c_size_t * na)(*[int(a) for a in axes]), *extra,
                             B.ptr(descIn.buf), B.ptr(descOut.buf), float(fct), int(nthreads),
                             B.stream_of(descIn.buf, descOut.buf))
    _lib.check(rc)


def r2r_fftpack(descOut: DataDesc, descIn: DataDesc, axes, real2hermitian: bool, forward: bool, fct: float = 1.0,
                nthreads: int = 1) -> None:
    """`r2r_fftpack` (cpp_pocketfft/pocketfft.nim:71-82 -> pocketfft_hdronly.h:3392-3403): FFTPACK halfcomplex
    real transform along every listed axis, in the given order."""
    _r2r_real("impulse_fft_r2r_fftpack", descOut, descIn, axes, fct, nthreads, (int(bool(real2hermitian)), int(bool(forward))))


def r2r_separable_hartley(descOut: DataDesc, descIn: DataDesc, axes, fct: float = 1.0, nthreads: int = 1) -> None:
    """`r2r_separable_hartley` (pocketfft.nim:84-94 -> pocketfft_hdronly.h:3405-3415)."""
    _r2r_real("impulse_fft_r2r_separable_hartley", descOut, descIn, axes, fct, nthreads)


def r2r_genuine_hartley(descOut: DataDesc, descIn: DataDesc, axes, fct: float = 1.0, nthreads: int = 1) -> None:
    """`r2r_genuine_hartley` (pocketfft.nim:96-106 -> pocketfft_hdronly.h:3417-3445)."""
    _r2r_real("impulse_fft_r2r_genuine_hartley", descOut, descIn, axes, fct, nthreads)


_ARG_CACHE: dict = {}


def apply(fft, descOut: DataDesc, descIn: DataDesc) -> None:
    if isinstance(fft, DCTDesc):
        return fft.apply(descOut, descIn)
    L = _lib.lib()
    if descIn.dtype != descOut.dtype:
        raise TypeError("input and output precision differ")
    if descIn.complex and descOut.complex:
        fn, shape = L.impulse_fft_c2c, descIn.shape
    elif descOut.complex:
        fn, shape = L.impulse_fft_r2c, descIn.shape
    elif descIn.complex:
        fn, shape = L.impulse_fft_c2r, descOut.shape  # real shape (hdronly.h:3352-3360)
    else:
        raise TypeError("Not implemented")  # pocketfft.nim:277
    nd, na = len(shape), len(fft.axes)
    if len(descIn.stride) != nd or len(descOut.stride) != nd:
        raise _lib.FFTError(-2, "stride dimension mismatch")
    if nd > _lib.MAX_DIMS or na > _lib.MAX_DIMS:
        raise _lib.FFTError(-1, "too many dimensions")
    if any(a < 0 for a in fft.axes):
        raise _lib.FFTError(-1, "bad axis number")
    stream = B.stream_of(descIn.buf, descOut.buf)
    # the ctypes argument arrays of a (shape, strides, axes) combination are built once: a 20 us kernel (BASELINE
    # config 1) is otherwise bound by the host side of the call
    key = (tuple(shape), tuple(descIn.stride), tuple(descOut.stride), tuple(fft.axes))
    args = _ARG_CACHE.get(key)
    if args is None:
        if len(_ARG_CACHE) >= 256:
            _ARG_CACHE.clear()
        args = _ARG_CACHE[key] = ((C.c_size_t * nd)(*shape), (C.c_ssize_t * nd)(*descIn.stride),
                                  (C.c_ssize_t * nd)(*descOut.stride), (C.c_size_t * na)(*fft.axes))
    rc = fn(descIn.dtype, nd, args[0], args[1], args[2], na, args[3], int(fft.forward),
            B.ptr(descIn.buf), B.ptr(descOut.buf), float(fft.scalingFactor), int(fft.nthreads), stream)
    _lib.check(rc)
