## B200 engine binding for Impulse's FFT module — `importc` veneer over libimpulse_fft_b200.so.
##
## NOT COMPILED IN THIS REPOSITORY: the build container has no Nim toolchain.  The tested
## boundary is the C ABI (include/impulse_fft_b200.h, include/pocketfft.h); this file is the
## thin binding a maintainer drops into `impulse/fft/` (see INTEGRATION.md).
##
## Usage:   nim c -d:impulseCuda --passL:"-L<dir> -limpulse_fft_b200" yourprog.nim
##
## It provides both API surfaces of the reference:
##   * the C backend symbols (`make_rfft_plan`, `rfft_forward`, ...) with the exact signatures of
##     impulse/fft/c_pocketfft/pocketfft.nim:71-81, so `fft`/`ifft`/`rfft`/`rfft_packed`,
##     `unpackFFT`, `symmetrize` and the Tensor overloads keep working unchanged;
##   * the C++ backend's `DataDesc` / `FFTDesc` / `apply`
##     (impulse/fft/cpp_pocketfft/pocketfft.nim:137-149,158-215,235-277), now without needing
##     `--backend:cpp`, routed to `impulse_fft_c2c/r2c/c2r`.
import std/complex

const libName* {.strdefine.} = "libimpulse_fft_b200.so"

# ---- C backend: the ten pocketfft symbols (include/pocketfft.h) ---------------------------------
type
  rfft_plan* = pointer
  cfft_plan* = pointer

proc make_cfft_plan*(length: csize_t): cfft_plan {.importc, dynlib: libName, cdecl.}
proc destroy_cfft_plan*(plan: cfft_plan) {.importc, dynlib: libName, cdecl.}
proc cfft_backward*(plan: cfft_plan; c: ptr Complex64; fct: cdouble): cint {.importc, dynlib: libName, cdecl.}
proc cfft_forward*(plan: cfft_plan; c: ptr Complex64; fct: cdouble): cint {.importc, dynlib: libName, cdecl.}
proc cfft_length*(plan: cfft_plan): csize_t {.importc, dynlib: libName, cdecl.}
proc make_rfft_plan*(length: csize_t): rfft_plan {.importc, dynlib: libName, cdecl.}
proc destroy_rfft_plan*(plan: rfft_plan) {.importc, dynlib: libName, cdecl.}
proc rfft_backward*(plan: rfft_plan; c: ptr cdouble; fct: cdouble): cint {.importc, dynlib: libName, cdecl.}
proc rfft_forward*(plan: rfft_plan; c: ptr cdouble; fct: cdouble): cint {.importc, dynlib: libName, cdecl.}
proc rfft_length*(plan: rfft_plan): csize_t {.importc, dynlib: libName, cdecl.}

# batched rows (what a caller looping `fft` over rows should use; SURVEY A.4-9)
proc impulse_fft_cfft_rows*(data: ptr Complex64; nrows, length: csize_t; forward: cint; fct: cdouble;
                            stream: pointer): cint {.importc, dynlib: libName, cdecl.}
proc impulse_fft_rfft_rows*(data: ptr cdouble; nrows, length: csize_t; forward: cint; fct: cdouble;
                            stream: pointer): cint {.importc, dynlib: libName, cdecl.}
proc impulse_fft_last_error*(): cstring {.importc, dynlib: libName, cdecl.}

# ---- C++ backend surface: DataDesc / FFTDesc / apply -------------------------------------------
proc impulse_fft_c2c(dtype: cint; ndim: csize_t; shape: ptr csize_t; strideIn, strideOut: ptr int;
                     naxes: csize_t; axes: ptr csize_t; forward: cint; dataIn, dataOut: pointer;
                     fct: cdouble; nthreads: csize_t; stream: pointer): cint {.importc, dynlib: libName, cdecl.}
proc impulse_fft_r2c(dtype: cint; ndim: csize_t; shape: ptr csize_t; strideIn, strideOut: ptr int;
                     naxes: csize_t; axes: ptr csize_t; forward: cint; dataIn, dataOut: pointer;
                     fct: cdouble; nthreads: csize_t; stream: pointer): cint {.importc, dynlib: libName, cdecl.}
proc impulse_fft_c2r(dtype: cint; ndim: csize_t; shape: ptr csize_t; strideIn, strideOut: ptr int;
                     naxes: csize_t; axes: ptr csize_t; forward: cint; dataIn, dataOut: pointer;
                     fct: cdouble; nthreads: csize_t; stream: pointer): cint {.importc, dynlib: libName, cdecl.}

proc impulse_fft_dct(dtype: cint; ndim: csize_t; shape: ptr csize_t; strideIn, strideOut: ptr int;
                     naxes: csize_t; axes: ptr csize_t; dctType: cint; dataIn, dataOut: pointer;
                     fct: cdouble; ortho: cint; nthreads: csize_t; stream: pointer): cint {.importc, dynlib: libName, cdecl.}
proc impulse_fft_dst(dtype: cint; ndim: csize_t; shape: ptr csize_t; strideIn, strideOut: ptr int;
                     naxes: csize_t; axes: ptr csize_t; dstType: cint; dataIn, dataOut: pointer;
                     fct: cdouble; ortho: cint; nthreads: csize_t; stream: pointer): cint {.importc, dynlib: libName, cdecl.}
proc impulse_fft_r2r_fftpack(dtype: cint; ndim: csize_t; shape: ptr csize_t; strideIn, strideOut: ptr int;
                             naxes: csize_t; axes: ptr csize_t; real2hermitian, forward: cint; dataIn, dataOut: pointer;
                             fct: cdouble; nthreads: csize_t; stream: pointer): cint {.importc, dynlib: libName, cdecl.}
proc impulse_fft_r2r_separable_hartley(dtype: cint; ndim: csize_t; shape: ptr csize_t; strideIn, strideOut: ptr int;
                                       naxes: csize_t; axes: ptr csize_t; dataIn, dataOut: pointer;
                                       fct: cdouble; nthreads: csize_t; stream: pointer): cint {.importc, dynlib: libName, cdecl.}
proc impulse_fft_r2r_genuine_hartley(dtype: cint; ndim: csize_t; shape: ptr csize_t; strideIn, strideOut: ptr int;
                                     naxes: csize_t; axes: ptr csize_t; dataIn, dataOut: pointer;
                                     fct: cdouble; nthreads: csize_t; stream: pointer): cint {.importc, dynlib: libName, cdecl.}

type
  DataDesc*[T] = object
    ## Descriptor of the data used in or out of the FFT (cpp_pocketfft/pocketfft.nim:137-142)
    shape*: seq[csize_t]
    stride*: seq[int]          ## bytes
    buf*: ptr UncheckedArray[T]

  FFTDesc*[T] = object
    ## Descriptor of the FFT (cpp_pocketfft/pocketfft.nim:144-149)
    axes*: seq[csize_t]
    scalingFactor*: T
    nthreads*: uint
    forward*: bool

func init*[T](_: type DataDesc[T], buffer: ptr T or ptr UncheckedArray[T],
              shape, stride: distinct openArray[SomeInteger]): DataDesc[T] =
  ## stride in elements of T (cpp_pocketfft/pocketfft.nim:158-176); `==`, not `=` (SURVEY A.4-3)
  assert shape.len == stride.len
  assert not buffer.isNil
  for i in 0 ..< shape.len:
    result.shape.add csize_t(shape[i])
    result.stride.add int(stride[i]) * sizeof(T)
  result.buf = cast[ptr UncheckedArray[T]](buffer)

func init*[T](_: type DataDesc[T], buffer: ptr T or ptr UncheckedArray[T],
              shape: openArray[SomeInteger]): DataDesc[T] =
  ## C-contiguous data (cpp_pocketfft/pocketfft.nim:178-199)
  assert not buffer.isNil
  result.stride = newSeq[int](shape.len)
  var accum = sizeof(T)
  for i in countdown(shape.len - 1, 0):
    result.stride[i] = accum
    accum *= int(shape[i])
  for s in shape: result.shape.add csize_t(s)
  result.buf = cast[ptr UncheckedArray[T]](buffer)

func init*[T](_: type FFTDesc[T], axes: varargs[int], forward: bool, scalingFactor: T = 1,
              nthreads = 1): FFTDesc[T] =
  for a in axes: result.axes.add csize_t(a)
  result.scalingFactor = scalingFactor
  result.forward = forward
  result.nthreads = uint nthreads

template dtypeCode(T: typedesc): cint =
  when T is float32: 0.cint else: 1.cint

proc check(rc: cint) =
  if rc != 0:
    raise newException(ValueError, "impulse_fft_b200: " & $impulse_fft_last_error())

proc apply*[In, Out](fft: FFTDesc, descOut: var DataDesc[Out], descIn: DataDesc[In]) =
  ## Same dispatch as cpp_pocketfft/pocketfft.nim:235-277.  The c2r branch passes the REAL
  ## (output) shape, which is what pocketfft::c2r expects (fixes SURVEY A.4-2).
  var axes = fft.axes
  when In is Complex and Out is Complex:
    var shape = descIn.shape
    check impulse_fft_c2c(dtypeCode(In.T), csize_t shape.len, shape[0].addr, descIn.stride[0].unsafeAddr,
                          descOut.stride[0].addr, csize_t axes.len, axes[0].addr, cint(fft.forward),
                          descIn.buf, descOut.buf, cdouble(fft.scalingFactor), csize_t(fft.nthreads), nil)
  elif Out is Complex:
    var shape = descIn.shape
    check impulse_fft_r2c(dtypeCode(In), csize_t shape.len, shape[0].addr, descIn.stride[0].unsafeAddr,
                          descOut.stride[0].addr, csize_t axes.len, axes[0].addr, cint(fft.forward),
                          descIn.buf, descOut.buf, cdouble(fft.scalingFactor), csize_t(fft.nthreads), nil)
  elif In is Complex:
    var shape = descOut.shape
    check impulse_fft_c2r(dtypeCode(Out), csize_t shape.len, shape[0].addr, descIn.stride[0].unsafeAddr,
                          descOut.stride[0].addr, csize_t axes.len, axes[0].addr, cint(fft.forward),
                          descIn.buf, descOut.buf, cdouble(fft.scalingFactor), csize_t(fft.nthreads), nil)
  else:
    {.error: "Not implemented".}

type
  DCTDesc*[T] = object
    ## cpp_pocketfft/pocketfft.nim:151-156
    axes*: seq[csize_t]
    dctType*: range[1'i32..4'i32]
    scalingFactor*: T
    nthreads*: uint
    ortho*: bool

func init*[T](_: type DCTDesc[T], axes: varargs[int], dctType: range[1'i32..4'i32] = 2'i32, ortho = false,
              scalingFactor: T = 1, nthreads = 1): DCTDesc[T] =
  ## cpp_pocketfft/pocketfft.nim:217-233
  for a in axes: result.axes.add csize_t(a)
  result.dctType = dctType
  result.ortho = ortho
  result.scalingFactor = scalingFactor
  result.nthreads = uint nthreads

proc apply*[T](dct: DCTDesc[T], descOut: var DataDesc[T], descIn: DataDesc[T]) =
  ## cpp_pocketfft/pocketfft.nim:279-295
  var axes = dct.axes
  var shape = descIn.shape
  check impulse_fft_dct(dtypeCode(T), csize_t shape.len, shape[0].addr, descIn.stride[0].unsafeAddr,
                        descOut.stride[0].addr, csize_t axes.len, axes[0].addr, cint(dct.dctType),
                        descIn.buf, descOut.buf, cdouble(dct.scalingFactor), cint(dct.ortho), csize_t(dct.nthreads), nil)

proc r2r_fftpack*[T](descOut: var DataDesc[T], descIn: DataDesc[T], axes: varargs[int],
                     real2hermitian, forward: bool, fct: T = 1) =
  ## cpp_pocketfft/pocketfft.nim:71-82 (imported there, reachable here)
  var ax = newSeq[csize_t](axes.len)
  for i, a in axes: ax[i] = csize_t a
  var shape = descIn.shape
  check impulse_fft_r2r_fftpack(dtypeCode(T), csize_t shape.len, shape[0].addr, descIn.stride[0].unsafeAddr,
                                descOut.stride[0].addr, csize_t ax.len, ax[0].addr, cint(real2hermitian), cint(forward),
                                descIn.buf, descOut.buf, cdouble(fct), 1, nil)

proc r2r_hartley*[T](descOut: var DataDesc[T], descIn: DataDesc[T], axes: varargs[int], genuine = false, fct: T = 1) =
  ## cpp_pocketfft/pocketfft.nim:84-106 (the reference's import of the separable variant has a stray `forward`)
  var ax = newSeq[csize_t](axes.len)
  for i, a in axes: ax[i] = csize_t a
  var shape = descIn.shape
  if genuine:
    check impulse_fft_r2r_genuine_hartley(dtypeCode(T), csize_t shape.len, shape[0].addr, descIn.stride[0].unsafeAddr,
                                          descOut.stride[0].addr, csize_t ax.len, ax[0].addr, descIn.buf, descOut.buf,
                                          cdouble(fct), 1, nil)
  else:
    check impulse_fft_r2r_separable_hartley(dtypeCode(T), csize_t shape.len, shape[0].addr, descIn.stride[0].unsafeAddr,
                                            descOut.stride[0].addr, csize_t ax.len, ax[0].addr, descIn.buf, descOut.buf,
                                            cdouble(fct), 1, nil)
