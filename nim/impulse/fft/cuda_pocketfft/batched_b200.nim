## Batched / axis / float32 / multi-GPU front end for Impulse's FFT module on the B200 engine (SURVEY 8(f) rank 4).
##
## NOT COMPILED IN THIS REPOSITORY (no Nim toolchain in the build container): like pocketfft_b200.nim this is the
## binding a maintainer drops in; every proc below is a few lines of shape / stride bookkeeping around ONE C-ABI call
## that the Python and C / C++ hosts of this repository exercise on the GPU (tests/test_gpu_parity.py,
## tests/test_gpu_dist_cabi.py, tests/cpp/).
##
## What it adds over the reference's C-backend surface:
##   * the Tensor overloads of impulse/fft/c_pocketfft/pocketfft_arraymancer.nim:33-94 flatten any-rank tensors to ONE
##     1-D transform (`data.size`, NA:37).  Here `fft(t, axis = ...)` transforms along one axis and batches the others —
##     one launch for the whole tensor instead of a Nim loop over rows (SURVEY A.4-9);
##   * float32 on the same API (the C backend is float64 only, NC:18-20);
##   * `fftRows` / `rfftRows` on raw buffers for callers without arraymancer;
##   * `DistFFT`: the single-process multi-GPU driver (`impulse_fft_dist_*`): batch sharding and the slab-decomposed fft2.
import std/complex
import arraymancer
import ./pocketfft_b200

# ---- raw C ABI used here (include/impulse_fft_b200.h) -------------------------------------------------------------
const
  kindC2C = 0.cint
  kindR2C = 1.cint
  kindC2R = 2.cint
  layoutHermitian = 0.cint
  maxDims = 8

type
  ImpulseFftDesc {.bycopy.} = object     ## impulse_fft_desc
    kind, dtype, realLayout, forward: int32
    ndim, naxes: uint32
    shape: array[maxDims, csize_t]
    strideIn, strideOut: array[maxDims, int]
    axes: array[maxDims, csize_t]
  ImpulseFftDist = pointer

proc impulse_fft_c2c(dtype: cint; ndim: csize_t; shape: ptr csize_t; strideIn, strideOut: ptr int; naxes: csize_t;
                     axes: ptr csize_t; forward: cint; dataIn, dataOut: pointer; fct: cdouble; nthreads: csize_t;
                     stream: pointer): cint {.importc, dynlib: libName, cdecl.}
proc impulse_fft_r2c(dtype: cint; ndim: csize_t; shape: ptr csize_t; strideIn, strideOut: ptr int; naxes: csize_t;
                     axes: ptr csize_t; forward: cint; dataIn, dataOut: pointer; fct: cdouble; nthreads: csize_t;
                     stream: pointer): cint {.importc, dynlib: libName, cdecl.}
proc impulse_fft_c2r(dtype: cint; ndim: csize_t; shape: ptr csize_t; strideIn, strideOut: ptr int; naxes: csize_t;
                     axes: ptr csize_t; forward: cint; dataIn, dataOut: pointer; fct: cdouble; nthreads: csize_t;
                     stream: pointer): cint {.importc, dynlib: libName, cdecl.}
proc impulse_fft_dist_create(dist: ptr ImpulseFftDist; mode: cint; desc: ptr ImpulseFftDesc; ndev: cint;
                             devices: ptr cint): cint {.importc, dynlib: libName, cdecl.}
proc impulse_fft_dist_execute(dist: ImpulseFftDist; dataIn, dataOut: pointer; fct: cdouble): cint {.importc, dynlib: libName, cdecl.}
proc impulse_fft_dist_shard(dist: ImpulseFftDist; index: cint; lo, hi: ptr csize_t): cint {.importc, dynlib: libName, cdecl.}
proc impulse_fft_dist_destroy(dist: ImpulseFftDist): cint {.importc, dynlib: libName, cdecl.}

proc check(rc: cint) =
  if rc != 0: raise newException(ValueError, "impulse_fft_b200: " & $impulse_fft_last_error())

template dtypeOf(T: typedesc): cint =
  when T is float32 or T is Complex32: 0.cint else: 1.cint

# ---- normalisation exactly as the reference (c_pocketfft/pocketfft.nim:189-204) -----------------------------------
type NormalizeKind* = enum nkBackward, nkOrtho, nkForward, nkCustom

func fctFor(kind: NormalizeKind; forward: bool; value: float; length: int): float =
  case kind
  of nkBackward: (if forward: 1.0 else: 1.0 / float(length))
  of nkForward: (if forward: 1.0 / float(length) else: 1.0)
  of nkOrtho: 1.0 / sqrt(float(length))
  of nkCustom: value

# ---- batched rows on raw buffers ----------------------------------------------------------------------------------
proc fftRows*[T: Complex32 | Complex64](data: ptr T; nrows, length: int; forward = true; fct = 1.0) =
  ## `nrows` contiguous rows of `length` complex points, in place, ONE launch (host or device memory).
  var shape = [csize_t nrows, csize_t length]
  var st = [length * sizeof(T), sizeof(T)]
  var axes = [csize_t 1]
  check impulse_fft_c2c(dtypeOf(T), 2, shape[0].addr, st[0].addr, st[0].addr, 1, axes[0].addr, cint(forward), data, data,
                        cdouble fct, 0, nil)

proc rfftRows*[T: float32 | float64](dataIn: ptr T; dataOut: ptr Complex[T]; nrows, length: int; fct = 1.0) =
  ## real rows -> `length div 2 + 1` complex bins per row (pocketfft::r2c layout, README.md:98)
  var shape = [csize_t nrows, csize_t length]
  var sIn = [length * sizeof(T), sizeof(T)]
  var sOut = [(length div 2 + 1) * sizeof(Complex[T]), sizeof(Complex[T])]
  var axes = [csize_t 1]
  check impulse_fft_r2c(dtypeOf(T), 2, shape[0].addr, sIn[0].addr, sOut[0].addr, 1, axes[0].addr, 1, dataIn, dataOut,
                        cdouble fct, 0, nil)

# ---- Tensor overloads with an axis (batched over every other axis) --------------------------------------------------
proc byteStrides[T](t: Tensor[T]): seq[int] =
  for s in t.strides: result.add s * sizeof(T)

proc fft*[T: Complex32 | Complex64](t: Tensor[T]; axis: int; forward = true; normalize = nkBackward;
                                    normValue = 1.0): Tensor[T] =
  ## Complex transform along `axis` of a tensor of any rank; the other axes are batch dimensions.
  ## (`fft(t)` without an axis keeps the reference's flatten-to-1-D behaviour, pocketfft_arraymancer.nim:37.)
  result = newTensorUninit[T](t.shape)
  var shape = newSeq[csize_t](t.rank)
  for i in 0 ..< t.rank: shape[i] = csize_t t.shape[i]
  var sIn = t.byteStrides
  var sOut = result.byteStrides
  var axes = [csize_t axis]
  let fct = fctFor(normalize, forward, normValue, t.shape[axis])
  check impulse_fft_c2c(dtypeOf(T), csize_t t.rank, shape[0].addr, sIn[0].addr, sOut[0].addr, 1, axes[0].addr,
                        cint(forward), t.unsafe_raw_offset.distinctBase, result.unsafe_raw_offset.distinctBase,
                        cdouble fct, 0, nil)

proc ifft*[T: Complex32 | Complex64](t: Tensor[T]; axis: int; normalize = nkBackward; normValue = 1.0): Tensor[T] =
  fft(t, axis, forward = false, normalize = normalize, normValue = normValue)

proc rfft*[T: float32 | float64](t: Tensor[T]; axis: int; normalize = nkBackward; normValue = 1.0): Tensor[Complex[T]] =
  ## Real transform along `axis` -> shape[axis] div 2 + 1 complex bins (the Hermitian half; `unpackFFT` is not needed).
  var oshape = t.shape.toSeq
  oshape[axis] = t.shape[axis] div 2 + 1
  result = newTensorUninit[Complex[T]](oshape)
  var shape = newSeq[csize_t](t.rank)
  for i in 0 ..< t.rank: shape[i] = csize_t t.shape[i]
  var sIn = t.byteStrides
  var sOut = result.byteStrides
  var axes = [csize_t axis]
  let fct = fctFor(normalize, true, normValue, t.shape[axis])
  check impulse_fft_r2c(dtypeOf(T), csize_t t.rank, shape[0].addr, sIn[0].addr, sOut[0].addr, 1, axes[0].addr, 1,
                        t.unsafe_raw_offset.distinctBase, result.unsafe_raw_offset.distinctBase, cdouble fct, 0, nil)

proc irfft*[T: float32 | float64](t: Tensor[Complex[T]]; axis, n: int; normalize = nkBackward; normValue = 1.0): Tensor[T] =
  ## Inverse of `rfft`: `n` real points along `axis` from n div 2 + 1 bins (pocketfft::c2r; shape = the REAL shape,
  ## pocketfft_hdronly.h:3352-3360 — the reference's FFTDesc.apply passes the wrong one, SURVEY A.4-2).
  var rshape = t.shape.toSeq
  rshape[axis] = n
  result = newTensorUninit[T](rshape)
  var shape = newSeq[csize_t](t.rank)
  for i in 0 ..< t.rank: shape[i] = csize_t rshape[i]
  var sIn = t.byteStrides
  var sOut = result.byteStrides
  var axes = [csize_t axis]
  let fct = fctFor(normalize, false, normValue, n)
  check impulse_fft_c2r(dtypeOf(T), csize_t t.rank, shape[0].addr, sIn[0].addr, sOut[0].addr, 1, axes[0].addr, 0,
                        t.unsafe_raw_offset.distinctBase, result.unsafe_raw_offset.distinctBase, cdouble fct, 0, nil)

# ---- several GPUs from one process ---------------------------------------------------------------------------------
type
  DistMode* = enum dmBatchShard = 0, dmSlab2D = 1
  DistFFT* = object
    ## `impulse_fft_dist_create/execute/destroy`: dimension 0 split over the devices (no communication), or one 2-D
    ## complex transform as row slabs whose column kernels read every device's rows over NVLink.
    handle: ImpulseFftDist

proc `=destroy`*(d: DistFFT) =
  if d.handle != nil: discard impulse_fft_dist_destroy(d.handle)
proc `=copy`*(a: var DistFFT; b: DistFFT) {.error: "a DistFFT owns device buffers: move it".}

proc initDistFFT*[T: Complex32 | Complex64](_: type T; mode: DistMode; shape: openArray[int]; axes: openArray[int];
                                            forward: bool; devices: openArray[int]): DistFFT =
  ## C-contiguous complex array of `shape`; BATCH_SHARD: axis 0 must not be in `axes`; SLAB_2D: shape.len == 2, axes = [0, 1].
  var d: ImpulseFftDesc
  d.kind = kindC2C; d.dtype = dtypeOf(T); d.realLayout = layoutHermitian; d.forward = int32(forward)
  d.ndim = uint32 shape.len; d.naxes = uint32 axes.len
  var acc = sizeof(T)
  for i in countdown(shape.len - 1, 0):
    d.shape[i] = csize_t shape[i]; d.strideIn[i] = acc; d.strideOut[i] = acc
    acc *= shape[i]
  for i, a in axes: d.axes[i] = csize_t a
  var devs = newSeq[cint](devices.len)
  for i, x in devices: devs[i] = cint x
  check impulse_fft_dist_create(result.handle.addr, cint(ord(mode)), d.addr, cint devs.len, devs[0].addr)

proc apply*[T](d: DistFFT; dataOut: var Tensor[T]; dataIn: Tensor[T]; fct = 1.0) =
  ## Whole arrays in host memory; every device stages and transforms its shard concurrently.
  check impulse_fft_dist_execute(d.handle, dataIn.unsafe_raw_offset.distinctBase, dataOut.unsafe_raw_offset.distinctBase,
                                 cdouble fct)

proc shard*(d: DistFFT; index: int): Slice[int] =
  var lo, hi: csize_t
  check impulse_fft_dist_shard(d.handle, cint index, lo.addr, hi.addr)
  int(lo) ..< int(hi)
