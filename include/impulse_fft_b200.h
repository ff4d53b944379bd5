/* impulse_fft_b200.h — C ABI of libimpulse_fft_b200.so, the B200 (sm_100a) FFT engine that
 * drops in behind the FFT path of SciNim/impulse.
 *
 * Two groups of entry points:
 *
 *  (1) the ten pocketfft symbols the reference's C backend binds with importc
 *      (impulse/fft/c_pocketfft/pocketfft.h:18-32, bound at
 *      impulse/fft/c_pocketfft/pocketfft.nim:71-81) are declared in include/pocketfft.h and
 *      exported verbatim: a Nim build links this library instead of compiling pocketfft.c.
 *
 *  (2) the descriptor API below, which carries what the reference's C++ backend passes to
 *      pocketfft::c2c / r2c / c2r (impulse/fft/cpp_pocketfft/pocketfft_hdronly.h:3272-3390,
 *      called from FFTDesc.apply at impulse/fft/cpp_pocketfft/pocketfft.nim:235-277):
 *      shape, byte strides, axes, direction, scale factor — plus batching and the real-data
 *      layouts of the C backend (FFTPACK halfcomplex; full symmetric spectrum).
 *
 * Pointers may be device pointers (asynchronous on `stream`) or host pointers (staged through
 * the device in pipelined chunks; synchronous).  There is no CPU transform path: without a
 * CUDA device every call returns IMPULSE_FFT_ERR_NO_DEVICE.
 *
 * All functions return 0 on success or a negative impulse_fft_status; the message for the last
 * failure on the calling thread is available from impulse_fft_last_error().  These codes
 * replace the C++ exceptions of pocketfft_hdronly.h:446-476.
 */
#ifndef IMPULSE_FFT_B200_H
#define IMPULSE_FFT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IMPULSE_FFT_MAX_DIMS 8

typedef enum {
  IMPULSE_FFT_OK = 0,
  IMPULSE_FFT_ERR_INVALID = -1,     /* bad ndim / axis / null pointer / zero length       */
  IMPULSE_FFT_ERR_STRIDE = -2,      /* stride mismatch, non-element stride, misalignment   */
  IMPULSE_FFT_ERR_UNSUPPORTED = -3, /* valid request this build cannot run                 */
  IMPULSE_FFT_ERR_NOMEM = -4,
  IMPULSE_FFT_ERR_CUDA = -5,
  IMPULSE_FFT_ERR_NO_DEVICE = -6
} impulse_fft_status;

typedef enum { IMPULSE_FFT_C2C = 0, IMPULSE_FFT_R2C = 1, IMPULSE_FFT_C2R = 2 } impulse_fft_kind;
typedef enum { IMPULSE_FFT_F32 = 0, IMPULSE_FFT_F64 = 1 } impulse_fft_dtype;

/* layout of the complex side of a real transform */
typedef enum {
  IMPULSE_FFT_HERMITIAN = 0,   /* shape[axis]/2+1 complex bins (pocketfft_hdronly.h:3329)        */
  IMPULSE_FFT_HALFCOMPLEX = 1, /* FFTPACK packed reals, N per line (pocketfft.nim:228-238); 1 axis */
  IMPULSE_FFT_FULLSYM = 2      /* all N bins incl. conjugates = `symmetrize` (pocketfft.nim:160-171); r2c, 1 axis */
} impulse_fft_real_layout;

/* Transform descriptor.  `shape` is the array shape for C2C and the shape of the REAL array for
 * R2C / C2R (as pocketfft does, README_pocketfft.md:90-92).  Strides are in BYTES and may be
 * negative.  For R2C/C2R the real transform runs along axes[naxes-1]. */
typedef struct {
  int32_t kind;        /* impulse_fft_kind        */
  int32_t dtype;       /* impulse_fft_dtype       */
  int32_t real_layout; /* impulse_fft_real_layout */
  int32_t forward;     /* 1: exp(-2 pi i jk/N), 0: exp(+...)  (pocketfft.c:946-951) */
  uint32_t ndim;
  uint32_t naxes;
  size_t shape[IMPULSE_FFT_MAX_DIMS];
  ptrdiff_t stride_in[IMPULSE_FFT_MAX_DIMS];
  ptrdiff_t stride_out[IMPULSE_FFT_MAX_DIMS];
  size_t axes[IMPULSE_FFT_MAX_DIMS];
} impulse_fft_desc;

typedef struct impulse_fft_plan_s *impulse_fft_plan;

/* Plans are immutable after creation and may be executed concurrently from several threads /
 * streams (the contract of impulse/fft/c_pocketfft/README.md:32-36).  Twiddle, permutation and
 * Bluestein tables live in a per-device cache shared between plans. */
int impulse_fft_plan_create(impulse_fft_plan *out, const impulse_fft_desc *desc);
int impulse_fft_plan_destroy(impulse_fft_plan plan);

/* result = fct * transform(in).  `stream` is a cudaStream_t (NULL = default stream).
 * in == out is allowed for C2C with equal strides and for HALFCOMPLEX real transforms. */
int impulse_fft_execute(impulse_fft_plan plan, const void *in, void *out, double fct, void *stream);

/* One-shot forms with the argument list of pocketfft::c2c / r2c / c2r
 * (pocketfft_hdronly.h:3272-3275, 3334-3337, 3366-3369); plans are cached internally the way
 * get_plan does (pocketfft_hdronly.h:2655-2706).  nthreads is accepted and ignored. */
int impulse_fft_c2c(int dtype, size_t ndim, const size_t *shape, const ptrdiff_t *stride_in,
                    const ptrdiff_t *stride_out, size_t naxes, const size_t *axes, int forward,
                    const void *data_in, void *data_out, double fct, size_t nthreads, void *stream);
int impulse_fft_r2c(int dtype, size_t ndim, const size_t *shape_in, const ptrdiff_t *stride_in,
                    const ptrdiff_t *stride_out, size_t naxes, const size_t *axes, int forward,
                    const void *data_in, void *data_out, double fct, size_t nthreads, void *stream);
int impulse_fft_c2r(int dtype, size_t ndim, const size_t *shape_out, const ptrdiff_t *stride_in,
                    const ptrdiff_t *stride_out, size_t naxes, const size_t *axes, int forward,
                    const void *data_in, void *data_out, double fct, size_t nthreads, void *stream);

/* FFTPACK halfcomplex real transforms on every listed axis, and the two Hartley transforms: the
 * flattened argument lists of pocketfft::r2r_fftpack / r2r_separable_hartley / r2r_genuine_hartley
 * (pocketfft_hdronly.h:3392-3445; imported at impulse/fft/cpp_pocketfft/pocketfft.nim:71-106).  All arrays
 * are real with the same shape; strides in bytes.  Hartley sign convention as the reference: out[k] =
 * Re F[k] + Im F[k] with F the FORWARD transform.  r2r_fftpack follows the vendored engine exactly: the
 * direction of the real transform is chosen by `forward` (true: real -> halfcomplex, false: halfcomplex ->
 * real); when real2hermitian != forward, elements 2, 4, 6, ... along the axis change sign (on the output
 * for real2hermitian, on the input otherwise), as pocketfft_hdronly.h:3134-3140 computes it. */
int impulse_fft_r2r_fftpack(int dtype, size_t ndim, const size_t *shape, const ptrdiff_t *stride_in,
                            const ptrdiff_t *stride_out, size_t naxes, const size_t *axes, int real2hermitian, int forward,
                            const void *data_in, void *data_out, double fct, size_t nthreads, void *stream);
int impulse_fft_r2r_separable_hartley(int dtype, size_t ndim, const size_t *shape, const ptrdiff_t *stride_in,
                                      const ptrdiff_t *stride_out, size_t naxes, const size_t *axes, const void *data_in,
                                      void *data_out, double fct, size_t nthreads, void *stream);
int impulse_fft_r2r_genuine_hartley(int dtype, size_t ndim, const size_t *shape, const ptrdiff_t *stride_in,
                                    const ptrdiff_t *stride_out, size_t naxes, const size_t *axes, const void *data_in,
                                    void *data_out, double fct, size_t nthreads, void *stream);

/* impulse_fft_c2c with a pointwise multiply fused into the store of the last pass: the output element at
 * element offset o of the output array (positive strides; o counts padding between rows too, so a padded array
 * takes a multiplier padded alike) is multiplied by mul[o % mul_elems] — mul_elems = the size of one image
 * broadcasts one filter spectrum over a batch.  This is the FFT -> multiply half of an FFT convolution without
 * the extra pass over the spectrum.  Device pointers. */
int impulse_fft_c2c_mul(int dtype, size_t ndim, const size_t *shape, const ptrdiff_t *stride_in,
                        const ptrdiff_t *stride_out, size_t naxes, const size_t *axes, int forward, const void *data_in,
                        void *data_out, double fct, const void *mul, size_t mul_elems, void *stream);

/* Convolution along ONE axis of a complex array in the frequency domain:
 *   out = fct * IFFT_axis( FFT_axis(in) .* mul[o % mul_elems] ),   o = element offset in the output layout.
 * A strided power-of-two axis of 512 ... 4096 points over adjacent lines (unit stride across the lines) runs as ONE
 * pass with the whole axis in shared memory (colconvw_kernel); with equal, dense in/out layouts and a strided axis
 * whose length splits into register-kernel factors (1024 ... 16384) it runs as three passes; in both the spectrum is
 * never written to memory.  Otherwise it is impulse_fft_c2c_mul followed by the inverse transform.
 * Device pointers; in place allowed. */
int impulse_fft_convolve_axis(int dtype, size_t ndim, const size_t *shape, const ptrdiff_t *stride_in,
                              const ptrdiff_t *stride_out, size_t axis, const void *data_in, void *data_out, double fct,
                              const void *mul, size_t mul_elems, void *stream);

/* Discrete cosine / sine transforms of type 1..4 over `axes`, real to real, with the argument list of
 * pocketfft::dct / pocketfft::dst (pocketfft_hdronly.h:3284-3318; FFTW's REDFT/RODFT definitions;
 * `ortho` as documented at README_pocketfft.md:220-241).  Called by DCTDesc.apply
 * (impulse/fft/cpp_pocketfft/pocketfft.nim:279-295).  in == out allowed with equal strides. */
int impulse_fft_dct(int dtype, size_t ndim, const size_t *shape, const ptrdiff_t *stride_in,
                    const ptrdiff_t *stride_out, size_t naxes, const size_t *axes, int type, const void *data_in,
                    void *data_out, double fct, int ortho, size_t nthreads, void *stream);
int impulse_fft_dst(int dtype, size_t ndim, const size_t *shape, const ptrdiff_t *stride_in,
                    const ptrdiff_t *stride_out, size_t naxes, const size_t *axes, int type, const void *data_in,
                    void *data_out, double fct, int ortho, size_t nthreads, void *stream);

/* Batched forms of the C backend's in-place row transforms: `nrows` contiguous rows of
 * `length` complex (cfft) or real (rfft, FFTPACK halfcomplex) doubles, device or host memory.
 * This is what a caller looping fft() over rows should call instead. */
int impulse_fft_cfft_rows(double *data, size_t nrows, size_t length, int forward, double fct, void *stream);
int impulse_fft_rfft_rows(double *data, size_t nrows, size_t length, int forward, double fct, void *stream);

/* Pointwise step of FFT-based filtering (r2c -> multiply -> c2r): out[b][i] = a[b][i]*filter[i]*scale
 * over complex arrays, i < n_inner, b < n_batch.  Device pointers; out may alias a. */
int impulse_fft_cmul(int dtype, const void *a, const void *filter, void *out, size_t n_inner, size_t n_batch,
                     double scale, void *stream);

/* Batched out-of-place transpose of complex matrices, out[b][c][r] = in[b][r][c]; leading dimensions in
 * elements; batch b starts at b*rows*ld_in / b*cols*ld_out.  Used to re-block slabs around the all-to-all
 * of the multi-GPU 2-D transform.  Device pointers. */
int impulse_fft_transpose(int dtype, const void *in, void *out, size_t rows, size_t cols, size_t ld_in,
                          size_t ld_out, size_t batch, void *stream);

/* Batched strided copy of complex matrices, out[b][r][c] = in[b][r][c], leading dimensions and batch
 * strides in elements.  Packs the per-peer column blocks of a row slab into contiguous send buffers
 * ahead of the all-to-all of the multi-GPU 2-D transform.  Device pointers. */
int impulse_fft_copy2d(int dtype, const void *in, void *out, size_t rows, size_t cols, size_t ld_in, size_t ld_out,
                       size_t batch, size_t bs_in, size_t bs_out, void *stream);

/* Column pass of a 2-D complex transform whose ROWS are distributed over `nparts` allocations of
 * `rows_per_part` rows each (leading dimension `ld_part` elements) — typically the row slabs of the other
 * GPUs of the box, mapped through CUDA IPC: out[k][c] = fct * sum_r part[r / rpp][(r % rpp)*ld_part + col0 + c]
 * * exp(-+2 pi i r k / R), R = nparts*rows_per_part, c < ncols, written with leading dimension ld_out.
 * The kernel loads straight from the peers' memory over NVLink (no pack, no all-to-all staging): the
 * exchange is fused into the load phase of the transform.  nparts <= 8.  Device pointers. */
int impulse_fft_cols_from_parts(int dtype, size_t nparts, const void *const *parts, size_t rows_per_part, size_t ld_part,
                                size_t col0, size_t ncols, void *out, size_t ld_out, int forward, double fct, void *stream);

/* The exchange of the slab 2-D transform as a copy: the column block [col0, col0 + ncols) of `nparts` row slabs
 * (rows_per_part rows each, leading dimension ld_part; the peers' slabs mapped through CUDA IPC) is gathered into the
 * local array out[nparts * rows_per_part][ld_out].  A small grid (`ctas`, 0 = 32) with many loads in flight per thread
 * keeps NVLink busy and leaves the other SMs to the column transform of the previous column chunk, which is how
 * SlabFFT2P2P(pull_chunks = J) overlaps exchange and arithmetic.  Offsets and leading dimensions in elements, multiples
 * of 16 bytes.  Reference analogue: the per-axis line gather of general_nd, pocketfft_hdronly.h:2955-3004. */
int impulse_fft_gather_parts(int dtype, size_t nparts, const void *const *parts, size_t rows_per_part, size_t ld_part,
                             size_t col0, size_t ncols, void *out, size_t ld_out, int ctas, void *stream);

/* Device buffers that other PROCESSES of the same box can map (CUDA IPC), for the row slabs of the
 * multi-GPU 2-D transform.  alloc: cudaMalloc + cudaIpcGetMemHandle (handle = 64 opaque bytes to send to the
 * peers).  open: maps a peer's buffer into the CURRENT device's address space with peer access enabled
 * (cudaIpcOpenMemHandle, lazy peer access) so that kernels of this device can load from it over NVLink. */
int impulse_fft_ipc_alloc(size_t bytes, void **ptr, void *handle64);
int impulse_fft_ipc_free(void *ptr);
int impulse_fft_ipc_open(const void *handle64, void **ptr);
int impulse_fft_ipc_close(void *ptr);

/* Let kernels of the CURRENT device dereference memory that lives on `peer_device` (cudaDeviceEnablePeerAccess;
 * "already enabled" is not an error).  Needed once per peer before impulse_fft_cols_from_parts is given
 * pointers into other GPUs' memory. */
int impulse_fft_enable_peer_access(int peer_device);

/* Single-process multi-GPU driver over the devices of one NVSwitch box.  Two partitionings:
 *   BATCH_SHARD  any transform whose dimension 0 is a batch dimension (not in `axes`): its shape[0] entries are split
 *                contiguously over the devices, device g holding [g*B/G, (g+1)*B/G); no communication.
 *   SLAB_2D      one 2-D complex transform over both axes, split into row slabs; row FFTs locally, then every device's
 *                column kernels load their column block out of all row slabs over NVLink (peer access; no pack, no
 *                all-to-all buffer).  Rows and columns must be divisible by ndev.
 * The reference's counterpart is general_nd splitting the lines of one call over its thread pool
 * (pocketfft_hdronly.h:3012-3050); `desc` is what impulse_fft_plan_create takes, for the WHOLE array.
 *   execute        whole arrays in HOST memory; every device stages its shard (one host thread per device for
 *                  BATCH_SHARD).  SLAB_2D returns the result in natural [R, C] layout.  Synchronous.
 *   execute_parts  per-device DEVICE pointers: in_parts[g] = the shard of device g (strides of `desc`); out_parts[g] =
 *                  the transformed shard (BATCH_SHARD, strides of `desc`) or the dense column slab [R, C/ndev] holding
 *                  columns [g*C/ndev, (g+1)*C/ndev) (SLAB_2D).  Returns after every device has finished. */
typedef struct impulse_fft_dist_s *impulse_fft_dist;
typedef enum { IMPULSE_FFT_DIST_BATCH_SHARD = 0, IMPULSE_FFT_DIST_SLAB_2D = 1 } impulse_fft_dist_mode;
int impulse_fft_dist_create(impulse_fft_dist *out, int mode, const impulse_fft_desc *desc, int ndev, const int *devices);
int impulse_fft_dist_execute(impulse_fft_dist dist, const void *in, void *out, double fct);
int impulse_fft_dist_execute_parts(impulse_fft_dist dist, const void *const *in_parts, void *const *out_parts, double fct);
int impulse_fft_dist_shard(impulse_fft_dist dist, int index, size_t *lo, size_t *hi);   /* rows of dimension 0 held by devices[index] */
int impulse_fft_dist_destroy(impulse_fft_dist dist);

/* Host-side placement for host-pointer calls: binds the CALLING thread (and threads it creates later) to the CPUs of
 * the NUMA node that `device`'s PCIe root belongs to (sysfs), so that pinned buffers allocated afterwards are local to
 * the GPU that will DMA them.  With one process per GPU this keeps 8 ranks from sharing node 0.  *numa_node receives
 * the node, or -1 when the box exposes no NUMA information (nothing is changed then; not an error). */
int impulse_fft_bind_host_to_device(int device, int *numa_node);

/* Introspection (tests, benchmarks). */
typedef struct {
  uint32_t n_steps;        /* kernel launches per execute                       */
  uint32_t n_fft;          /* shared-memory FFT length of the first step        */
  uint32_t bluestein;      /* first step runs Bluestein                         */
  uint32_t lines_per_cta;  /* first step                                        */
  uint32_t threads;        /* first step                                        */
  uint32_t smem_bytes;     /* first step                                        */
  uint64_t tmp_bytes;      /* device scratch held by the plan                   */
  uint32_t n_radices;
  uint32_t radices[32];    /* first step's radix schedule                       */
} impulse_fft_plan_info;
int impulse_fft_plan_get_info(impulse_fft_plan plan, impulse_fft_plan_info *info);

/* number of kernels launched by this library since load (bench.py's gpu_launches) */
uint64_t impulse_fft_launch_count(void);

/* name of the kernel most recently launched by the calling thread ("" if none) */
const char *impulse_fft_last_kernel(void);

const char *impulse_fft_last_error(void);
const char *impulse_fft_version(void);

#ifdef __cplusplus
}
#endif
#endif /* IMPULSE_FFT_B200_H */
