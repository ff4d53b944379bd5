// Launcher interface between the ABI layer (abi.cu) and the kernels (fft_kernels.cu).
#pragma once
#include <cstddef>
#include <cstdint>

#include "fft_types.h"

namespace impulse {
// name of the kernel most recently launched by this thread (introspection for bench/tests)
extern thread_local const char *g_last_kernel;
// raise the dynamic shared-memory limit of every kernel (once per device); returns cudaError_t
int configure_kernels(size_t max_dyn_smem);
int init_sched_slots();   // row-scheduler words of the register kernels (fast_kernels.cu); once per device
// enqueue one LineJob on `stream` (a cudaStream_t); returns cudaError_t
int launch_line_job(const LineJob &job, int threads, size_t smem_bytes, uint64_t n_tiles, void *stream);
// specialised kernels selected by LineJob::fast_id (fast_kernels.cu); returns cudaError_t
int launch_fast_job(const LineJob &job, int sm_count, void *stream);
// out[b][i] = a[b][i] * f[i] * scale over complex arrays (i < n_inner, b < n_batch)
int launch_cmul(int dtype, const void *a, const void *f, void *out, size_t n_inner, size_t n_batch, double scale,
                int sm_count, void *stream);
// out[b][c][r] = in[b][r][c], complex elements, leading dimensions in elements
int launch_transpose(int dtype, const void *in, void *out, size_t rows, size_t cols, size_t ld_in, size_t ld_out,
                     size_t batch, int sm_count, void *stream);
// out[b][r][c] = in[b][r][c] for complex matrices with independent leading dimensions / batch strides
int launch_gather_parts(const void *const *parts, uint32_t nparts, uint32_t rpp, uint64_t ld_part16, uint64_t col0_16,
                        uint32_t ncols16, void *out, uint64_t ld_out16, int ctas, void *stream);
int launch_copy2d(int dtype, const void *in, void *out, size_t rows, size_t cols, size_t ld_in, size_t ld_out,
                  size_t batch, size_t bs_in, size_t bs_out, int sm_count, void *stream);
// both launches of a split column transform as one persistent kernel (colfuse_kernels.cu).  The two jobs carry
// their final pointers / fct; `scratch` = colfuse_scratch_bytes() bytes of device memory (ring + control words; the
// control words must be zero).  Returns cudaError_t, or -1 when there is no fused kernel for the pair.
size_t colfuse_scratch_bytes(const LineJob &a, const LineJob &b, uint32_t tiles, size_t *ctrl_off, size_t *ctrl_bytes);
int launch_colfuse_pair(const LineJob &a, const LineJob &b, uint32_t tiles, uint32_t g0n, void *scratch, int sm_count, void *stream);
bool colfuse_pair_supported(uint32_t fast_id_a, uint32_t fast_id_b);
// elementwise conversion / chirp pass around a long transform (AuxJob)
int launch_aux(const AuxJob &job, int sm_count, void *stream);
// genuine Hartley fold of a contiguous half spectrum into the real output (CombineJob)
int launch_hartley_combine(const CombineJob &job, int sm_count, void *stream);
}  // namespace impulse
