// Host-side helpers shared by the register-kernel translation units (fast_kernels.cu, fast3_kernels.cu).
#pragma once
#include <cuda_runtime.h>

#include <cstdlib>

#include "fft_types.h"

namespace impulse {

constexpr int kMaxDevices = 64;
inline int cur_dev() {
  int d = 0;
  cudaGetDevice(&d);
  return (d >= 0 && d < kMaxDevices) ? d : 0;
}
// function attributes (dynamic shared memory size, carve-out) are per device: remember where they were set
struct PerDeviceFlag {
  bool done[kMaxDevices] = {};
  bool &here() { return done[cur_dev()]; }
};
// Launch with programmatic stream serialisation (the kernel must call griddep_wait() before it touches row data):
// the prologue of launch N+1 overlaps the tail of launch N.  IMPULSE_FFT_PDL=0 launches plainly (A/B runs).
inline bool pdl_enabled() {
  static const int v = [] { const char *e = getenv("IMPULSE_FFT_PDL"); return e ? atoi(e) : 1; }();
  return v != 0;
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*k)(KArgs...), unsigned grid, unsigned block, size_t smem, cudaStream_t s, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid, 1, 1); cfg.blockDim = dim3(block, 1, 1); cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = at; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, k, KArgs(args)...);
}
// Grid of the row-claiming three-pass kernels: min(rows, resident CTAs).  IMPULSE_FFT_F3_GRID=1 evens out short
// batches (rows <= 8 x resident CTAs): with r = ceil(rows / cap) rows per CTA anyway, ceil(rows / r) CTAs finish at the
// same time as cap CTAs would and contend less for HBM (config 1: 1024 rows -> 256 CTAs of 4 rows instead of 296
// of 3 or 4).  Any grid >= 1 is correct: rows are claimed dynamically.
inline uint64_t f3_grid(uint64_t rows, uint64_t cap) {
  static const int mode = [] { const char *e = getenv("IMPULSE_FFT_F3_GRID"); return e ? atoi(e) : 0; }();
  if (rows <= cap) return rows;
  if (mode == 1 && rows <= 8 * cap) { const uint64_t r = (rows + cap - 1) / cap; return (rows + r - 1) / r; }
  return cap;
}
// a zero-initialised pair of device words {next row, CTAs done} for one launch with dynamic row claims (fast_kernels.cu)
unsigned int *sched_slot();
int init_sched_slots();   // once per device, before the first launch (abi.cu: get_ctx)
// the three-pass register kernels (fast3_kernels.cu); returns cudaError_t, cudaErrorInvalidValue for other ids
int launch_fast3_job(const LineJob &job, int sm_count, void *stream);
// staged (TMA) variants, fast3t_kernels.cu: -1 = none for this job, launch the direct-load kernel
int launch_fast3_staged_job(const LineJob &job, int sm_count, void *stream);
int launch_fastblue_job(const LineJob &job, int sm_count, void *stream);   // fastblue_kernels.cu
int launch_colconvw_job(const LineJob &job, int sm_count, void *stream);   // colconvw_kernels.cu

}  // namespace impulse
