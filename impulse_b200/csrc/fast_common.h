// Host-side helpers shared by the register-kernel translation units (fast_kernels.cu, fast3_kernels.cu).
#pragma once
#include <cuda_runtime.h>

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
// a zero-initialised pair of device words {next row, CTAs done} for one launch with dynamic row claims (fast_kernels.cu)
unsigned int *sched_slot();
int init_sched_slots();   // once per device, before the first launch (abi.cu: get_ctx)
// the three-pass register kernels (fast3_kernels.cu); returns cudaError_t, cudaErrorInvalidValue for other ids
int launch_fast3_job(const LineJob &job, int sm_count, void *stream);
// staged (TMA) variants, fast3t_kernels.cu: -1 = none for this job, launch the direct-load kernel
int launch_fast3_staged_job(const LineJob &job, int sm_count, void *stream);
int launch_fastblue_job(const LineJob &job, int sm_count, void *stream);   // fastblue_kernels.cu

}  // namespace impulse
