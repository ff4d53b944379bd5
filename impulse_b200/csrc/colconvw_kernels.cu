// Launcher of the whole-axis convolution kernel (colconvw_kernel, colconvw_device.cuh): FFT -> multiply -> inverse FFT
// along a strided axis in one pass over the data, the axis resident in shared memory.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>

#include "colconvw_device.cuh"
#include "fast_common.h"
#include "fft_kernels.h"

namespace impulse {

namespace {
template <typename T, int R1, int R2, int R3, int W, int LP, int TT, bool PLAIN = false>
int launch_colconvw(const LineJob &J, int sm_count, cudaStream_t s) {
  constexpr int N = R1 * R2 * R3;
  const size_t smem = sizeof(cx<T>) * ((size_t)N * W + (size_t)N + (size_t)R2 * R3) + 16;   // + the tensor-memory address slot
  if (!J.f3_tw1 || !J.f3_tw2) return (int)cudaErrorInvalidValue;
  if (!PLAIN && (!J.umul || !J.umul_mod)) return (int)cudaErrorInvalidValue;
  const bool bwd = PLAIN && (J.flags & F_CONJ_SEQ) != 0;
  // the LP lines of a thread move as one vector when every address involved is a multiple of the vector
  constexpr uint64_t VB = LP * sizeof(cx<T>) >= 16 ? 16 : 8, VE = VB / sizeof(cx<T>) ? VB / sizeof(cx<T>) : 1;
  auto mult = [&](int64_t v) { return v % (int64_t)VE == 0; };
  const bool gv = (uintptr_t)J.in % VB == 0 && (uintptr_t)J.out % VB == 0 && (PLAIN || (uintptr_t)J.umul % VB == 0) && mult(J.es_in) &&
                  mult(J.es_out) && mult(J.bs_in[1]) && mult(J.bs_in[2]) && mult(J.bs_out[1]) && mult(J.bs_out[2]) &&
                  (PLAIN || J.umul_mod % VE == 0);
  // The next tile staged in tensor memory while the current one is transformed (aligned arrays).  Measured
  // (profiles/r02_ab_convw_tmem.txt): complex128, whose 256-thread CTAs have registers to spare for the points in flight,
  // gains 1-3 % on the convolution and 3-13 % on the plain transform; complex64 (512 threads, 128 registers: the staged
  // points spill) loses 1-8 %.  IMPULSE_FFT_CONVW_TMEM = 0: off, 1 (default): complex128, 2: both.
  static const int tm_env = [] { const char *e = getenv("IMPULSE_FFT_CONVW_TMEM"); return e ? atoi(e) : 1; }();
  const bool tm = gv && (tm_env >= 2 || (tm_env == 1 && sizeof(T) == 8));
  typedef void (*kern_t)(const LineJob);
  kern_t k;
  if constexpr (PLAIN) {
    k = bwd ? (gv ? (tm ? colconvw_kernel<T, R1, R2, R3, W, LP, TT, true, CW_BWD, true> : colconvw_kernel<T, R1, R2, R3, W, LP, TT, true, CW_BWD>)
                  : colconvw_kernel<T, R1, R2, R3, W, LP, TT, false, CW_BWD>)
            : (gv ? (tm ? colconvw_kernel<T, R1, R2, R3, W, LP, TT, true, CW_FWD, true> : colconvw_kernel<T, R1, R2, R3, W, LP, TT, true, CW_FWD>)
                  : colconvw_kernel<T, R1, R2, R3, W, LP, TT, false, CW_FWD>);
  } else {
    k = gv ? (tm ? colconvw_kernel<T, R1, R2, R3, W, LP, TT, true, CW_CONV, true> : colconvw_kernel<T, R1, R2, R3, W, LP, TT, true>)
           : colconvw_kernel<T, R1, R2, R3, W, LP, TT, false>;
  }
  static PerDeviceFlag flag[8];
  static int ctas_per_sm[8][kMaxDevices] = {};
  const int vi = (gv ? 1 : 0) + (bwd ? 2 : 0) + (tm ? 4 : 0);
  bool &configured = flag[vi].here();
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return (int)e;
    int nb = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k, TT, smem);
    if (e != cudaSuccess) return (int)e;
    ctas_per_sm[vi][cur_dev()] = nb > 0 ? nb : 1;
    configured = true;
  }
  // tile order: GF adjacent groups on CTAs that run side by side (IMPULSE_FFT_CONVW_GF overrides: A/B runs)
  static const int gf_env = [] { const char *e = getenv("IMPULSE_FFT_CONVW_GF"); return e ? atoi(e) : 0; }();
  // (measured on config 5, 32-byte runs: GF = 1: 9.61 ms, 2: 9.45, 4: 12.5, 8: 11.5, 16: 10.9 — side-by-side requests for
  // the pieces of one line contend instead of merging; default 1)
  const uint32_t gf = gf_env > 0 ? (uint32_t)gf_env : 1u;
  static const int pf_env = [] { const char *e = getenv("IMPULSE_FFT_CONVW_PF"); return e ? atoi(e) : 1; }();
  LineJob Jg = J;
  Jg.n_load = gf;
  Jg.n_store = (uint32_t)pf_env;
  const uint64_t g0n = (J.bdim[0] + W - 1) / W;
  const uint64_t tiles = (g0n + gf - 1) / gf * gf * J.bdim[1] * J.bdim[2];
  if (tiles == 0) return 0;
  if (tiles >= (1ull << 31)) return (int)cudaErrorInvalidValue;
  uint64_t grid = (uint64_t)sm_count * ctas_per_sm[vi][cur_dev()];
  if (grid > tiles) grid = tiles;
  static thread_local char name[96];
  snprintf(name, sizeof(name), "colconvw_kernel<%s,%d,%d,%d,%d,%d,%d>%s%s", sizeof(T) == 8 ? "double" : "float", R1, R2, R3, W, LP, TT,
           PLAIN ? (bwd ? "+bwd" : "+fwd") : "", gv ? (tm ? "+tmem" : "") : "+scalar");
  g_last_kernel = name;
  k<<<(unsigned)grid, TT, smem, s>>>(Jg);
  return (int)cudaGetLastError();
}
}  // namespace

int launch_colconvw_job(const LineJob &J, int sm_count, void *stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  // IMPULSE_FFT_CW_NARROW=1 (A/B runs): 1024-point axes as half-width tiles (64-byte runs, two CTAs per SM)
  static const int narrow = [] { const char *e = getenv("IMPULSE_FFT_CW_NARROW"); return e ? atoi(e) : 0; }();
  if (narrow) {
    switch (J.fast_id) {
      case COLCONVW_1024_F32: return launch_colconvw<float, 16, 8, 8, 8, 2, 256>(J, sm_count, s);
      case COLCONVW_1024_F64: return launch_colconvw<double, 16, 8, 8, 4, 2, 128>(J, sm_count, s);
      case COLW_1024_F32: return launch_colconvw<float, 16, 8, 8, 8, 2, 256, true>(J, sm_count, s);
      case COLW_1024_F64: return launch_colconvw<double, 16, 8, 8, 4, 2, 128, true>(J, sm_count, s);
      default: break;
    }
  }
  switch (J.fast_id) {
    // <T, R1, R2, R3, W lines per tile, LP lines per thread, threads>
    // 512 / 1024 points: 128-byte runs (16 float32 or 8 float64 lines); 2048 / 4096 points fit only 64- / 32-byte runs
    case COLCONVW_512_F32: return launch_colconvw<float, 8, 8, 8, 16, 2, 512>(J, sm_count, s);
    case COLCONVW_1024_F32: return launch_colconvw<float, 16, 8, 8, 16, 2, 512>(J, sm_count, s);
    case COLCONVW_2048_F32: return launch_colconvw<float, 16, 16, 8, 8, 2, 512>(J, sm_count, s);
    case COLCONVW_4096_F32: return launch_colconvw<float, 16, 16, 16, 4, 2, 512>(J, sm_count, s);
    case COLCONVW_512_F64: return launch_colconvw<double, 8, 8, 8, 8, 2, 256>(J, sm_count, s);
    case COLCONVW_1024_F64: return launch_colconvw<double, 16, 8, 8, 8, 2, 256>(J, sm_count, s);
    case COLCONVW_2048_F64: return launch_colconvw<double, 16, 16, 8, 4, 2, 256>(J, sm_count, s);
    case COLCONVW_4096_F64: return launch_colconvw<double, 16, 16, 16, 2, 2, 256>(J, sm_count, s);
    case COLW_1024_F32: return launch_colconvw<float, 16, 8, 8, 16, 2, 512, true>(J, sm_count, s);
    case COLW_2048_F32: return launch_colconvw<float, 16, 16, 8, 8, 2, 512, true>(J, sm_count, s);
    case COLW_1024_F64: return launch_colconvw<double, 16, 8, 8, 8, 2, 256, true>(J, sm_count, s);
    case COLW_2048_F64: return launch_colconvw<double, 16, 16, 8, 4, 2, 256, true>(J, sm_count, s);
    default: return (int)cudaErrorInvalidValue;
  }
}

}  // namespace impulse
