// Fused Bluestein kernels (fastblue_kernel, fastblue_device.cuh): launcher and the table of instantiated work lengths.
// A translation unit of its own, like fast3_kernels.cu: its instances were the longest part of fast_kernels.cu's
// compile, and the three register-kernel files now build in parallel.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>

#include "fast3_device.cuh"
#include "fast_common.h"
#include "fastblue_device.cuh"
#include "fft_device.cuh"
#include "fft_kernels.h"

namespace impulse {

namespace {
// Chirp table in shared memory (8192-point work length only: one CTA per SM either way).  Measured on config 3c
// (16384 x 4099 fp64): r2c 1.717 -> 1.584 ms, c2r 1.749 -> 1.481 ms.  IMPULSE_FFT_BLUE_BK_SMEM=0 switches it off.
inline bool blue_bk_smem() {
  static const int v = [] { const char *e = getenv("IMPULSE_FFT_BLUE_BK_SMEM"); return e ? atoi(e) : 1; }();
  return v != 0;
}
// Multiply by FFT(b)/M inside the first transform's last pass, half of the multipliers requested ahead of the
// butterfly.  Measured on config 3c: r2c 1.584 -> 1.276 ms, c2r 1.481 -> 1.386 ms.  IMPULSE_FFT_BLUE_BF_EARLY=0: off.
inline bool blue_bf_early() {
  static const int v = [] { const char *e = getenv("IMPULSE_FFT_BLUE_BF_EARLY"); return e ? atoi(e) : 1; }();
  return v != 0;
}
// The 8192-point work array on the four-pass core (512 threads x 16 points, 16 warps per SM instead of 8).  Measured
// on config 3c (profiles/r02_ab_round2.txt): r2c 1.277 -> 1.142 ms, c2r 1.386 -> 1.135 ms.  IMPULSE_FFT_BLUE_FOUR=0
// restores the three-pass 256-thread core.
inline bool blue_four_pass() {
  static const int v = [] { const char *e = getenv("IMPULSE_FFT_BLUE_FOUR"); return e ? atoi(e) : 1; }();
  return v != 0;
}
// FFT(b)/M in tensor memory (four-pass core with the early multiply): every thread keeps its 16 multipliers in its own
// columns for the life of the CTA instead of streaming 131 KB per unit from L2.  Measured on config 3c
// (profiles/r02_ab_blue_tmem.txt): r2c 1.138 -> 1.074 ms, c2r 1.130 -> 1.079 ms.  IMPULSE_FFT_BLUE_TMEM=0 switches it off.
inline bool blue_tmem() {
  static const int v = [] { const char *e = getenv("IMPULSE_FFT_BLUE_TMEM"); return e ? atoi(e) : 1; }();
  return v != 0;
}
template <typename T, int R1, int R2, int R3, int E>
int launch_fastblue(const LineJob &J, int sm_count, cudaStream_t s) {
  using F = Fft3<T, R1, R2, R3, E>;
  const bool bks = (R1 * R2 * R3 == 8192) && blue_bk_smem();
  const bool bfe = bks && blue_bf_early();
  const size_t smem = sizeof(cx<T>) * ((size_t)F::BUFN + (size_t)R2 * R3 + kBlueMaxDef + (bks ? (size_t)J.n_seq : 0)) + 32;   // + scheduler words, tensor-memory slot
  const int kind = J.store_mode == ST_HERM_HALF ? BL_R2C_PAIR : J.load_mode == LD_HERM_FULL ? BL_C2R_PAIR : BL_C2C;
  const bool bwd = kind == BL_C2C ? (J.flags & F_CONJ_SEQ) != 0 : kind == BL_R2C_PAIR ? (J.flags & F_CONJ_RESULT) != 0 : (J.flags & F_CONJ_IN) != 0;
  typedef void (*kern_t)(const void *, void *, uint64_t, int64_t, int64_t, uint32_t, uint32_t, const cx<T> *, const cx<T> *,
                         const cx<T> *, const cx<T> *, const cx<T> *, T, unsigned int *);
  kern_t k = nullptr;
  bool four = false, tmb = false;
  switch (kind * 2 + (bwd ? 1 : 0)) {
    case 0: k = fastblue_kernel<T, R1, R2, R3, E, BL_C2C, false>; break;
    case 1: k = fastblue_kernel<T, R1, R2, R3, E, BL_C2C, true>; break;
    case 2: k = fastblue_kernel<T, R1, R2, R3, E, BL_R2C_PAIR, false>; break;
    case 3: k = fastblue_kernel<T, R1, R2, R3, E, BL_R2C_PAIR, true>; break;
    case 4: k = fastblue_kernel<T, R1, R2, R3, E, BL_C2R_PAIR, false>; break;
    default: k = fastblue_kernel<T, R1, R2, R3, E, BL_C2R_PAIR, true>; break;
  }
  if constexpr (R1 * R2 * R3 == 8192) {
    if (bks) {
      switch (kind * 2 + (bwd ? 1 : 0)) {
        case 0: k = fastblue_kernel<T, R1, R2, R3, E, BL_C2C, false, true>; break;
        case 1: k = fastblue_kernel<T, R1, R2, R3, E, BL_C2C, true, true>; break;
        case 2: k = fastblue_kernel<T, R1, R2, R3, E, BL_R2C_PAIR, false, true>; break;
        case 3: k = fastblue_kernel<T, R1, R2, R3, E, BL_R2C_PAIR, true, true>; break;
        case 4: k = fastblue_kernel<T, R1, R2, R3, E, BL_C2R_PAIR, false, true>; break;
        default: k = fastblue_kernel<T, R1, R2, R3, E, BL_C2R_PAIR, true, true>; break;
      }
      g_last_kernel = sizeof(T) == 8 ? "fastblue_kernel<double,16,16,32,E32>+bk_smem" : "fastblue_kernel<float,16,16,32,E32>+bk_smem";
      if (bfe) {
        switch (kind * 2 + (bwd ? 1 : 0)) {
          case 0: k = fastblue_kernel<T, R1, R2, R3, E, BL_C2C, false, true, true>; break;
          case 1: k = fastblue_kernel<T, R1, R2, R3, E, BL_C2C, true, true, true>; break;
          case 2: k = fastblue_kernel<T, R1, R2, R3, E, BL_R2C_PAIR, false, true, true>; break;
          case 3: k = fastblue_kernel<T, R1, R2, R3, E, BL_R2C_PAIR, true, true, true>; break;
          case 4: k = fastblue_kernel<T, R1, R2, R3, E, BL_C2R_PAIR, false, true, true>; break;
          default: k = fastblue_kernel<T, R1, R2, R3, E, BL_C2R_PAIR, true, true, true>; break;
        }
        g_last_kernel = sizeof(T) == 8 ? "fastblue_kernel<double,16,16,32,E32>+bk_smem+bf_early" : "fastblue_kernel<float,16,16,32,E32>+bk_smem+bf_early";
      }
      if constexpr (sizeof(T) == 8) {
      if (blue_four_pass()) {
        four = true;
        if (bfe) {
          switch (kind * 2 + (bwd ? 1 : 0)) {
            case 0: k = fastblue_kernel<T, 16, 16, 32, 16, BL_C2C, false, true, true, true>; break;
            case 1: k = fastblue_kernel<T, 16, 16, 32, 16, BL_C2C, true, true, true, true>; break;
            case 2: k = fastblue_kernel<T, 16, 16, 32, 16, BL_R2C_PAIR, false, true, true, true>; break;
            case 3: k = fastblue_kernel<T, 16, 16, 32, 16, BL_R2C_PAIR, true, true, true, true>; break;
            case 4: k = fastblue_kernel<T, 16, 16, 32, 16, BL_C2R_PAIR, false, true, true, true>; break;
            default: k = fastblue_kernel<T, 16, 16, 32, 16, BL_C2R_PAIR, true, true, true, true>; break;
          }
          g_last_kernel = "fastblue_kernel<double,16,16,16,2,E16>+bk_smem+bf_early";
          if (blue_tmem()) {
            tmb = true;
            switch (kind * 2 + (bwd ? 1 : 0)) {
              case 0: k = fastblue_kernel<T, 16, 16, 32, 16, BL_C2C, false, true, true, true, true>; break;
              case 1: k = fastblue_kernel<T, 16, 16, 32, 16, BL_C2C, true, true, true, true, true>; break;
              case 2: k = fastblue_kernel<T, 16, 16, 32, 16, BL_R2C_PAIR, false, true, true, true, true>; break;
              case 3: k = fastblue_kernel<T, 16, 16, 32, 16, BL_R2C_PAIR, true, true, true, true, true>; break;
              case 4: k = fastblue_kernel<T, 16, 16, 32, 16, BL_C2R_PAIR, false, true, true, true, true>; break;
              default: k = fastblue_kernel<T, 16, 16, 32, 16, BL_C2R_PAIR, true, true, true, true, true>; break;
            }
            g_last_kernel = "fastblue_kernel<double,16,16,16,2,E16>+bk_smem+bf_early+bf_tmem";
          }
        } else {
          switch (kind * 2 + (bwd ? 1 : 0)) {
            case 0: k = fastblue_kernel<T, 16, 16, 32, 16, BL_C2C, false, true, false, true>; break;
            case 1: k = fastblue_kernel<T, 16, 16, 32, 16, BL_C2C, true, true, false, true>; break;
            case 2: k = fastblue_kernel<T, 16, 16, 32, 16, BL_R2C_PAIR, false, true, false, true>; break;
            case 3: k = fastblue_kernel<T, 16, 16, 32, 16, BL_R2C_PAIR, true, true, false, true>; break;
            case 4: k = fastblue_kernel<T, 16, 16, 32, 16, BL_C2R_PAIR, false, true, false, true>; break;
            default: k = fastblue_kernel<T, 16, 16, 32, 16, BL_C2R_PAIR, true, true, false, true>; break;
          }
          g_last_kernel = "fastblue_kernel<double,16,16,16,2,E16>+bk_smem";
        }
      }
      }
    }
  }
  const int threads = four ? 512 : F::TT;
  // the dynamic shared-memory size depends on L when the chirp table is resident: always raise the limit to the maximum
  const size_t smem_max = bks ? (size_t)227 * 1024 : smem;
  if (smem > smem_max) return (int)cudaErrorInvalidValue;
  static PerDeviceFlag flags[36];
  bool &configured_here = flags[(tmb ? 30 : four ? 18 + (bfe ? 6 : 0) : bfe ? 12 : bks ? 6 : 0) + kind * 2 + (bwd ? 1 : 0)].here();
  if (!configured_here) {
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return (int)e;
    configured_here = true;
  }
  const uint64_t units = kind == BL_C2C ? J.n_lines : (J.n_lines + 1) / 2;
  int per_sm = 1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, threads, smem);
  if (per_sm < 1) per_sm = 1;
  uint64_t grid = units;
  const uint64_t cap = (uint64_t)sm_count * per_sm;
  if (grid > cap) grid = cap;
  unsigned int *sched = sched_slot();
  if (!sched) return (int)cudaErrorMemoryAllocation;
  if (J.n_lines > 0xfff00000ull) return (int)cudaErrorInvalidValue;
  k<<<(unsigned)grid, threads, smem, s>>>(J.in, J.out, J.n_lines, J.bs_in[0], J.bs_out[0], J.n_seq, J.fb_d, (const cx<T> *)J.f3_tw1,
                                         (const cx<T> *)J.f3_tw2, (const cx<T> *)J.bk, (const cx<T> *)J.fb_bf,
                                         (const cx<T> *)J.fb_corr, (T)J.fct, sched);
  return (int)cudaGetLastError();
}
}  // namespace

int launch_fastblue_job(const LineJob &J, int sm_count, void *stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  switch (J.fast_id) {
    case FASTBLUE_2048_F64: g_last_kernel = "fastblue_kernel<double,16,16,8,E16>"; return launch_fastblue<double, 16, 16, 8, 16>(J, sm_count, s);
    case FASTBLUE_4096_F64: g_last_kernel = "fastblue_kernel<double,16,16,16,E16>"; return launch_fastblue<double, 16, 16, 16, 16>(J, sm_count, s);
    case FASTBLUE_8192_F64: g_last_kernel = "fastblue_kernel<double,16,16,32,E32>"; return launch_fastblue<double, 16, 16, 32, 32>(J, sm_count, s);
    case FASTBLUE_2048_F32: g_last_kernel = "fastblue_kernel<float,16,16,8,E16>"; return launch_fastblue<float, 16, 16, 8, 16>(J, sm_count, s);
    case FASTBLUE_4096_F32: g_last_kernel = "fastblue_kernel<float,16,16,16,E16>"; return launch_fastblue<float, 16, 16, 16, 16>(J, sm_count, s);
    case FASTBLUE_8192_F32: g_last_kernel = "fastblue_kernel<float,16,16,32,E32>"; return launch_fastblue<float, 16, 16, 32, 32>(J, sm_count, s);
    default: return (int)cudaErrorInvalidValue;
  }
}

}  // namespace impulse
