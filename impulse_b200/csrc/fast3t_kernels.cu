// Staged variants of the three-pass register kernels (fast3_kernel<..., TMA = true>, fast3_device.cuh): the next
// claimed row streams into a shared staging buffer by cp.async.bulk while the current one is transformed.
// A translation unit of its own (compiled in parallel with fast3_kernels.cu); launch_fast3_job tries this table first.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>

#include "fast3_device.cuh"
#include "fast_common.h"
#include "fft_device.cuh"
#include "fft_kernels.h"

namespace impulse {

// IMPULSE_FFT_F3_TMA: 0 = off (the direct-load kernels of fast3_kernels.cu), 1 = staged rows with one exchange buffer,
// 2 = staged rows with the second exchange buffer, -1 / unset = the per-shape default below (A/B runs decide it).
int f3_tma_mode() {
  static const int v = [] { const char *e = getenv("IMPULSE_FFT_F3_TMA"); return e ? atoi(e) : -1; }();
  return v;
}

namespace {
// DBV: 0 = one exchange buffer, 1 = two.  MINB: resident CTAs per SM the shared memory allows for that variant.
template <typename T, int R1, int R2, int R3, int E, int KIND, bool DBV, int MINB>
int launch_staged(const LineJob &J, int sm_count, cudaStream_t s, bool bwd) {
  constexpr int N = R1 * R2 * R3, TT = N / E, M1 = N / R1, S = sizeof(T) == 8 ? 8 : 16;
  constexpr int P1 = ((M1 + S - 1) / S) * S + 1;
  constexpr int BUFN = (R1 * P1 > N + 1) ? R1 * P1 : N + 1;
  constexpr bool PAIR = KIND != F3_C2C;
  constexpr size_t smem = sizeof(cx<T>) * ((size_t)(DBV ? 2 : 1) * BUFN + (size_t)R2 * R3) + 16 + 128 + F3Stage<T, N, KIND>::BYTES + 16;
  // the bulk copy wants 16-byte aligned rows, and the rounded-up copy must stay inside the row stride
  const size_t esz = KIND == F3_R2C ? sizeof(T) : sizeof(cx<T>);
  const size_t stride_b = (size_t)J.bs_in[0] * esz;
  if (((uintptr_t)J.in & 15) || (stride_b & 15) || (J.n_lines > 1 && stride_b < F3Stage<T, N, KIND>::BYTES) || J.bs_in[0] <= 0) return -1;
  if (F3Stage<T, N, KIND>::BYTES != F3Stage<T, N, KIND>::ROW_BYTES && stride_b < F3Stage<T, N, KIND>::BYTES) return -1;
  typedef void (*kern_t)(const void *, void *, uint64_t, int64_t, int64_t, const cx<T> *, const cx<T> *, const cx<T> *, T, unsigned int *);
  kern_t k = bwd ? (kern_t)fast3_kernel<T, R1, R2, R3, E, KIND, true, MINB, PAIR, false, DBV, true>
                 : (kern_t)fast3_kernel<T, R1, R2, R3, E, KIND, false, MINB, PAIR, false, DBV, true>;
  static PerDeviceFlag flags[2];
  bool &configured_here = flags[bwd ? 1 : 0].here();
  if (!configured_here) {
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return (int)e;
    configured_here = true;
  }
  const uint64_t grid = f3_grid(J.n_lines, (uint64_t)sm_count * MINB);
  unsigned int *sched = sched_slot();
  if (!sched) return (int)cudaErrorMemoryAllocation;
  if (J.n_lines > 0xfff00000ull) return (int)cudaErrorInvalidValue;
  {
    static thread_local char name[112];
    snprintf(name, sizeof(name), "fast3_kernel<%s,%d,%d,%d,E%d>%s+tma%s", sizeof(T) == 8 ? "double" : "float", R1, R2, R3, E,
             PAIR ? "+pair" : "", DBV ? "+db" : "");
    g_last_kernel = name;
  }
  cudaError_t le = launch_pdl(k, (unsigned)grid, (unsigned)TT, smem, s, (const void *)J.in, (void *)J.out, (uint64_t)J.n_lines, (int64_t)J.bs_in[0],
                              (int64_t)J.bs_out[0], (const cx<T> *)J.f3_tw1, (const cx<T> *)J.f3_tw2, (const cx<T> *)J.tw_r, (T)J.fct, sched);
  return le != cudaSuccess ? (int)le : (int)cudaGetLastError();
}
}  // namespace

// returns -1 when there is no staged variant for this job (shape, kind, alignment, or switched off): the caller
// launches the direct-load kernel instead
int launch_fast3_staged_job(const LineJob &J, int sm_count, void *stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int mode = f3_tma_mode();
  if (mode == 0) return -1;
  const int kind = J.store_mode == ST_R2C_EVEN ? F3_R2C : J.load_mode == LD_HERM_EVEN ? F3_C2R : F3_C2C;
  const bool bwd = kind == F3_C2C ? (J.flags & F_CONJ_SEQ) != 0 : kind == F3_R2C ? (J.flags & F_CONJ_RESULT) != 0 : (J.flags & F_CONJ_IN) != 0;
  // per-shape default when mode < 0: DEF = 0 off, 1 = staged, 2 = staged + second exchange buffer — the better of the
  // two on the B200 (profiles/r02_ab_tma.txt: r2c 3888 4.12 -> 4.71 TB/s, c2r 3888 4.22 -> 4.58, c2r 4096 4.92 -> 5.48,
  // r2c 4096 5.19 -> 5.37, real 1000 +3-7 %, c2c 2048 6.17 -> 6.56, r2c 4096 fp32 4.70 -> 5.24)
#define STAGED(DEF, T, R1, R2, R3, E, KIND, MINB1, MINB2)                                                   \
  {                                                                                                          \
    const int m = mode < 0 ? DEF : mode;                                                                     \
    if (m == 1) return launch_staged<T, R1, R2, R3, E, KIND, false, MINB1>(J, sm_count, s, bwd);             \
    if (m >= 2) return launch_staged<T, R1, R2, R3, E, KIND, true, MINB2>(J, sm_count, s, bwd);              \
    return -1;                                                                                               \
  }
#define STAGED1(DEF, T, R1, R2, R3, E, KIND, MINB1)                                                         \
  {                                                                                                          \
    const int m = mode < 0 ? DEF : mode;                                                                     \
    if (m >= 1) return launch_staged<T, R1, R2, R3, E, KIND, false, MINB1>(J, sm_count, s, bwd);             \
    return -1;                                                                                               \
  }
  switch (J.fast_id) {
    case FAST3_2048_F64:   // real rows of 4096 points (config 1) and complex rows of 2048
      if (kind == F3_R2C) STAGED(2, double, 16, 16, 8, 16, F3_R2C, 3, 2)
      if (kind == F3_C2C) STAGED1(1, double, 16, 16, 8, 16, F3_C2C, 3)
      return -1;
    case FAST3C_2048_F64: if (kind == F3_C2R) STAGED(1, double, 8, 16, 16, 16, F3_C2R, 3, 2) return -1;
    case FAST3R_500_F64: if (kind == F3_R2C) STAGED(2, double, 10, 10, 5, 10, F3_R2C, 8, 8) return -1;      // config 3a
    case FAST3_500_F64: if (kind == F3_C2R) STAGED(2, double, 5, 10, 10, 10, F3_C2R, 8, 8) return -1;
    case FAST3R_1944_F64: if (kind == F3_R2C) STAGED(1, double, 18, 18, 6, 18, F3_R2C, 3, 2) return -1;     // config 3b
    case FAST3_1944_F64: if (kind == F3_C2R) STAGED(1, double, 6, 18, 18, 18, F3_C2R, 3, 2) return -1;
    case FAST3_2048_F32: if (kind == F3_R2C) STAGED(2, float, 16, 16, 8, 16, F3_R2C, 4, 4) return -1;       // config 5 rows
    case FAST3C_2048_F32: if (kind == F3_C2R) STAGED(2, float, 8, 16, 16, 16, F3_C2R, 4, 4) return -1;
    default: return -1;
  }
#undef STAGED
#undef STAGED1
}

}  // namespace impulse
