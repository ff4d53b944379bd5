// TEST INFRASTRUCTURE — thread-level host emulation of the column register kernels (colfast2 / colpipe2 / colconv2).
//
// Compiles impulse_b200/csrc/col_device.cuh (the source nvcc compiles) with g++: one OS thread per CUDA thread,
// __syncthreads() = a pthread barrier, CTAs of the 3-D grid one after another.  emu.cpp hands it the LineJobs the
// product's planner built (strided axes of N-D transforms, both launches of the four-step split, the fused
// convolution middle pass), so the jobs, tables and kernels are checked together.  The kernel/shape selection below
// mirrors launch_fast_job / launch_colfast2 (fast_kernels.cu).  NOT a product code path.
#include <pthread.h>

#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)
#define __shared__
#define __grid_constant__
#define __align__(n) __attribute__((aligned(n)))
struct emu_dim3 { unsigned x = 1, y = 1, z = 1; };
static thread_local emu_dim3 threadIdx;
static emu_dim3 blockIdx, gridDim;
static pthread_barrier_t g_bar;
static inline void __syncthreads() { pthread_barrier_wait(&g_bar); }
static inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }
static inline unsigned atomicAdd(unsigned *p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
template <typename V> static inline V __ldg(const V *p) { return *p; }
using std::min;

namespace impulse { alignas(128) unsigned char smem_raw[256 * 1024]; }

#include "../../impulse_b200/csrc/col_device.cuh"
#include "../../impulse_b200/csrc/colconvw_device.cuh"

using namespace impulse;

namespace {
template <typename F>
void launch(emu_dim3 grid, int threads, F &&body) {
  gridDim = grid;
  pthread_barrier_init(&g_bar, nullptr, threads);
  for (unsigned z = 0; z < grid.z; ++z)
    for (unsigned y = 0; y < grid.y; ++y)
      for (unsigned x = 0; x < grid.x; ++x) {
        blockIdx.x = x; blockIdx.y = y; blockIdx.z = z;
        std::memset(smem_raw, 0xCD, sizeof(smem_raw));
        std::vector<std::thread> th;
        for (int t = 0; t < threads; ++t) th.emplace_back([&, t] { threadIdx.x = (unsigned)t; body(); });
        for (auto &q : th) q.join();
      }
  pthread_barrier_destroy(&g_bar);
}

template <typename T, int R1, int R2, int LPC>
int run_colfast2(const LineJob &J, uint32_t pipe_groups) {
  const bool rows_in = J.col_in_rows != 0, bwd = (J.flags & F_CONJ_SEQ) != 0;
  const uint64_t g0n = (J.bdim[0] + LPC - 1) / LPC;
  if constexpr (sizeof(T) == 8 && R1 * R2 <= 128) {   // software-pipelined variant: same condition as launch_colfast2
    if (!rows_in && !J.seg_len && !J.umul_mod && pipe_groups > 1 && J.bdim[0] >= 2 * LPC) {
      const uint32_t G = pipe_groups;
      emu_dim3 grid; grid.x = (unsigned)((g0n + G - 1) / G); grid.y = (unsigned)J.bdim[1]; grid.z = (unsigned)J.bdim[2];
      if (bwd) launch(grid, LPC * R2, [&] { colpipe2_kernel<T, R1, R2, LPC, true>(J, G); });
      else launch(grid, LPC * R2, [&] { colpipe2_kernel<T, R1, R2, LPC, false>(J, G); });
      return 0;
    }
  }
  emu_dim3 grid; grid.x = (unsigned)g0n; grid.y = (unsigned)J.bdim[1]; grid.z = (unsigned)J.bdim[2];
  if (rows_in) {
    if (bwd) launch(grid, LPC * R2, [&] { colfast2_kernel<T, R1, R2, LPC, true, false, true>(J); });
    else launch(grid, LPC * R2, [&] { colfast2_kernel<T, R1, R2, LPC, false, false, true>(J); });
  } else {
    if (bwd) launch(grid, LPC * R2, [&] { colfast2_kernel<T, R1, R2, LPC, true, false, false>(J); });
    else launch(grid, LPC * R2, [&] { colfast2_kernel<T, R1, R2, LPC, false, false, false>(J); });
  }
  return 0;
}

template <typename T, int R1, int R2, int LPC>
int run_colconv2(const LineJob &J) {
  if (!J.umul || !J.umul_mod || !J.tw4_n) return -2;
  emu_dim3 grid; grid.x = (unsigned)((J.bdim[0] + LPC - 1) / LPC); grid.y = (unsigned)J.bdim[1]; grid.z = (unsigned)J.bdim[2];
  launch(grid, LPC * R2, [&] { colconv2_kernel<T, R1, R2, LPC, false>(J); });
  return 0;
}

// whole-axis convolution: persistent CTAs striding over the tiles (three CTAs here, so that the tile loop and the
// merged "I1 of this tile + F1 of the next" step are exercised)
template <typename T, int R1, int R2, int R3, int W, int LP, int TT, bool PLAIN = false>
int run_colconvw(const LineJob &J) {
  if (!J.f3_tw1 || !J.f3_tw2 || (!PLAIN && (!J.umul || !J.umul_mod))) return -2;
  const bool bwd = PLAIN && (J.flags & F_CONJ_SEQ) != 0;
  const uint32_t gf = 2;   // the product's default is 1 (IMPULSE_FFT_CONVW_GF); 2 exercises the block decode as well
  LineJob Jg = J;
  Jg.n_load = gf;
  Jg.n_store = 1;
  const uint64_t tiles = ((J.bdim[0] + W - 1) / W + gf - 1) / gf * gf * J.bdim[1] * J.bdim[2];
  emu_dim3 grid; grid.x = (unsigned)std::min<uint64_t>(tiles, 3);
  // the vector / scalar choice of launch_colconvw (colconvw_kernels.cu)
  constexpr uint64_t VB = LP * sizeof(cx<T>) >= 16 ? 16 : 8, VE = VB / sizeof(cx<T>) ? VB / sizeof(cx<T>) : 1;
  auto mult = [&](int64_t v) { return v % (int64_t)VE == 0; };
  const bool gv = (uintptr_t)J.in % VB == 0 && (uintptr_t)J.out % VB == 0 && (PLAIN || (uintptr_t)J.umul % VB == 0) && mult(J.es_in) &&
                  mult(J.es_out) && mult(J.bs_in[1]) && mult(J.bs_in[2]) && mult(J.bs_out[1]) && mult(J.bs_out[2]) &&
                  (PLAIN || J.umul_mod % VE == 0);
  const int tm_env = getenv("IMPULSE_FFT_CONVW_TMEM") ? atoi(getenv("IMPULSE_FFT_CONVW_TMEM")) : 1;   // as launch_colconvw
  const bool tm = gv && (tm_env >= 2 || (tm_env == 1 && sizeof(T) == 8));
  if (tm) {   // the tensor-memory staging logic, with a thread-private array standing in for the lanes' columns
    if constexpr (PLAIN) {
      if (bwd) launch(grid, TT, [&] { colconvw_kernel<T, R1, R2, R3, W, LP, TT, true, CW_BWD, true>(Jg); });
      else launch(grid, TT, [&] { colconvw_kernel<T, R1, R2, R3, W, LP, TT, true, CW_FWD, true>(Jg); });
    } else {
      launch(grid, TT, [&] { colconvw_kernel<T, R1, R2, R3, W, LP, TT, true, CW_CONV, true>(Jg); });
    }
    return 0;
  }
  if constexpr (PLAIN) {
    if (bwd && gv) launch(grid, TT, [&] { colconvw_kernel<T, R1, R2, R3, W, LP, TT, true, CW_BWD>(Jg); });
    else if (bwd) launch(grid, TT, [&] { colconvw_kernel<T, R1, R2, R3, W, LP, TT, false, CW_BWD>(Jg); });
    else if (gv) launch(grid, TT, [&] { colconvw_kernel<T, R1, R2, R3, W, LP, TT, true, CW_FWD>(Jg); });
    else launch(grid, TT, [&] { colconvw_kernel<T, R1, R2, R3, W, LP, TT, false, CW_FWD>(Jg); });
    return 0;
  }
  if (gv) launch(grid, TT, [&] { colconvw_kernel<T, R1, R2, R3, W, LP, TT, true>(Jg); });
  else launch(grid, TT, [&] { colconvw_kernel<T, R1, R2, R3, W, LP, TT, false>(Jg); });
  return 0;
}
}  // namespace

// returns 0 when the job ran on an emulated column kernel, 1 when its fast_id is not a column kernel, < 0 on error
int emu_run_col_job(const LineJob &J, unsigned pipe_groups) {
  switch (J.fast_id) {
    case COL2_32_F64: return run_colfast2<double, 8, 4, 8>(J, pipe_groups);
    case COL2_512_F64: return run_colfast2<double, 32, 16, 8>(J, pipe_groups);
    case COL2_32_F32: return run_colfast2<float, 8, 4, 16>(J, pipe_groups);
    case COL2_512_F32: return run_colfast2<float, 32, 16, 16>(J, pipe_groups);
    case COL2_64_F64: return J.bdim[0] >= 64 ? run_colfast2<double, 8, 8, 16>(J, pipe_groups) : run_colfast2<double, 8, 8, 8>(J, pipe_groups);
    case COL2_128_F64: return J.bdim[0] >= 64 ? run_colfast2<double, 16, 8, 16>(J, pipe_groups) : run_colfast2<double, 16, 8, 8>(J, pipe_groups);
    case COL2_256_F64: return run_colfast2<double, 16, 16, 8>(J, pipe_groups);
    case COL2_64_F32: return run_colfast2<float, 8, 8, 16>(J, pipe_groups);
    case COL2_128_F32: return run_colfast2<float, 16, 8, 16>(J, pipe_groups);
    case COL2_256_F32: return run_colfast2<float, 16, 16, 16>(J, pipe_groups);
    case COLCONV_32_F64: return run_colconv2<double, 8, 4, 8>(J);
    case COLCONV_64_F64: return run_colconv2<double, 8, 8, 8>(J);
    case COLCONV_128_F64: return run_colconv2<double, 16, 8, 8>(J);
    case COLCONV_32_F32: return run_colconv2<float, 8, 4, 16>(J);
    case COLCONV_64_F32: return run_colconv2<float, 8, 8, 16>(J);
    case COLCONV_128_F32: return run_colconv2<float, 16, 8, 16>(J);
    case COLCONVW_512_F32: return run_colconvw<float, 8, 8, 8, 16, 2, 512>(J);
    case COLCONVW_1024_F32: return run_colconvw<float, 16, 8, 8, 16, 2, 512>(J);
    case COLCONVW_2048_F32: return run_colconvw<float, 16, 16, 8, 8, 2, 512>(J);
    case COLCONVW_4096_F32: return run_colconvw<float, 16, 16, 16, 4, 2, 512>(J);
    case COLCONVW_512_F64: return run_colconvw<double, 8, 8, 8, 8, 2, 256>(J);
    case COLCONVW_1024_F64: return run_colconvw<double, 16, 8, 8, 8, 2, 256>(J);
    case COLCONVW_2048_F64: return run_colconvw<double, 16, 16, 8, 4, 2, 256>(J);
    case COLCONVW_4096_F64: return run_colconvw<double, 16, 16, 16, 2, 2, 256>(J);
    case COLW_1024_F32: return run_colconvw<float, 16, 8, 8, 16, 2, 512, true>(J);
    case COLW_2048_F32: return run_colconvw<float, 16, 16, 8, 8, 2, 512, true>(J);
    case COLW_1024_F64: return run_colconvw<double, 16, 8, 8, 8, 2, 256, true>(J);
    case COLW_2048_F64: return run_colconvw<double, 16, 16, 8, 4, 2, 256, true>(J);
    default: return 1;
  }
}
