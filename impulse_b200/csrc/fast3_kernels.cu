// Three-pass register kernels (fast3_kernel, fast3_device.cuh): launcher and the table of instantiated shapes.
// A translation unit of its own: its ~150 kernel instances dominate the build time, and nvcc compiles the two
// register-kernel files in parallel.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>

#include "fast3_device.cuh"
#include "fast_common.h"
#include "fft_device.cuh"
#include "fft_kernels.h"

namespace impulse {

namespace {
// IMPULSE_FFT_R2C_PAIR=0 / IMPULSE_FFT_C2R_PAIR=0 restore the post- / pre-twiddle through shared memory (A/B runs)
inline bool r2c_pair_enabled() {
  static const int v = [] { const char *e = getenv("IMPULSE_FFT_R2C_PAIR"); return e ? atoi(e) : 1; }();
  return v != 0;
}
// Register prefetch of the next claimed row during pass 3 (PF variants).  Measured on the B200: +7 % for c2c rows of
// 2048 points (5.74 -> 6.14 TB/s), within +-2 % for the paired r2c shapes, slower wherever the extra live
// registers spill (c2c 4096: 5.24 -> 4.71).  So it is on by default only where the launcher says so (PFDEF);
// IMPULSE_FFT_F3_PF=0 / 1 forces it off / on for every instantiated variant (A/B runs).
inline bool f3_prefetch_enabled(bool dflt) {
  static const int v = [] { const char *e = getenv("IMPULSE_FFT_F3_PF"); return e ? atoi(e) : -1; }();
  return v < 0 ? dflt : v != 0;
}
// Second exchange buffer (two barriers per row instead of four) where instantiated.  Measured on the B200
// (profiles/r02_ab_round2.txt): r2c 4096 f64 4.67 -> 5.18 TB/s, c2r 4096 4.77 -> 4.94, the fp32 rows of the same
// shapes +2-4 %, 1000-point real rows +1 %; r2c 3888 LOSES 7 % (the 18-point shape drops a resident CTA) and c2c
// 2048 f64 2 %.  So it is the default only for the kinds a shape lists in DBDEF; IMPULSE_FFT_F3_DB=0 / 1 forces it
// off / on for every instantiated variant (A/B runs).
inline bool f3_double_buffer(bool dflt) {
  static const int v = [] { const char *e = getenv("IMPULSE_FFT_F3_DB"); return e ? atoi(e) : -1; }();
  return v < 0 ? dflt : v != 0;
}
inline bool c2r_pair_enabled() {
  static const int v = [] { const char *e = getenv("IMPULSE_FFT_C2R_PAIR"); return e ? atoi(e) : 1; }();
  return v != 0;
}

// KINDS: which kinds the shape is instantiated for (1 = c2c, 2 = r2c, 4 = c2r).
// PAIRS: 1 = r2c post-twiddle in pass 3 (pair units) unless switched off, 2 = r2c always paired (the shape exists
//        for that variant only); 4 / 8 = the same for the c2r pre-twiddle in pass 1.
// PFK:   kinds for which the register-prefetch variant is instantiated (same bits as KINDS; real kinds: pair only)
// PFDEF: kinds for which it is the default
// DBK:   kinds for which the second-exchange-buffer variant is instantiated (real kinds: pair only)
// DBDEF: kinds for which it is the default
template <typename T, int R1, int R2, int R3, int E, int MINB, int KINDS = 7, int PAIRS = 0, int PFK = 0, int PFDEF = 0, int DBK = 0, int DBDEF = 0>
int launch_fast3(const LineJob &J, int sm_count, cudaStream_t s) {
  constexpr int N = R1 * R2 * R3, TT = N / E, M1 = N / R1, S = sizeof(T) == 8 ? 8 : 16;
  constexpr int P1 = ((M1 + S - 1) / S) * S + 1;
  constexpr int BUFN = (R1 * P1 > N + 1) ? R1 * P1 : N + 1;
  size_t smem = sizeof(cx<T>) * ((size_t)BUFN + (size_t)R2 * R3) + 16;
  const int kind = J.store_mode == ST_R2C_EVEN ? F3_R2C : J.load_mode == LD_HERM_EVEN ? F3_C2R : F3_C2C;
  const bool bwd = kind == F3_C2C ? (J.flags & F_CONJ_SEQ) != 0 : kind == F3_R2C ? (J.flags & F_CONJ_RESULT) != 0 : (J.flags & F_CONJ_IN) != 0;
  typedef void (*kern_t)(const void *, void *, uint64_t, int64_t, int64_t, const cx<T> *, const cx<T> *, const cx<T> *, T, unsigned int *);
  kern_t k = nullptr;
  bool pair = false, pf = false;
  if (kind == F3_C2C) {
    if constexpr ((KINDS & 1) != 0) k = bwd ? (kern_t)fast3_kernel<T, R1, R2, R3, E, F3_C2C, true, MINB> : (kern_t)fast3_kernel<T, R1, R2, R3, E, F3_C2C, false, MINB>;
    if constexpr ((KINDS & 1) != 0 && (PFK & 1) != 0) {
      if (f3_prefetch_enabled((PFDEF & 1) != 0)) {
        pf = true;
        k = bwd ? (kern_t)fast3_kernel<T, R1, R2, R3, E, F3_C2C, true, MINB, false, true> : (kern_t)fast3_kernel<T, R1, R2, R3, E, F3_C2C, false, MINB, false, true>;
      }
    }
  } else if (kind == F3_R2C) {
    if constexpr ((KINDS & 2) != 0) {
      if constexpr ((PAIRS & 3) != 0) {
        if ((PAIRS & 2) != 0 || r2c_pair_enabled()) {
          pair = true;
          k = bwd ? (kern_t)fast3_kernel<T, R1, R2, R3, E, F3_R2C, true, MINB, true> : (kern_t)fast3_kernel<T, R1, R2, R3, E, F3_R2C, false, MINB, true>;
          if constexpr ((PFK & 2) != 0) {
            if (f3_prefetch_enabled((PFDEF & 2) != 0)) {
              pf = true;
              k = bwd ? (kern_t)fast3_kernel<T, R1, R2, R3, E, F3_R2C, true, MINB, true, true> : (kern_t)fast3_kernel<T, R1, R2, R3, E, F3_R2C, false, MINB, true, true>;
            }
          }
        }
      }
      if constexpr ((PAIRS & 2) == 0) {
        if (!pair) k = bwd ? (kern_t)fast3_kernel<T, R1, R2, R3, E, F3_R2C, true, MINB> : (kern_t)fast3_kernel<T, R1, R2, R3, E, F3_R2C, false, MINB>;
      }
    }
  } else {
    if constexpr ((KINDS & 4) != 0) {
      if constexpr ((PAIRS & 12) != 0) {
        if ((PAIRS & 8) != 0 || c2r_pair_enabled()) {
          pair = true;
          k = bwd ? (kern_t)fast3_kernel<T, R1, R2, R3, E, F3_C2R, true, MINB, true> : (kern_t)fast3_kernel<T, R1, R2, R3, E, F3_C2R, false, MINB, true>;
        }
      }
      if constexpr ((PAIRS & 8) == 0) {
        if (!pair) k = bwd ? (kern_t)fast3_kernel<T, R1, R2, R3, E, F3_C2R, true, MINB> : (kern_t)fast3_kernel<T, R1, R2, R3, E, F3_C2R, false, MINB>;
      }
    }
  }
  bool db = false;
  if constexpr (DBK != 0) {
    if (f3_double_buffer((DBDEF & (1 << kind)) != 0)) {
      if constexpr ((DBK & 1) != 0 && (KINDS & 1) != 0) {
        if (kind == F3_C2C) { db = true; k = bwd ? (kern_t)fast3_kernel<T, R1, R2, R3, E, F3_C2C, true, MINB, false, false, true> : (kern_t)fast3_kernel<T, R1, R2, R3, E, F3_C2C, false, MINB, false, false, true>; }
      }
      if constexpr ((DBK & 2) != 0 && (KINDS & 2) != 0 && (PAIRS & 3) != 0) {
        if (kind == F3_R2C && pair) { db = true; k = bwd ? (kern_t)fast3_kernel<T, R1, R2, R3, E, F3_R2C, true, MINB, true, false, true> : (kern_t)fast3_kernel<T, R1, R2, R3, E, F3_R2C, false, MINB, true, false, true>; }
      }
      if constexpr ((DBK & 4) != 0 && (KINDS & 4) != 0 && (PAIRS & 12) != 0) {
        if (kind == F3_C2R && pair) { db = true; k = bwd ? (kern_t)fast3_kernel<T, R1, R2, R3, E, F3_C2R, true, MINB, true, false, true> : (kern_t)fast3_kernel<T, R1, R2, R3, E, F3_C2R, false, MINB, true, false, true>; }
      }
      if (db) { pf = false; smem += sizeof(cx<T>) * (size_t)BUFN; }
    }
  }
  if (!k) return (int)cudaErrorInvalidValue;   // the planner never selects a shape for a kind it is not built for
  if (pair || pf || db) {  // reported by impulse_fft_last_kernel()
    static thread_local char name[96];
    snprintf(name, sizeof(name), "%s%s%s%s", g_last_kernel, pair ? "+pair" : "", pf ? "+pf" : "", db ? "+db" : "");
    g_last_kernel = name;
  }
  static PerDeviceFlag flags[30];
  bool &configured_here = flags[db ? 24 + kind * 2 + (bwd ? 1 : 0) : (pf ? 12 : 0) + (pair ? 6 : 0) + kind * 2 + (bwd ? 1 : 0)].here();
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
  cudaError_t le = launch_pdl(k, (unsigned)grid, (unsigned)TT, smem, s, (const void *)J.in, (void *)J.out, (uint64_t)J.n_lines, (int64_t)J.bs_in[0],
                              (int64_t)J.bs_out[0], (const cx<T> *)J.f3_tw1, (const cx<T> *)J.f3_tw2, (const cx<T> *)J.tw_r, (T)J.fct, sched);
  return le != cudaSuccess ? (int)le : (int)cudaGetLastError();
}
}  // namespace

// IMPULSE_FFT_F3_MINB4=1: 2048-point fp64 shapes at four CTAs per SM (128 registers) instead of three (A/B runs)
static int f3_minb4() {
  static int v = -1;
  if (v < 0) { const char *e = getenv("IMPULSE_FFT_F3_MINB4"); v = e ? atoi(e) : 0; }
  return v;
}

int launch_fast3_job(const LineJob &J, int sm_count, void *stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  {
    const int rc = launch_fast3_staged_job(J, sm_count, stream);   // rows staged by bulk copies where that variant exists
    if (rc != -1) return rc;
  }
  switch (J.fast_id) {
    case FAST3R_256_F64: g_last_kernel = "fast3_kernel<double,8,8,4,E8>"; return launch_fast3<double, 8, 8, 4, 8, 16, 6, 1>(J, sm_count, s);
    case FAST3R_512_F64: g_last_kernel = "fast3_kernel<double,8,8,8,E8>"; return launch_fast3<double, 8, 8, 8, 8, 8, 6, 0>(J, sm_count, s);
    case FAST3R_1024_F64: g_last_kernel = "fast3_kernel<double,16,8,8,E16>"; return launch_fast3<double, 16, 8, 8, 16, 8, 6, 1>(J, sm_count, s);
    case FAST3R_256_F32: g_last_kernel = "fast3_kernel<float,8,8,4,E8>"; return launch_fast3<float, 8, 8, 4, 8, 16, 6, 1>(J, sm_count, s);
    case FAST3R_512_F32: g_last_kernel = "fast3_kernel<float,8,8,8,E8>"; return launch_fast3<float, 8, 8, 8, 8, 12, 6, 0>(J, sm_count, s);
    case FAST3R_1024_F32: g_last_kernel = "fast3_kernel<float,16,8,8,E16>"; return launch_fast3<float, 16, 8, 8, 16, 12, 6, 1>(J, sm_count, s);
    case FAST3_2048_F64:
      if (f3_minb4()) { g_last_kernel = "fast3_kernel<double,16,16,8,E16,minb4>"; return launch_fast3<double, 16, 16, 8, 16, 4, 7, 1, 3, 1>(J, sm_count, s); }
      g_last_kernel = "fast3_kernel<double,16,16,8,E16>"; return launch_fast3<double, 16, 16, 8, 16, 3, 7, 1, 3, 1, 3, 2>(J, sm_count, s);
    case FAST3_4096_F64: g_last_kernel = "fast3_kernel<double,16,16,16,E16>"; return launch_fast3<double, 16, 16, 16, 16, 2, 7, 0>(J, sm_count, s);
    case FAST3_8192_F64: g_last_kernel = "fast3_kernel<double,16,16,32,E32>"; return launch_fast3<double, 16, 16, 32, 32, 1, 7, 0>(J, sm_count, s);
    case FAST3_500_F64: g_last_kernel = "fast3_kernel<double,5,10,10,E10>"; return launch_fast3<double, 5, 10, 10, 10, 8, 7, 4, 0, 0, 4, 4>(J, sm_count, s);
    case FAST3_1944_F64: g_last_kernel = "fast3_kernel<double,6,18,18,E18>"; return launch_fast3<double, 6, 18, 18, 18, 4, 7, 4, 0, 0, 4>(J, sm_count, s);
    case FAST3_1000_F64: g_last_kernel = "fast3_kernel<double,10,10,10,E10>"; return launch_fast3<double, 10, 10, 10, 10, 5, 7, 0>(J, sm_count, s);
    case FAST3R_500_F64: g_last_kernel = "fast3_kernel<double,10,10,5,E10>"; return launch_fast3<double, 10, 10, 5, 10, 8, 2, 2, 2, 0, 2, 2>(J, sm_count, s);
    case FAST3R_1944_F64: g_last_kernel = "fast3_kernel<double,18,18,6,E18>"; return launch_fast3<double, 18, 18, 6, 18, 4, 2, 2, 0, 0, 2>(J, sm_count, s);
    case FAST3C_2048_F64:
      if (f3_minb4()) { g_last_kernel = "fast3_kernel<double,8,16,16,E16,minb4>"; return launch_fast3<double, 8, 16, 16, 16, 4, 4, 8>(J, sm_count, s); }
      g_last_kernel = "fast3_kernel<double,8,16,16,E16>"; return launch_fast3<double, 8, 16, 16, 16, 3, 4, 8, 0, 0, 4, 4>(J, sm_count, s);
    case FAST3C_2048_F32: g_last_kernel = "fast3_kernel<float,8,16,16,E16>"; return launch_fast3<float, 8, 16, 16, 16, 4, 4, 8, 0, 0, 4, 4>(J, sm_count, s);
    case FAST3C_1024_F64: g_last_kernel = "fast3_kernel<double,8,8,16,E16>"; return launch_fast3<double, 8, 8, 16, 16, 8, 4, 8>(J, sm_count, s);
    case FAST3C_1024_F32: g_last_kernel = "fast3_kernel<float,8,8,16,E16>"; return launch_fast3<float, 8, 8, 16, 16, 12, 4, 8>(J, sm_count, s);
    case FAST3_500_F32: g_last_kernel = "fast3_kernel<float,5,10,10,E10>"; return launch_fast3<float, 5, 10, 10, 10, 16, 7, 4>(J, sm_count, s);
    case FAST3_1944_F32: g_last_kernel = "fast3_kernel<float,6,18,18,E18>"; return launch_fast3<float, 6, 18, 18, 18, 6, 7, 4>(J, sm_count, s);
    case FAST3_1000_F32: g_last_kernel = "fast3_kernel<float,10,10,10,E10>"; return launch_fast3<float, 10, 10, 10, 10, 8, 7, 0>(J, sm_count, s);
    case FAST3R_500_F32: g_last_kernel = "fast3_kernel<float,10,10,5,E10>"; return launch_fast3<float, 10, 10, 5, 10, 16, 2, 2>(J, sm_count, s);
    case FAST3R_1944_F32: g_last_kernel = "fast3_kernel<float,18,18,6,E18>"; return launch_fast3<float, 18, 18, 6, 18, 6, 2, 2>(J, sm_count, s);
    case FAST3P_512_F64: g_last_kernel = "fast3_kernel<double,8,8,8,E16>"; return launch_fast3<double, 8, 8, 8, 16, 16, 6, 10>(J, sm_count, s);
    case FAST3P_512_F32: g_last_kernel = "fast3_kernel<float,8,8,8,E16>"; return launch_fast3<float, 8, 8, 8, 16, 24, 6, 10>(J, sm_count, s);
    case FAST3_8192_F32: g_last_kernel = "fast3_kernel<float,16,16,32,E32>"; return launch_fast3<float, 16, 16, 32, 32, 2, 7, 0>(J, sm_count, s);
    case FAST3_2048_F32: g_last_kernel = "fast3_kernel<float,16,16,8,E16>"; return launch_fast3<float, 16, 16, 8, 16, 4, 7, 1, 3, 0, 3, 3>(J, sm_count, s);
    case FAST3_4096_F32: g_last_kernel = "fast3_kernel<float,16,16,16,E16>"; return launch_fast3<float, 16, 16, 16, 16, 3, 7, 0>(J, sm_count, s);
    case FAST3_1536_F64: g_last_kernel = "fast3_kernel<double,8,24,8,E24>"; return launch_fast3<double, 8, 24, 8, 24, 4, 7, 5>(J, sm_count, s);
    case FAST3_2000_F64: g_last_kernel = "fast3_kernel<double,10,20,10,E20>"; return launch_fast3<double, 10, 20, 10, 20, 3, 7, 5>(J, sm_count, s);
    case FAST3_4000_F64: g_last_kernel = "fast3_kernel<double,10,20,20,E20>"; return launch_fast3<double, 10, 20, 20, 20, 2, 7, 4>(J, sm_count, s);
    case FAST3_2187_F64: g_last_kernel = "fast3_kernel<double,27,9,9,E27>"; return launch_fast3<double, 27, 9, 9, 27, 2, 7, 0>(J, sm_count, s);
    case FAST3_3000_F64: g_last_kernel = "fast3_kernel<double,10,30,10,E30>"; return launch_fast3<double, 10, 30, 10, 30, 2, 7, 0>(J, sm_count, s);
    case FAST3_6561_F64: g_last_kernel = "fast3_kernel<double,27,27,9,E27>"; return launch_fast3<double, 27, 27, 9, 27, 1, 7, 0>(J, sm_count, s);
    case FAST3_1536_F32: g_last_kernel = "fast3_kernel<float,8,24,8,E24>"; return launch_fast3<float, 8, 24, 8, 24, 6, 7, 0>(J, sm_count, s);
    case FAST3_2000_F32: g_last_kernel = "fast3_kernel<float,10,20,10,E20>"; return launch_fast3<float, 10, 20, 10, 20, 5, 7, 0>(J, sm_count, s);
    case FAST3_4000_F32: g_last_kernel = "fast3_kernel<float,10,20,20,E20>"; return launch_fast3<float, 10, 20, 20, 20, 3, 7, 0>(J, sm_count, s);
    case FAST3_2187_F32: g_last_kernel = "fast3_kernel<float,27,9,9,E27>"; return launch_fast3<float, 27, 9, 9, 27, 4, 7, 0>(J, sm_count, s);
    case FAST3_3000_F32: g_last_kernel = "fast3_kernel<float,10,30,10,E30>"; return launch_fast3<float, 10, 30, 10, 30, 4, 7, 0>(J, sm_count, s);
    case FAST3_6561_F32: g_last_kernel = "fast3_kernel<float,27,27,9,E27>"; return launch_fast3<float, 27, 27, 9, 27, 2, 7, 0>(J, sm_count, s);
    default: return (int)cudaErrorInvalidValue;
  }
}

}  // namespace impulse
