// TEST INFRASTRUCTURE — host emulation of the three-pass register kernel (fast3_kernel).
//
// Compiles impulse_b200/csrc/fast3_device.cuh with g++: one OS thread per CUDA thread of a CTA,
// __syncthreads() is a pthread barrier, shared memory is a static buffer, CTAs run one after another.
// It checks the kernel's index algebra (register ownership, shared-memory layouts, the Hermitian
// post-/pre-twiddle, the in-register pair units, the dynamic row claims) in the GPU-less build
// container before GPU time is spent.  NOT a product code path: the product library has no CPU transform.
#include <pthread.h>

#include <atomic>
#include <cmath>
#include <complex>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

// ---- CUDA shims ------------------------------------------------------------------------------
#define __global__
#define __device__
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)
#define __shared__
#define __align__(n) __attribute__((aligned(n)))
struct emu_dim3 { unsigned x = 1, y = 1, z = 1; };
static thread_local emu_dim3 threadIdx;
static emu_dim3 blockIdx, gridDim;
static pthread_barrier_t g_bar;
static inline void __syncthreads() { pthread_barrier_wait(&g_bar); }
static inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }
static inline unsigned atomicAdd(unsigned *p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
template <typename V> static inline V __ldg(const V *p) { return *p; }

namespace impulse { alignas(128) unsigned char smem_raw[256 * 1024]; }

#include "../../impulse_b200/csrc/fast3_device.cuh"

using namespace impulse;

namespace {
typedef std::complex<long double> cld;
cld root(uint64_t num, uint64_t den) {  // exp(-2 pi i num/den)
  const long double a = -2.0L * 3.141592653589793238462643383279502884L * (long double)(num % den) / (long double)den;
  return cld(cosl(a), sinl(a));
}
template <typename T> std::vector<cx<T>> conv(const std::vector<cld> &v) {
  std::vector<cx<T>> r(v.size());
  for (size_t i = 0; i < v.size(); ++i) r[i] = mk<T>((T)v[i].real(), (T)v[i].imag());
  return r;
}

template <typename T, int R1, int R2, int R3, int E, int KIND, bool BWD, bool PAIR, bool PF = false, bool DB = false, bool TMA = false>
void run(const void *in, void *out, uint64_t nrows, int64_t rs_in, int64_t rs_out, double fct, unsigned ctas) {
  constexpr int N = R1 * R2 * R3, TT = N / E, M1 = N / R1;
  std::vector<cld> a((size_t)R1 * M1), b((size_t)R2 * R3), r((size_t)N + 1);
  for (int k1 = 0; k1 < R1; ++k1)
    for (int i1 = 0; i1 < M1; ++i1) a[(size_t)k1 * M1 + i1] = root((uint64_t)i1 * k1, N);
  for (int k2 = 0; k2 < R2; ++k2)
    for (int i2 = 0; i2 < R3; ++i2) b[(size_t)k2 * R3 + i2] = root((uint64_t)R1 * i2 * k2, N);
  for (int k = 0; k <= N; ++k) r[k] = root(k, 2 * (uint64_t)N);   // W_2N^k, the real post/pre-twiddle
  auto tw1 = conv<T>(a), tw2 = conv<T>(b), twr = conv<T>(r);
  unsigned sched[2] = {0u, 0u};
  gridDim.x = ctas;
  pthread_barrier_init(&g_bar, nullptr, TT);
  for (unsigned c = 0; c < ctas; ++c) {
    blockIdx.x = c;
    std::memset(smem_raw, 0xCD, sizeof(smem_raw));   // poison: reads of unwritten slots show up as garbage
    std::vector<std::thread> th;
    for (int t = 0; t < TT; ++t)
      th.emplace_back([&, t] {
        threadIdx.x = (unsigned)t;
        fast3_kernel<T, R1, R2, R3, E, KIND, BWD, 1, PAIR, PF, DB, TMA>(in, out, nrows, rs_in, rs_out, tw1.data(), tw2.data(), twr.data(), (T)fct, sched);
      });
    for (auto &x : th) x.join();
  }
  pthread_barrier_destroy(&g_bar);
}

template <typename T, int R1, int R2, int R3, int E>
int dispatch(int kind, int bwd, int flags, const void *in, void *out, uint64_t nrows, int64_t rs_in, int64_t rs_out, double fct, unsigned ctas) {
  const bool pair = (flags & 1) != 0, pf = (flags & 2) != 0, db = (flags & 4) != 0;   // flags: 1 = pair units, 2 = register prefetch, 4 = second exchange buffer
  const bool tma = (flags & 8) != 0;                                                  //        8 = next row staged in shared memory by a bulk copy
#define GOT(K, P, D) (bwd ? run<T, R1, R2, R3, E, K, true, P, false, D, true>(in, out, nrows, rs_in, rs_out, fct, ctas) \
                          : run<T, R1, R2, R3, E, K, false, P, false, D, true>(in, out, nrows, rs_in, rs_out, fct, ctas))
  if (tma) {
    if (pf) return -2;
    if (kind == F3_C2C) { if (db) GOT(F3_C2C, false, true); else GOT(F3_C2C, false, false); return 0; }
    if (!pair) return -2;
    if (kind == F3_R2C) { if constexpr ((R1 * R2) % 2 == 0) { if (db) GOT(F3_R2C, true, true); else GOT(F3_R2C, true, false); return 0; } return -2; }
    if constexpr ((R2 * R3) % 2 == 0) { if (db) GOT(F3_C2R, true, true); else GOT(F3_C2R, true, false); return 0; }
    return -2;
  }
#undef GOT
#define GO(K, B, P, F) run<T, R1, R2, R3, E, K, B, P, F>(in, out, nrows, rs_in, rs_out, fct, ctas)
#define GO2(K, P, F) (bwd ? GO(K, true, P, F) : GO(K, false, P, F))
#define GODB(K, P) (bwd ? run<T, R1, R2, R3, E, K, true, P, false, true>(in, out, nrows, rs_in, rs_out, fct, ctas) \
                        : run<T, R1, R2, R3, E, K, false, P, false, true>(in, out, nrows, rs_in, rs_out, fct, ctas))
  if (db) {
    if (pf) return -2;
    if (kind == F3_C2C) { GODB(F3_C2C, false); return 0; }
    if (!pair) return -2;
    if (kind == F3_R2C) { if constexpr ((R1 * R2) % 2 == 0) { GODB(F3_R2C, true); return 0; } return -2; }
    if constexpr ((R2 * R3) % 2 == 0) { GODB(F3_C2R, true); return 0; }
    return -2;
  }
  if (kind == F3_R2C) {
    if (pair) {
      if constexpr ((R1 * R2) % 2 == 0) { if (pf) GO2(F3_R2C, true, true); else GO2(F3_R2C, true, false); return 0; }
      return -2;
    }
    if (pf) return -2;
    GO2(F3_R2C, false, false);
  } else if (kind == F3_C2R) {
    if (pair) {
      if (pf) return -2;   // (no register prefetch for c2r)
      if constexpr ((R2 * R3) % 2 == 0) { GO2(F3_C2R, true, false); return 0; }
      return -2;
    }
    if (pf) return -2;
    GO2(F3_C2R, false, false);
  } else {
    if (pf) GO2(F3_C2C, false, true); else GO2(F3_C2C, false, false);
  }
#undef GODB
#undef GO2
#undef GO
  return 0;
}
}  // namespace

extern "C" {
// shape = R1*1000000 + R2*10000 + R3*100 + E; dtype 1 = f64, 0 = f32; kind 0 c2c / 1 r2c / 2 c2r; row strides in
// elements of the row's own type (reals for the real side), as LineJob::bs_in/bs_out
int emu_fast3(int shape, int dtype, int kind, int bwd, int flags, const void *in, void *out, uint64_t nrows, int64_t rs_in,
              int64_t rs_out, double fct, unsigned ctas) {
#define SHAPE(A, B, C, D)                                                                                                   \
  if (shape == A * 1000000 + B * 10000 + C * 100 + D)                                                                       \
    return dtype ? dispatch<double, A, B, C, D>(kind, bwd, flags, in, out, nrows, rs_in, rs_out, fct, ctas)                  \
                 : dispatch<float, A, B, C, D>(kind, bwd, flags, in, out, nrows, rs_in, rs_out, fct, ctas);
  SHAPE(16, 16, 8, 16)
  SHAPE(16, 8, 8, 16)
  SHAPE(8, 8, 4, 8)
  SHAPE(8, 8, 8, 8)
  SHAPE(10, 10, 5, 10)
  SHAPE(18, 18, 6, 18)
  SHAPE(5, 10, 10, 10)
  SHAPE(6, 18, 18, 18)
  SHAPE(16, 16, 16, 16)
  SHAPE(8, 16, 16, 16)
  SHAPE(8, 8, 16, 16)
  SHAPE(8, 8, 8, 16)
  SHAPE(4, 8, 8, 8)
  SHAPE(10, 10, 10, 10)
  SHAPE(8, 24, 8, 24)
  SHAPE(10, 20, 10, 20)
  SHAPE(10, 20, 20, 20)
  SHAPE(27, 9, 9, 27)
  SHAPE(10, 30, 10, 30)
  SHAPE(27, 27, 9, 27)
#undef SHAPE
  return -1;
}
}
