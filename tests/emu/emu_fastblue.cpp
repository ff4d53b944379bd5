// TEST INFRASTRUCTURE — host emulation of the fused Bluestein kernel (fastblue_kernel), thread level.
//
// Same scheme as emu_fast3.cpp: the kernel body (impulse_b200/csrc/fastblue_device.cuh, the source nvcc compiles) built
// with g++, one OS thread per CUDA thread, __syncthreads() = a pthread barrier.  The tables (three-pass twiddles,
// chirp b_k, FFT(b)/M, alias corrections) come from the product's own planner (planner.cpp linked in with a host
// allocator), so table conventions are checked together with the kernel.  NOT a product code path.
#include <pthread.h>

#include <atomic>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

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

#include "../../impulse_b200/csrc/fastblue_device.cuh"
#include "../../impulse_b200/csrc/planner.h"

using namespace impulse;

namespace {
struct HostAlloc : TableAlloc {
  void *upload(const void *h, size_t n) override { void *p = std::malloc(n ? n : 1); if (p) std::memcpy(p, h, n); return p; }
  void release(void *p) override { std::free(p); }
};
HostAlloc g_alloc;
PlanCache *g_cache = nullptr;

template <typename T, int R3, int E, int KIND, bool BWD, bool BKS, bool BFE, bool FOUR = false, bool TMB = false>
int run(uint32_t L, const void *in, void *out, uint64_t nrows, int64_t rs_in, int64_t rs_out, double fct, unsigned ctas) {
  constexpr int M = 16 * 16 * R3, TT = M / E;
  if (!g_cache) g_cache = new PlanCache(&g_alloc);
  std::string err;
  const void *tw1 = nullptr, *tw2 = nullptr, *bf = nullptr, *corr = nullptr;
  uint32_t d = 0;
  const Engine1D *eng = nullptr;
  constexpr int DT = sizeof(T) == 8 ? DT_F64 : DT_F32;
  if (g_cache->fast3_tables(M, 16, 16, R3, DT, &tw1, &tw2, &err)) return -3;
  if (g_cache->fastblue_tables(L, M, DT, &bf, &corr, &d, &err)) return -4;
  if (g_cache->status_engine(L, DT, &eng, &err) || !eng->d_bk) return -5;
  unsigned sched[2] = {0u, 0u};
  gridDim.x = ctas;
  pthread_barrier_init(&g_bar, nullptr, TT);
  for (unsigned c = 0; c < ctas; ++c) {
    blockIdx.x = c;
    std::memset(smem_raw, 0xCD, sizeof(smem_raw));
    std::vector<std::thread> th;
    for (int t = 0; t < TT; ++t)
      th.emplace_back([&, t] {
        threadIdx.x = (unsigned)t;
        fastblue_kernel<T, 16, 16, R3, E, KIND, BWD, BKS, BFE, FOUR, TMB>(in, out, nrows, rs_in, rs_out, L, d, (const cx<T> *)tw1, (const cx<T> *)tw2,
                                                                     (const cx<T> *)eng->d_bk, (const cx<T> *)bf, (const cx<T> *)corr, (T)fct, sched);
      });
    for (auto &x : th) x.join();
  }
  pthread_barrier_destroy(&g_bar);
  return 0;
}

template <typename T, int R3, int E, bool BKS, bool BFE, bool FOUR = false, bool TMB = false>
int by_kind(int kind, int bwd, uint32_t L, const void *in, void *out, uint64_t nrows, int64_t rs_in, int64_t rs_out, double fct, unsigned ctas) {
#define GO(K) (bwd ? run<T, R3, E, K, true, BKS, BFE, FOUR, TMB>(L, in, out, nrows, rs_in, rs_out, fct, ctas) : run<T, R3, E, K, false, BKS, BFE, FOUR, TMB>(L, in, out, nrows, rs_in, rs_out, fct, ctas))
  if (kind == BL_C2C) return GO(BL_C2C);
  if (kind == BL_R2C_PAIR) return GO(BL_R2C_PAIR);
  return GO(BL_C2R_PAIR);
#undef GO
}
}  // namespace

namespace {
template <bool BWD>
int run_fast4(const void *in, void *out, uint64_t nrows, int64_t rs_in, int64_t rs_out, double fct, unsigned ctas) {
  if (!g_cache) g_cache = new PlanCache(&g_alloc);
  std::string err;
  const void *tw1 = nullptr, *tw2 = nullptr;
  if (g_cache->fast3_tables(8192, 16, 16, 32, DT_F64, &tw1, &tw2, &err)) return -3;
  unsigned sched[2] = {0u, 0u};
  gridDim.x = ctas;
  pthread_barrier_init(&g_bar, nullptr, 512);
  for (unsigned c = 0; c < ctas; ++c) {
    blockIdx.x = c;
    std::memset(smem_raw, 0xCD, sizeof(smem_raw));
    std::vector<std::thread> th;
    for (int t = 0; t < 512; ++t)
      th.emplace_back([&, t] {
        threadIdx.x = (unsigned)t;
        fast4_8192_kernel<double, BWD>((const cx<double> *)in, (cx<double> *)out, nrows, rs_in, rs_out, (const cx<double> *)tw1,
                                       (const cx<double> *)tw2, fct, sched);
      });
    for (auto &x : th) x.join();
  }
  pthread_barrier_destroy(&g_bar);
  return 0;
}
}  // namespace

extern "C" {
// c2c rows of 8192 points on the four-pass core (fast4_8192_kernel)
int emu_fast4_8192(int bwd, const void *in, void *out, uint64_t nrows, int64_t rs_in, int64_t rs_out, double fct, unsigned ctas) {
  return bwd ? run_fast4<true>(in, out, nrows, rs_in, rs_out, fct, ctas) : run_fast4<false>(in, out, nrows, rs_in, rs_out, fct, ctas);
}
// kind: 0 c2c, 1 r2c (row pairs), 2 c2r (row pairs); flags: 1 = chirp table in shared memory, 2 = multipliers inside the
// first transform's last pass, 4 = four-pass core (512 threads), 8 = float32, 16 = multipliers in tensor memory (with 2 and 4); row strides in elements of the row's own type (LineJob::bs_in / bs_out)
int emu_fastblue(int kind, int bwd, int flags, uint32_t L, const void *in, void *out, uint64_t nrows, int64_t rs_in, int64_t rs_out,
                 double fct, unsigned ctas) {
  const uint32_t need = 2 * L - 1;
  const bool bks = flags & 1, bfe = flags & 2, f32 = flags & 8;
#define ARGS kind, bwd, L, in, out, nrows, rs_in, rs_out, fct, ctas
  if (f32) {   // float32: the variants the launcher would pick (chirp table in shared memory + early multipliers at 8192)
    if (need <= 2048 + 8) return by_kind<float, 8, 16, false, false>(ARGS);
    if (need <= 4096 + 8) return by_kind<float, 16, 16, false, false>(ARGS);
    if (need > 8192 + 8) return -1;
    if (bks && bfe) return by_kind<float, 32, 32, true, true>(ARGS);
    return by_kind<float, 32, 32, false, false>(ARGS);
  }
  if (need <= 2048 + 8) return by_kind<double, 8, 16, false, false>(ARGS);
  if (need <= 4096 + 8) return by_kind<double, 16, 16, false, false>(ARGS);
  if (need > 8192 + 8) return -1;
  if (flags & 4) {   // four-pass core: 512 threads x 16 points
    if (bfe && (flags & 16)) return by_kind<double, 32, 16, true, true, true, true>(ARGS);   // FFT(b)/M in tensor memory
    if (bfe) return by_kind<double, 32, 16, true, true, true>(ARGS);
    return by_kind<double, 32, 16, true, false, true>(ARGS);
  }
  if (bks && bfe) return by_kind<double, 32, 32, true, true>(ARGS);
  if (bks) return by_kind<double, 32, 32, true, false>(ARGS);
  return by_kind<double, 32, 32, false, false>(ARGS);
#undef ARGS
}
}
