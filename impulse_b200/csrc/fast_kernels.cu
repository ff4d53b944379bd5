// Register-resident batched c2c kernels for the headline power-of-two row lengths (sm_100a).
//
// Where the generic engine (fft_kernels.cu) round-trips every radix pass through shared memory,
// these kernels keep a whole row in the registers of TPR threads and touch shared memory only
// for the transposes between passes:
//
//   two-pass  N = R1*R2 (N <= 1024 in fp64, <= 4096 in fp32), TPR = R2 threads per row, E = R1
//             elements per thread, one exchange, warp-synchronous (a row never leaves its warp):
//     LDG.128 x[i + R2*j]            (coalesced: lanes walk i)
//     radix-R1 DFT in registers, twiddle W_N^(i*k1) from a per-CTA shared table
//     STS  S[k1*(R2+1) + i]  |  __syncwarp  |  LDS  S[k1*(R2+1) + j]     (both conflict-free)
//     radix-R2 DFT in registers, STG.128 X[k1 + R1*k2]  (coalesced: lanes walk k1)
//
// HBM traffic is exactly one read and one write per element; the grid is persistent
// (CTAs/SM x 148) and each warp streams rows.  Replaces pass_all + pass2..pass11 of
// pocketfft.c:300-929 for these lengths; same mathematics (DIF Stockham), fp64 twiddles rounded
// from long double on the host.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>

#include "col_device.cuh"
#include "fast3_device.cuh"
#include "fast_common.h"
#include "fastblue_device.cuh"
#include "fft_device.cuh"
#include "fft_kernels.h"
#include "tma_device.cuh"

namespace impulse {



template <typename T, int R1, int R2, int WARPS, int MINB, bool BWD>
__global__ void __launch_bounds__(WARPS * 32, MINB)
fast2_kernel(const cx<T> *__restrict__ in, cx<T> *__restrict__ out, uint64_t nrows, int64_t rs_in, int64_t rs_out,
             const cx<T> *__restrict__ twN, T fct) {
  // R2 threads per row, GPW = floor(32 / R2) rows per warp: R2 need not divide 32 (100 = 10*10: three rows on 30 lanes,
  // 243 = 27*9: three rows on 27 lanes, 625 = 25*25: one row on 25 lanes); the spare lanes idle
  constexpr int N = R1 * R2, TPR = R2, GPW = 32 / TPR, NB2 = R1 / R2, PITCH = R2 + 1;
  static_assert(R2 <= 32 && R1 % R2 == 0, "two-pass shape");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cx<T> *tw = reinterpret_cast<cx<T> *>(smem_raw);
  cx<T> *xbuf = tw + N;
  griddep_launch_dependents();
  for (int idx = threadIdx.x; idx < N; idx += WARPS * 32) {
    const int k1 = idx / R2, i = idx % R2;
    tw[idx] = twN[k1 * i];
  }
  __syncthreads();
  griddep_wait();   // rows may still be written by the previous kernel of the stream (programmatic dependent launch)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool lane_ok = (32 % TPR == 0) || lane < GPW * TPR;
  const int g = lane_ok ? lane / TPR : 0, i = lane_ok ? lane % TPR : 0;
  cx<T> *S = xbuf + (size_t)(warp * GPW + g) * (R1 * PITCH);
  constexpr uint64_t RPC = (uint64_t)WARPS * GPW;  // rows per CTA per sweep
  for (uint64_t wrow = (uint64_t)blockIdx.x * RPC + (uint64_t)warp * GPW; wrow < nrows; wrow += (uint64_t)gridDim.x * RPC) {
    const uint64_t row = wrow + g;
    const bool active = lane_ok && row < nrows;
    cx<T> x[R1];
    {
      const cx<T> *src = in + (int64_t)row * rs_in + i;
#pragma unroll
      for (int j = 0; j < R1; ++j) {
        x[j] = active ? src[j * R2] : mk<T>((T)0, (T)0);
        if (BWD) x[j].y = -x[j].y;
      }
    }
    RegFFT<T, R1>::run(x);
#pragma unroll
    for (int k1 = 1; k1 < R1; ++k1) x[k1] = cmul(x[k1], tw[k1 * R2 + i]);
    if (lane_ok) {
#pragma unroll
      for (int k1 = 0; k1 < R1; ++k1) S[k1 * PITCH + i] = x[k1];
    }
    __syncwarp();
    cx<T> *dst = out + (int64_t)row * rs_out;
#pragma unroll
    for (int m = 0; m < NB2; ++m) {
      const int k1 = i + R2 * m;
      cx<T> y[R2];
#pragma unroll
      for (int j = 0; j < R2; ++j) y[j] = S[k1 * PITCH + j];
      RegFFT<T, R2>::run(y);
      if (active) {
#pragma unroll
        for (int k2 = 0; k2 < R2; ++k2) {
          cx<T> v = y[k2];
          v.x *= fct;
          v.y *= BWD ? -fct : fct;
          dst[k1 + R1 * k2] = v;
        }
      }
    }
    __syncwarp();
  }
}

namespace {
template <typename T, int R1, int R2, int WARPS, int MINB>
int launch_fast2(const LineJob &J, int sm_count, cudaStream_t s) {
  constexpr int N = R1 * R2, GPW = 32 / R2;
  const size_t smem = sizeof(cx<T>) * ((size_t)N + (size_t)WARPS * GPW * R1 * (R2 + 1));
  const bool bwd = (J.flags & F_CONJ_SEQ) != 0;
  auto kf = fast2_kernel<T, R1, R2, WARPS, MINB, false>;
  auto kb = fast2_kernel<T, R1, R2, WARPS, MINB, true>;
  static PerDeviceFlag flag;
  bool &configured = flag.here();
  if (!configured) {
    for (auto k : {kf, kb}) {
      cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return (int)e;
      e = cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
      if (e != cudaSuccess) return (int)e;
    }
    configured = true;
  }
  const uint64_t rpc = (uint64_t)WARPS * GPW;
  uint64_t grid = (J.n_lines + rpc - 1) / rpc;
  const uint64_t cap = (uint64_t)sm_count * MINB;
  if (grid > cap) grid = cap;
  cudaError_t le = launch_pdl(bwd ? kb : kf, (unsigned)grid, (unsigned)(WARPS * 32), smem, s, (const cx<T> *)J.in, (cx<T> *)J.out, (uint64_t)J.n_lines,
                              (int64_t)J.bs_in[0], (int64_t)J.bs_out[0], (const cx<T> *)J.tw, (T)J.fct);
  return le != cudaSuccess ? (int)le : (int)cudaGetLastError();
}
}  // namespace


// Short REAL rows: r2c / c2r of N = 2M points with M = R1*R2 in 16...128, several rows per warp.  The packed
// row (x[2m], x[2m+1]) is an M-point complex line for the two-pass scheme of fast2_kernel; the Hermitian
// post-twiddle (r2c) / pre-twiddle (c2r) pairs bin k with bin M-k through the group's shared buffer.
constexpr int F2R_R2C = 1, F2R_C2R = 2;
template <typename T, int R1, int R2, int WARPS, int MINB, int KIND, bool BWD>
__global__ void __launch_bounds__(WARPS * 32, MINB)
fast2r_kernel(const void *__restrict__ in_v, void *__restrict__ out_v, uint64_t nrows, int64_t rs_in, int64_t rs_out,
              const cx<T> *__restrict__ twN, const cx<T> *__restrict__ twr, T fct) {
  constexpr int M = R1 * R2, TPR = R2, GPW = 32 / TPR, NB2 = R1 / R2, PITCH = R2 + 1, BUF = R1 * PITCH;
  static_assert(R2 <= 32 && 32 % R2 == 0 && R1 % R2 == 0 && BUF >= M + 1, "two-pass shape");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cx<T> *tw = reinterpret_cast<cx<T> *>(smem_raw);
  cx<T> *xbuf = tw + M;
  for (int idx = threadIdx.x; idx < M; idx += WARPS * 32) {
    const int k1 = idx / R2, i = idx % R2;
    tw[idx] = twN[k1 * i];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane / TPR, i = lane % TPR;
  cx<T> *S = xbuf + (size_t)(warp * GPW + g) * BUF;
  constexpr uint64_t RPC = (uint64_t)WARPS * GPW;
  for (uint64_t wrow = (uint64_t)blockIdx.x * RPC + (uint64_t)warp * GPW; wrow < nrows; wrow += (uint64_t)gridDim.x * RPC) {
    const uint64_t row = wrow + g;
    const bool active = row < nrows;
    cx<T> x[R1];
    if (KIND == F2R_R2C) {
      const cx<T> *src = reinterpret_cast<const cx<T> *>(reinterpret_cast<const T *>(in_v) + (int64_t)row * rs_in) + i;
#pragma unroll
      for (int j = 0; j < R1; ++j) x[j] = active ? src[j * R2] : mk<T>((T)0, (T)0);
    } else {
      const cx<T> *src = reinterpret_cast<const cx<T> *>(in_v) + (int64_t)row * rs_in;
      for (int n = i; n <= M; n += TPR) S[n] = active ? src[n] : mk<T>((T)0, (T)0);
      __syncwarp();
#pragma unroll
      for (int j = 0; j < R1; ++j) {
        const int n = i + R2 * j;
        cx<T> a = S[n], b = S[M - n];
        if (BWD) { a.y = -a.y; b.y = -b.y; }       // c2r with forward=true conjugates its input
        if (n == 0) { a.y = (T)0; b.y = (T)0; }    // imaginary parts of bins 0 and M are ignored
        const cx<T> w = cconj(__ldg(twr + n));
        const cx<T> s = cadd(a, cconj(b)), d = csub(a, cconj(b));
        x[j] = cconj(cadd(s, mul_pi(cmul(w, d))));  // backward = conj(FFT(conj z))
      }
      __syncwarp();
    }
    RegFFT<T, R1>::run(x);
#pragma unroll
    for (int k1 = 1; k1 < R1; ++k1) x[k1] = cmul(x[k1], tw[k1 * R2 + i]);
#pragma unroll
    for (int k1 = 0; k1 < R1; ++k1) S[k1 * PITCH + i] = x[k1];
    __syncwarp();
#pragma unroll
    for (int m = 0; m < NB2; ++m) {
      const int k1 = i + R2 * m;
      cx<T> y[R2];
#pragma unroll
      for (int j = 0; j < R2; ++j) y[j] = S[k1 * PITCH + j];
      RegFFT<T, R2>::run(y);
#pragma unroll
      for (int k2 = 0; k2 < R2; ++k2) x[m * R2 + k2] = y[k2];   // bin k1 + R1*k2
    }
    __syncwarp();
    if (KIND == F2R_C2R) {
      cx<T> *dst = reinterpret_cast<cx<T> *>(reinterpret_cast<T *>(out_v) + (int64_t)row * rs_out);
      if (active) {
#pragma unroll
        for (int m = 0; m < NB2; ++m)
#pragma unroll
          for (int k2 = 0; k2 < R2; ++k2) {
            cx<T> v = x[m * R2 + k2];
            v.x *= fct; v.y *= -fct;                  // undo the conjugation of the backward trick
            dst[i + R2 * m + R1 * k2] = v;            // (x[2n], x[2n+1])
          }
      }
    } else {
#pragma unroll
      for (int m = 0; m < NB2; ++m)
#pragma unroll
        for (int k2 = 0; k2 < R2; ++k2) S[i + R2 * m + R1 * k2] = x[m * R2 + k2];
      __syncwarp();
      cx<T> *dst = reinterpret_cast<cx<T> *>(out_v) + (int64_t)row * rs_out;
      const T h = (T)0.5;
      if (active) {
        for (int n = i; n <= M; n += TPR) {
          const cx<T> a = S[n == M ? 0 : n], b = cconj(S[n == 0 ? 0 : M - n]);
          const cx<T> Ev = mk<T>((a.x + b.x) * h, (a.y + b.y) * h), Dv = mk<T>((a.x - b.x) * h, (a.y - b.y) * h);
          cx<T> v = cadd(Ev, cmul(__ldg(twr + n), mul_mi(Dv)));
          v.x *= fct; v.y *= BWD ? -fct : fct;       // r2c with forward=false returns the conjugate spectrum
          dst[n] = v;
        }
      }
    }
    __syncwarp();
  }
}

namespace {
template <typename T, int R1, int R2, int WARPS, int MINB>
int launch_fast2r(const LineJob &J, int sm_count, cudaStream_t s) {
  constexpr int M = R1 * R2, GPW = 32 / R2;
  const size_t smem = sizeof(cx<T>) * ((size_t)M + (size_t)WARPS * GPW * R1 * (R2 + 1));
  const int kind = J.store_mode == ST_R2C_EVEN ? F2R_R2C : F2R_C2R;
  const bool bwd = kind == F2R_R2C ? (J.flags & F_CONJ_RESULT) != 0 : (J.flags & F_CONJ_IN) != 0;
  typedef void (*kern_t)(const void *, void *, uint64_t, int64_t, int64_t, const cx<T> *, const cx<T> *, T);
  kern_t k = kind == F2R_R2C ? (bwd ? (kern_t)fast2r_kernel<T, R1, R2, WARPS, MINB, F2R_R2C, true> : (kern_t)fast2r_kernel<T, R1, R2, WARPS, MINB, F2R_R2C, false>)
                             : (bwd ? (kern_t)fast2r_kernel<T, R1, R2, WARPS, MINB, F2R_C2R, true> : (kern_t)fast2r_kernel<T, R1, R2, WARPS, MINB, F2R_C2R, false>);
  static PerDeviceFlag flags[4];
  bool &configured = flags[(kind == F2R_R2C ? 0 : 2) + (bwd ? 1 : 0)].here();
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  const uint64_t rpc = (uint64_t)WARPS * GPW;
  uint64_t grid = (J.n_lines + rpc - 1) / rpc;
  const uint64_t cap = (uint64_t)sm_count * MINB;
  if (grid > cap) grid = cap;
  k<<<(unsigned)grid, WARPS * 32, smem, s>>>(J.in, J.out, J.n_lines, J.bs_in[0], J.bs_out[0], (const cx<T> *)J.tw,
                                             (const cx<T> *)J.tw_r, (T)J.fct);
  return (int)cudaGetLastError();
}
}  // namespace


// Two-pass kernel with TMA row prefetch.  Each warp owns two shared buffers: while it computes
// row r out of buffer b (first as the staged input, then as the exchange buffer), one elected
// lane has already issued a cp.async.bulk for row r+1 into buffer b^1.  Global-load latency is
// therefore hidden by one whole row of work per warp, independent of register pressure.
template <typename T, int R1, int R2, int WARPS, bool BWD>
__global__ void __launch_bounds__(WARPS * 32, 1)
fast2p_kernel(const cx<T> *__restrict__ in, cx<T> *__restrict__ out, uint64_t nrows, int64_t rs_in, int64_t rs_out,
              const cx<T> *__restrict__ twN, T fct, unsigned int *__restrict__ sched /* [0]=next row, [1]=CTAs done */) {
  constexpr int N = R1 * R2, TPR = R2, GPW = 32 / TPR, NB2 = R1 / R2, PITCH = R2 + 1;
  constexpr int BUF = R1 * PITCH;  // elements per group buffer (>= N)
  static_assert(R2 <= 32 && 32 % R2 == 0 && R1 % R2 == 0, "two-pass shape");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  cx<T> *tw = reinterpret_cast<cx<T> *>(smem_raw);
  cx<T> *bufs = tw + N;                                             // [WARPS][2][GPW][BUF]
  uint64_t *bars = reinterpret_cast<uint64_t *>(bufs + (size_t)WARPS * 2 * GPW * BUF);  // [WARPS][2]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  griddep_launch_dependents();
  if (lane == 0) { mbar_init(&bars[warp * 2], 1); mbar_init(&bars[warp * 2 + 1], 1); }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  for (int idx = threadIdx.x; idx < N; idx += WARPS * 32) {
    const int k1 = idx / R2, i = idx % R2;
    tw[idx] = twN[k1 * i];
  }
  __syncthreads();
  griddep_wait();   // rows may still be written by the previous kernel of the stream (programmatic dependent launch)
  const int g = lane / TPR, i = lane % TPR;
  cx<T> *wbuf = bufs + (size_t)warp * 2 * GPW * BUF;
  uint64_t *bar = &bars[warp * 2];
  constexpr uint64_t RPC = (uint64_t)WARPS * GPW;
  constexpr uint32_t ROW_BYTES = N * sizeof(cx<T>);
  auto issue = [&](uint64_t first_row, int b) {  // lane 0 only
    uint32_t nact = 0;
#pragma unroll
    for (int q = 0; q < GPW; ++q) nact += (first_row + q < nrows) ? 1u : 0u;
    fence_proxy_async();
    mbar_expect_tx(&bar[b], nact * ROW_BYTES);
#pragma unroll
    for (int q = 0; q < GPW; ++q)
      if (first_row + q < nrows)
        bulk_g2s(wbuf + ((size_t)b * GPW + q) * BUF, in + (int64_t)(first_row + q) * rs_in, ROW_BYTES, &bar[b]);
  };
  // Rows are claimed dynamically (SMs do not run at equal speed: a static split left the slowest
  // SM 1.5x behind the fastest).  The claim for iteration it+2 is issued at iteration it, so the
  // atomic's latency never sits on the critical path.
  uint32_t cur = 0, nxt = 0, nxt2 = 0;
  if (lane == 0) { cur = atomicAdd(&sched[0], (unsigned)GPW); nxt = atomicAdd(&sched[0], (unsigned)GPW); }
  cur = __shfl_sync(0xffffffffu, cur, 0);
  nxt = __shfl_sync(0xffffffffu, nxt, 0);
  if (cur < nrows && lane == 0) issue(cur, 0);
  uint32_t ph0 = 0u, ph1 = 0u;
  for (uint32_t it = 0; cur < nrows; ++it) {
    const int b = it & 1;
    const uint64_t wrow = cur;
    if (lane == 0) {
      nxt2 = atomicAdd(&sched[0], (unsigned)GPW);
      if (nxt < nrows) issue(nxt, b ^ 1);
    }
    mbar_wait(&bar[b], b ? ph1 : ph0);
    if (b) ph1 ^= 1u; else ph0 ^= 1u;
    const uint64_t row = wrow + g;
    const bool active = row < nrows;
    cx<T> *S = wbuf + ((size_t)b * GPW + g) * BUF;
    cx<T> x[R1];
#pragma unroll
    for (int j = 0; j < R1; ++j) {
      x[j] = S[i + j * R2];
      if (BWD) x[j].y = -x[j].y;
    }
    __syncwarp();
    RegFFT<T, R1>::run(x);
#pragma unroll
    for (int k1 = 1; k1 < R1; ++k1) x[k1] = cmul(x[k1], tw[k1 * R2 + i]);
#pragma unroll
    for (int k1 = 0; k1 < R1; ++k1) S[k1 * PITCH + i] = x[k1];
    __syncwarp();
    cx<T> *dst = out + (int64_t)row * rs_out;
#pragma unroll
    for (int m = 0; m < NB2; ++m) {
      const int k1 = i + R2 * m;
      cx<T> y[R2];
#pragma unroll
      for (int j = 0; j < R2; ++j) y[j] = S[k1 * PITCH + j];
      RegFFT<T, R2>::run(y);
      if (active) {
#pragma unroll
        for (int k2 = 0; k2 < R2; ++k2) {
          cx<T> v = y[k2];
          v.x *= fct;
          v.y *= BWD ? -fct : fct;
          dst[k1 + R1 * k2] = v;
        }
      }
    }
    __syncwarp();
    cur = nxt;
    nxt = __shfl_sync(0xffffffffu, nxt2, 0);
  }
  // the last CTA to leave re-arms the scheduler words for the next launch that uses this slot
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned done = atomicAdd(&sched[1], 1u);
    if (done == gridDim.x - 1) { sched[0] = 0u; sched[1] = 0u; __threadfence(); }
  }
}

// scheduler slots: zero-initialised device words, one pair per in-flight launch (ring); shared with fast3_kernels.cu.
// The ring of the current device is allocated and zeroed by init_sched_slots(), which the ABI layer calls once per
// device under its context lock and BEFORE any launch (cudaMemset + cudaDeviceSynchronize: the zeros are visible to
// every stream, including non-blocking ones, and nothing is allocated lazily under a CUDA-graph capture).
namespace {
constexpr int kSchedSlots = 65536;  // a slot is reused after this many launches: far more than can be queued across streams
unsigned int *g_sched_base[kMaxDevices] = {};
unsigned g_sched_next = 0;
}  // namespace
int init_sched_slots() {
  unsigned int *&base = g_sched_base[cur_dev()];
  if (base) return 0;
  unsigned int *p = nullptr;
  cudaError_t e = cudaMalloc(&p, sizeof(unsigned int) * 2 * kSchedSlots);
  if (e != cudaSuccess) return (int)e;
  e = cudaMemset(p, 0, sizeof(unsigned int) * 2 * kSchedSlots);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { cudaFree(p); return (int)e; }
  __atomic_store_n(&base, p, __ATOMIC_RELEASE);
  return 0;
}
unsigned int *sched_slot() {
  unsigned int *base = __atomic_load_n(&g_sched_base[cur_dev()], __ATOMIC_ACQUIRE);
  if (!base) return nullptr;
  const unsigned s = __atomic_fetch_add(&g_sched_next, 1u, __ATOMIC_RELAXED) % kSchedSlots;
  return base + 2 * s;
}

namespace {
template <typename T, int R1, int R2, int WARPS>
int launch_fast2p(const LineJob &J, int sm_count, cudaStream_t s) {
  constexpr int N = R1 * R2, GPW = 32 / R2, BUF = R1 * (R2 + 1);
  const size_t smem = sizeof(cx<T>) * ((size_t)N + (size_t)WARPS * 2 * GPW * BUF) + sizeof(uint64_t) * WARPS * 2;
  const bool bwd = (J.flags & F_CONJ_SEQ) != 0;
  auto kf = fast2p_kernel<T, R1, R2, WARPS, false>;
  auto kb = fast2p_kernel<T, R1, R2, WARPS, true>;
  static PerDeviceFlag flag;
  bool &configured = flag.here();
  if (!configured) {
    for (auto k : {kf, kb}) {
      cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return (int)e;
      e = cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
      if (e != cudaSuccess) return (int)e;
    }
    configured = true;
  }
  const uint64_t rpc = (uint64_t)WARPS * GPW;
  uint64_t grid = (J.n_lines + rpc - 1) / rpc;
  if (grid > (uint64_t)sm_count) grid = (uint64_t)sm_count;
  unsigned int *sched = sched_slot();
  if (!sched) return (int)cudaErrorMemoryAllocation;
  if (J.n_lines > 0xfff00000ull) return (int)cudaErrorInvalidValue;  // 32-bit row claims
  cudaError_t le = launch_pdl(bwd ? kb : kf, (unsigned)grid, (unsigned)(WARPS * 32), smem, s, (const cx<T> *)J.in, (cx<T> *)J.out, (uint64_t)J.n_lines,
                              (int64_t)J.bs_in[0], (int64_t)J.bs_out[0], (const cx<T> *)J.tw, (T)J.fct, sched);
  return le != cudaSuccess ? (int)le : (int)cudaGetLastError();
}
}  // namespace

// fast_id encodes the specialised kernel chosen by the planner (0 = generic engine)



// Measured on the 64 x 4096 x 4096 filter (fp32: 128-thread CTAs) the prefetch is worth 12-25% per column
// pass; on the 8192 x 8192 complex128 transform (64-thread CTAs, twice as many resident) it costs 7%.
// IMPULSE_FFT_COL_PREFETCH=0/1 forces it off/on for A/B runs.
static inline bool col_prefetch_enabled(bool f32) {
  static const int mode = [] { const char *e = getenv("IMPULSE_FFT_COL_PREFETCH"); return e ? atoi(e) : -1; }();
  return mode < 0 ? f32 : mode != 0;
}
// grid for the column kernels (see col_group); returns false if the job cannot be launched
static inline bool col_grid(const LineJob &J, int lpc, dim3 *grid) {
  const uint64_t g0n = (J.bdim[0] + lpc - 1) / lpc, groups = g0n * J.bdim[1] * J.bdim[2];
  if (groups == 0 || groups > 0x7fffffffull) return false;
  static const int force1d = [] { const char *e = getenv("IMPULSE_FFT_COL_GRID1D"); return e ? atoi(e) : 0; }();
  if (!force1d && J.bdim[1] <= 65535 && J.bdim[2] <= 65535) *grid = dim3((unsigned)g0n, (unsigned)J.bdim[1], (unsigned)J.bdim[2]);
  else *grid = dim3((unsigned)groups, 1, 1);
  return true;
}


static inline uint32_t col_pipe_groups() {  // IMPULSE_FFT_COL_PIPE=G (0 or 1 turns the pipelined variant off)
  static const int g = [] { const char *e = getenv("IMPULSE_FFT_COL_PIPE"); return e ? atoi(e) : 2; }();
  return g > 1 ? (uint32_t)g : 0u;
}

namespace {
template <typename T, int R1, int R2, int LPC>
int launch_colfast2(const LineJob &J, cudaStream_t s) {
  const bool rows_in = J.col_in_rows != 0;
  // (64- and 128-point sub-transforms only: at 256 points the second register set costs more occupancy than
  // the overlap returns — measured 1.80 -> 1.64 TB/s on 65536-point rows)
  static const int pipe_f32 = [] { const char *e = getenv("IMPULSE_FFT_COL_PIPE_F32"); return e ? atoi(e) : 0; }();   // groups per CTA for complex64 (0 = off)
  if ((sizeof(T) == 8 || pipe_f32 > 1) && R1 * R2 <= 128 && !rows_in && !J.seg_len && !J.umul_mod && col_pipe_groups() && J.bdim[0] >= 2 * LPC) {
    const uint32_t G = sizeof(T) == 8 ? col_pipe_groups() : (uint32_t)pipe_f32;
    const size_t smem_p = sizeof(cx<T>) * (size_t)R1 * R2 * LPC;
    const bool bwd_p = (J.flags & F_CONJ_SEQ) != 0;
    auto kpf = colpipe2_kernel<T, (R1 <= 16 ? R1 : 16), R2, LPC, false>;
    auto kpb = colpipe2_kernel<T, (R1 <= 16 ? R1 : 16), R2, LPC, true>;
    static PerDeviceFlag pflag;
    bool &pconf = pflag.here();
    if (!pconf) {
      for (auto k : {kpf, kpb}) {
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_p);
        if (e != cudaSuccess) return (int)e;
        e = cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) return (int)e;
      }
      pconf = true;
    }
    const uint64_t g0n = (J.bdim[0] + LPC - 1) / LPC, chunks = (g0n + G - 1) / G, groups = chunks * J.bdim[1] * J.bdim[2];
    if (groups == 0 || groups > 0x7fffffffull) return (int)cudaErrorInvalidValue;
    dim3 grid = (J.bdim[1] <= 65535 && J.bdim[2] <= 65535) ? dim3((unsigned)chunks, (unsigned)J.bdim[1], (unsigned)J.bdim[2])
                                                           : dim3((unsigned)groups, 1, 1);
    g_last_kernel = "colpipe2_kernel";
    (bwd_p ? kpb : kpf)<<<grid, LPC * R2, smem_p, s>>>(J, G);
    return (int)cudaGetLastError();
  }
  const size_t smem = sizeof(cx<T>) * (rows_in ? (size_t)LPC * (R1 * R2 + 1) : (size_t)R1 * R2 * LPC);
  const bool bwd = (J.flags & F_CONJ_SEQ) != 0;
  const bool pf = !rows_in && sizeof(T) == 4 && col_prefetch_enabled(true);   // fp64: measured slower with it
  // strided lines without segmented input or a fused multiply run the slimmed (PLAIN) instantiation
  static const int plain_on = [] { const char *e = getenv("IMPULSE_FFT_COL_PLAIN"); return e ? atoi(e) : 1; }();
  const bool plain = plain_on && !rows_in && !J.seg_len && !J.umul_mod;
  typedef void (*kern_t)(const LineJob);
  kern_t kf, kb;
  if (rows_in) { kf = colfast2_kernel<T, R1, R2, LPC, false, false, true>; kb = colfast2_kernel<T, R1, R2, LPC, true, false, true>; }
  else if (plain && pf) { kf = colfast2_kernel<T, R1, R2, LPC, false, sizeof(T) == 4, false, true>; kb = colfast2_kernel<T, R1, R2, LPC, true, sizeof(T) == 4, false, true>; }
  else if (plain) { kf = colfast2_kernel<T, R1, R2, LPC, false, false, false, true>; kb = colfast2_kernel<T, R1, R2, LPC, true, false, false, true>; }
  else if (pf) { kf = colfast2_kernel<T, R1, R2, LPC, false, sizeof(T) == 4, false>; kb = colfast2_kernel<T, R1, R2, LPC, true, sizeof(T) == 4, false>; }
  else { kf = colfast2_kernel<T, R1, R2, LPC, false, false, false>; kb = colfast2_kernel<T, R1, R2, LPC, true, false, false>; }
  static PerDeviceFlag flags[5];
  bool &configured = flags[rows_in ? 2 : (plain ? 3 : 0) + (pf ? 1 : 0)].here();
  if (!configured) {
    for (auto k : {kf, kb}) {
      cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return (int)e;
      e = cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
      if (e != cudaSuccess) return (int)e;
    }
    configured = true;
  }
  dim3 grid;
  if (!col_grid(J, LPC, &grid)) return (int)cudaErrorInvalidValue;
  (bwd ? kb : kf)<<<grid, LPC * R2, smem, s>>>(J);
  return (int)cudaGetLastError();
}
}  // namespace


namespace {
template <typename T, int R1, int R2, int LPC>
int launch_colconv2(const LineJob &J, cudaStream_t s) {
  const size_t smem = 2 * sizeof(cx<T>) * (size_t)R1 * R2 * LPC;
  auto k = col_prefetch_enabled(sizeof(T) == 4) ? colconv2_kernel<T, R1, R2, LPC, true> : colconv2_kernel<T, R1, R2, LPC, false>;
  static PerDeviceFlag flag;
  bool &configured = flag.here();
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  if (!J.umul || !J.umul_mod || !J.tw4_n) return (int)cudaErrorInvalidValue;
  dim3 grid;
  if (!col_grid(J, LPC, &grid)) return (int)cudaErrorInvalidValue;
  k<<<grid, LPC * R2, smem, s>>>(J);
  return (int)cudaGetLastError();
}
}  // namespace



// IMPULSE_FFT_FAST4=1: c2c rows of 8192 points (fp64) on the four-pass 512-thread core instead of the three-pass
// 256-thread kernel.  Validated under the thread-level emulation only, not yet measured: off by default.
static int fast4_enabled() {
  static int v = -1;
  if (v < 0) { const char *e = getenv("IMPULSE_FFT_FAST4"); v = e ? atoi(e) : 0; }
  return v;
}
static int launch_fast4_8192(const LineJob &J, int sm_count, cudaStream_t s) {
  using F = Fft4_8192<double>;
  const size_t smem = sizeof(cx<double>) * ((size_t)F::BUFN + 16 * 32) + 16;
  const bool bwd = (J.flags & F_CONJ_SEQ) != 0;
  auto k = bwd ? fast4_8192_kernel<double, true> : fast4_8192_kernel<double, false>;
  static PerDeviceFlag flags[2];
  bool &configured = flags[bwd ? 1 : 0].here();
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  uint64_t grid = J.n_lines;
  if (grid > (uint64_t)sm_count) grid = (uint64_t)sm_count;
  unsigned int *sched = sched_slot();
  if (!sched) return (int)cudaErrorMemoryAllocation;
  if (J.n_lines > 0xfff00000ull) return (int)cudaErrorInvalidValue;
  g_last_kernel = "fast4_8192_kernel<double>";
  k<<<(unsigned)grid, F::TT, smem, s>>>((const cx<double> *)J.in, (cx<double> *)J.out, J.n_lines, J.bs_in[0], J.bs_out[0],
                                        (const cx<double> *)J.f3_tw1, (const cx<double> *)J.f3_tw2, J.fct, sched);
  return (int)cudaGetLastError();
}

static int fast2p_wide() {
  static int v = -1;
  if (v < 0) { const char *e = getenv("IMPULSE_FFT_FAST2P_WIDE"); v = e ? atoi(e) : 1; }   // measured: 256 points 4.53 -> 6.63 TB/s, 128 points 4.54 -> 5.86
  return v;
}
static int fast_variant() {  // IMPULSE_FFT_FAST_VARIANT=1 selects the non-TMA kernel (A/B measurements)
  static int v = -1;
  if (v < 0) { const char *e = getenv("IMPULSE_FFT_FAST_VARIANT"); v = e ? atoi(e) : 0; }
  return v;
}

// 16 adjacent complex128 lines per CTA (256-byte runs) for the 64- and 128-point column kernels: measured
// 1.51 -> 1.43 ms on the 8192 x 8192 transform, 1.36 ms together with the pipelined variant.
// IMPULSE_FFT_COL_LPC16=0 restores 8 lines for A/B runs.
static int col_lpc16() {
  static int v = -1;
  if (v < 0) { const char *e = getenv("IMPULSE_FFT_COL_LPC16"); v = e ? atoi(e) : 1; }
  return v;
}

int launch_fast_job(const LineJob &J, int sm_count, void *stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (J.fast_id == FAST3_8192_F64 && J.store_mode != ST_R2C_EVEN && J.load_mode != LD_HERM_EVEN && fast4_enabled())
    return launch_fast4_8192(J, sm_count, s);
  switch (J.fast_id) {
    case FAST2_1024_F64:
      if (fast_variant() == 0) { g_last_kernel = "fast2p_kernel<double,32,32,6>"; return launch_fast2p<double, 32, 32, 6>(J, sm_count, s); }
      g_last_kernel = "fast2_kernel<double,32,32,4,2>";
      return launch_fast2<double, 32, 32, 4, 2>(J, sm_count, s);
    case FAST2_512_F64:
      if (fast_variant() == 0) { g_last_kernel = "fast2p_kernel<double,32,16,6>"; return launch_fast2p<double, 32, 16, 6>(J, sm_count, s); }
      g_last_kernel = "fast2_kernel<double,32,16,4,2>";
      return launch_fast2<double, 32, 16, 4, 2>(J, sm_count, s);
    case FAST2_256_F64:
      // 32 points per thread, 8 threads per row, four rows (16 KB) per warp iteration — the per-warp shape of the
      // 1024- and 512-point kernels — against 16 x 16 (IMPULSE_FFT_FAST2P_WIDE=0 restores it)
      if (fast_variant() == 0 && fast2p_wide()) { g_last_kernel = "fast2p_kernel<double,32,8,6>"; return launch_fast2p<double, 32, 8, 6>(J, sm_count, s); }
      if (fast_variant() == 0) { g_last_kernel = "fast2p_kernel<double,16,16,8>"; return launch_fast2p<double, 16, 16, 8>(J, sm_count, s); }
      g_last_kernel = "fast2_kernel<double,16,16,4,4>";
      return launch_fast2<double, 16, 16, 4, 4>(J, sm_count, s);
    case FAST2_1024_F32:
      // measured: the TMA variant (12 warps, one CTA per SM) reaches 4.60 TB/s here against 5.19 TB/s for
      // four 4-warp CTAs per SM — complex64 rows are instruction-issue bound, not latency bound
      g_last_kernel = "fast2_kernel<float,32,32,4,4>";
      return launch_fast2<float, 32, 32, 4, 4>(J, sm_count, s);
    case FAST2_16_F64:   // short rows: plain loads beat the bulk copies (5.2-5.4 against 2.0-3.1 TB/s measured)
      g_last_kernel = "fast2_kernel<double,4,4,8,4>";
      return launch_fast2<double, 4, 4, 8, 4>(J, sm_count, s);
    case FAST2_32_F64:   // short rows: plain loads beat the bulk copies (5.2-5.4 against 2.0-3.1 TB/s measured)
      g_last_kernel = "fast2_kernel<double,8,4,8,4>";
      return launch_fast2<double, 8, 4, 8, 4>(J, sm_count, s);
    case FAST2_64_F64:   // short rows: plain loads beat the bulk copies (5.2-5.4 against 2.0-3.1 TB/s measured)
      g_last_kernel = "fast2_kernel<double,8,8,8,4>";
      return launch_fast2<double, 8, 8, 8, 4>(J, sm_count, s);
    case FAST2_128_F64:
      if (fast_variant() == 0 && fast2p_wide()) { g_last_kernel = "fast2p_kernel<double,32,4,5>"; return launch_fast2p<double, 32, 4, 5>(J, sm_count, s); }
      if (fast_variant() == 0) { g_last_kernel = "fast2p_kernel<double,16,8,10>"; return launch_fast2p<double, 16, 8, 10>(J, sm_count, s); }
      g_last_kernel = "fast2_kernel<double,16,8,8,4>";
      return launch_fast2<double, 16, 8, 8, 4>(J, sm_count, s);
    // non-power-of-two short rows: R2 threads per row, floor(32 / R2) rows per warp
    case FAST2_100_F64: g_last_kernel = "fast2_kernel<double,10,10,8,4>"; return launch_fast2<double, 10, 10, 8, 4>(J, sm_count, s);
    case FAST2_243_F64: g_last_kernel = "fast2_kernel<double,27,9,4,2>"; return launch_fast2<double, 27, 9, 4, 2>(J, sm_count, s);
    case FAST2_625_F64: g_last_kernel = "fast2_kernel<double,25,25,4,2>"; return launch_fast2<double, 25, 25, 4, 2>(J, sm_count, s);
    case FAST2_100_F32: g_last_kernel = "fast2_kernel<float,10,10,8,6>"; return launch_fast2<float, 10, 10, 8, 6>(J, sm_count, s);
    case FAST2_243_F32: g_last_kernel = "fast2_kernel<float,27,9,4,4>"; return launch_fast2<float, 27, 9, 4, 4>(J, sm_count, s);
    case FAST2_625_F32: g_last_kernel = "fast2_kernel<float,25,25,4,4>"; return launch_fast2<float, 25, 25, 4, 4>(J, sm_count, s);
    case FAST2_16_F32: g_last_kernel = "fast2_kernel<float,4,4,8,6>"; return launch_fast2<float, 4, 4, 8, 6>(J, sm_count, s);
    case FAST2_32_F32: g_last_kernel = "fast2_kernel<float,8,4,8,6>"; return launch_fast2<float, 8, 4, 8, 6>(J, sm_count, s);
    case FAST2_64_F32: g_last_kernel = "fast2_kernel<float,8,8,8,6>"; return launch_fast2<float, 8, 8, 8, 6>(J, sm_count, s);
    case FAST2_128_F32: g_last_kernel = "fast2_kernel<float,16,8,8,4>"; return launch_fast2<float, 16, 8, 8, 4>(J, sm_count, s);
    case FAST2_256_F32: g_last_kernel = "fast2_kernel<float,16,16,8,4>"; return launch_fast2<float, 16, 16, 8, 4>(J, sm_count, s);
    case FAST2_512_F32: g_last_kernel = "fast2_kernel<float,32,16,4,4>"; return launch_fast2<float, 32, 16, 4, 4>(J, sm_count, s);
    case FAST2_50_F64: g_last_kernel = "fast2_kernel<double,10,5,8,4>"; return launch_fast2<double, 10, 5, 8, 4>(J, sm_count, s);
    case FAST2_50_F32: g_last_kernel = "fast2_kernel<float,10,5,8,6>"; return launch_fast2<float, 10, 5, 8, 6>(J, sm_count, s);
    case FAST2_72_F64: g_last_kernel = "fast2_kernel<double,24,3,4,2>"; return launch_fast2<double, 24, 3, 4, 2>(J, sm_count, s);
    case FAST2_72_F32: g_last_kernel = "fast2_kernel<float,24,3,4,4>"; return launch_fast2<float, 24, 3, 4, 4>(J, sm_count, s);
    case FAST2_81_F64: g_last_kernel = "fast2_kernel<double,9,9,8,4>"; return launch_fast2<double, 9, 9, 8, 4>(J, sm_count, s);
    case FAST2_81_F32: g_last_kernel = "fast2_kernel<float,9,9,8,6>"; return launch_fast2<float, 9, 9, 8, 6>(J, sm_count, s);
    case FAST2_96_F64: g_last_kernel = "fast2_kernel<double,24,4,4,2>"; return launch_fast2<double, 24, 4, 4, 2>(J, sm_count, s);
    case FAST2_96_F32: g_last_kernel = "fast2_kernel<float,24,4,4,4>"; return launch_fast2<float, 24, 4, 4, 4>(J, sm_count, s);
    case FAST2_192_F64: g_last_kernel = "fast2_kernel<double,24,8,4,2>"; return launch_fast2<double, 24, 8, 4, 2>(J, sm_count, s);
    case FAST2_192_F32: g_last_kernel = "fast2_kernel<float,24,8,4,4>"; return launch_fast2<float, 24, 8, 4, 4>(J, sm_count, s);
    case FAST2_200_F64: g_last_kernel = "fast2_kernel<double,20,10,4,3>"; return launch_fast2<double, 20, 10, 4, 3>(J, sm_count, s);
    case FAST2_200_F32: g_last_kernel = "fast2_kernel<float,20,10,4,4>"; return launch_fast2<float, 20, 10, 4, 4>(J, sm_count, s);
    case FAST2_400_F64: g_last_kernel = "fast2_kernel<double,20,20,4,3>"; return launch_fast2<double, 20, 20, 4, 3>(J, sm_count, s);
    case FAST2_400_F32: g_last_kernel = "fast2_kernel<float,20,20,4,4>"; return launch_fast2<float, 20, 20, 4, 4>(J, sm_count, s);
    case FAST2_576_F64: g_last_kernel = "fast2_kernel<double,24,24,4,2>"; return launch_fast2<double, 24, 24, 4, 2>(J, sm_count, s);
    case FAST2_576_F32: g_last_kernel = "fast2_kernel<float,24,24,4,4>"; return launch_fast2<float, 24, 24, 4, 4>(J, sm_count, s);
    case FAST2_729_F64: g_last_kernel = "fast2_kernel<double,27,27,4,2>"; return launch_fast2<double, 27, 27, 4, 2>(J, sm_count, s);
    case FAST2_729_F32: g_last_kernel = "fast2_kernel<float,27,27,4,4>"; return launch_fast2<float, 27, 27, 4, 4>(J, sm_count, s);
    case FAST2_900_F64: g_last_kernel = "fast2_kernel<double,30,30,4,2>"; return launch_fast2<double, 30, 30, 4, 2>(J, sm_count, s);
    case FAST2_900_F32: g_last_kernel = "fast2_kernel<float,30,30,4,4>"; return launch_fast2<float, 30, 30, 4, 4>(J, sm_count, s);
    case FAST2_8_F64: g_last_kernel = "fast2_kernel<double,4,2,8,4>"; return launch_fast2<double, 4, 2, 8, 4>(J, sm_count, s);
    case FAST2_4_F64: g_last_kernel = "fast2_kernel<double,2,2,8,4>"; return launch_fast2<double, 2, 2, 8, 4>(J, sm_count, s);
    case FAST2_8_F32: g_last_kernel = "fast2_kernel<float,4,2,8,6>"; return launch_fast2<float, 4, 2, 8, 6>(J, sm_count, s);
    case FAST2_4_F32: g_last_kernel = "fast2_kernel<float,2,2,8,6>"; return launch_fast2<float, 2, 2, 8, 6>(J, sm_count, s);
    case FAST2R_8_F64: g_last_kernel = "fast2r_kernel<double,4,2>"; return launch_fast2r<double, 4, 2, 8, 4>(J, sm_count, s);
    case FAST2R_4_F64: g_last_kernel = "fast2r_kernel<double,2,2>"; return launch_fast2r<double, 2, 2, 8, 4>(J, sm_count, s);
    case FAST2R_8_F32: g_last_kernel = "fast2r_kernel<float,4,2>"; return launch_fast2r<float, 4, 2, 8, 6>(J, sm_count, s);
    case FAST2R_4_F32: g_last_kernel = "fast2r_kernel<float,2,2>"; return launch_fast2r<float, 2, 2, 8, 6>(J, sm_count, s);
    case FAST2R_16_F64: g_last_kernel = "fast2r_kernel<double,4,4>"; return launch_fast2r<double, 4, 4, 8, 4>(J, sm_count, s);
    case FAST2R_32_F64: g_last_kernel = "fast2r_kernel<double,8,4>"; return launch_fast2r<double, 8, 4, 8, 4>(J, sm_count, s);
    case FAST2R_64_F64: g_last_kernel = "fast2r_kernel<double,8,8>"; return launch_fast2r<double, 8, 8, 8, 4>(J, sm_count, s);
    case FAST2R_128_F64: g_last_kernel = "fast2r_kernel<double,16,8>"; return launch_fast2r<double, 16, 8, 8, 3>(J, sm_count, s);
    case FAST2R_16_F32: g_last_kernel = "fast2r_kernel<float,4,4>"; return launch_fast2r<float, 4, 4, 8, 6>(J, sm_count, s);
    case FAST2R_32_F32: g_last_kernel = "fast2r_kernel<float,8,4>"; return launch_fast2r<float, 8, 4, 8, 6>(J, sm_count, s);
    case FAST2R_64_F32: g_last_kernel = "fast2r_kernel<float,8,8>"; return launch_fast2r<float, 8, 8, 8, 6>(J, sm_count, s);
    case FAST2R_128_F32: g_last_kernel = "fast2r_kernel<float,16,8>"; return launch_fast2r<float, 16, 8, 8, 4>(J, sm_count, s);
    case FASTBLUE_2048_F64: case FASTBLUE_4096_F64: case FASTBLUE_8192_F64:
    case FASTBLUE_2048_F32: case FASTBLUE_4096_F32: case FASTBLUE_8192_F32:
      return launch_fastblue_job(J, sm_count, stream);   // fastblue_kernels.cu
    case COLCONVW_512_F32: case COLCONVW_1024_F32: case COLCONVW_2048_F32: case COLCONVW_4096_F32:
    case COLCONVW_512_F64: case COLCONVW_1024_F64: case COLCONVW_2048_F64: case COLCONVW_4096_F64:
    case COLW_1024_F32: case COLW_2048_F32: case COLW_1024_F64: case COLW_2048_F64:
      return launch_colconvw_job(J, sm_count, stream);   // colconvw_kernels.cu
    case COLCONV_32_F64: g_last_kernel = "colconv2_kernel<double,8,4,8>"; return launch_colconv2<double, 8, 4, 8>(J, s);
    case COLCONV_64_F64: g_last_kernel = "colconv2_kernel<double,8,8,8>"; return launch_colconv2<double, 8, 8, 8>(J, s);
    case COLCONV_128_F64: g_last_kernel = "colconv2_kernel<double,16,8,8>"; return launch_colconv2<double, 16, 8, 8>(J, s);
    case COLCONV_32_F32: g_last_kernel = "colconv2_kernel<float,8,4,16>"; return launch_colconv2<float, 8, 4, 16>(J, s);
    case COLCONV_64_F32: g_last_kernel = "colconv2_kernel<float,8,8,16>"; return launch_colconv2<float, 8, 8, 16>(J, s);
    case COLCONV_128_F32: g_last_kernel = "colconv2_kernel<float,16,8,16>"; return launch_colconv2<float, 16, 8, 16>(J, s);
    case COL2_32_F64: g_last_kernel = "colfast2_kernel<double,8,4,8>"; return launch_colfast2<double, 8, 4, 8>(J, s);
    case COL2_512_F64: g_last_kernel = "colfast2_kernel<double,32,16,8>"; return launch_colfast2<double, 32, 16, 8>(J, s);
    case COL2_32_F32: g_last_kernel = "colfast2_kernel<float,8,4,16>"; return launch_colfast2<float, 8, 4, 16>(J, s);
    case COL2_512_F32: g_last_kernel = "colfast2_kernel<float,32,16,16>"; return launch_colfast2<float, 32, 16, 16>(J, s);
    case COL2_64_F64:
      if (col_lpc16() && J.bdim[0] >= 64) { g_last_kernel = "colfast2_kernel<double,8,8,16>"; return launch_colfast2<double, 8, 8, 16>(J, s); }
      g_last_kernel = "colfast2_kernel<double,8,8,8>"; return launch_colfast2<double, 8, 8, 8>(J, s);
    case COL2_128_F64:
      if (col_lpc16() && J.bdim[0] >= 64) { g_last_kernel = "colfast2_kernel<double,16,8,16>"; return launch_colfast2<double, 16, 8, 16>(J, s); }
      g_last_kernel = "colfast2_kernel<double,16,8,8>"; return launch_colfast2<double, 16, 8, 8>(J, s);
    case COL2_256_F64: g_last_kernel = "colfast2_kernel<double,16,16,8>"; return launch_colfast2<double, 16, 16, 8>(J, s);
    case COL2_64_F32: g_last_kernel = "colfast2_kernel<float,8,8,16>"; return launch_colfast2<float, 8, 8, 16>(J, s);
    case COL2_128_F32: g_last_kernel = "colfast2_kernel<float,16,8,16>"; return launch_colfast2<float, 16, 8, 16>(J, s);
    case COL2_256_F32: g_last_kernel = "colfast2_kernel<float,16,16,16>"; return launch_colfast2<float, 16, 16, 16>(J, s);
    default: return launch_fast3_job(J, sm_count, stream);   // three-pass register kernels: fast3_kernels.cu
  }
}

}  // namespace impulse
