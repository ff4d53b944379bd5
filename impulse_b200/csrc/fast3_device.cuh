// Three-pass register kernel (fast3_kernel) and the in-register DFT helpers shared by every register kernel.
// Split out of fast_kernels.cu so that tests/emu/emu_fast3.cpp can compile the SAME kernel body for the host
// (one OS thread per CUDA thread, __syncthreads() = a barrier) and check its index algebra in the GPU-less
// build container.  The emulation is test infrastructure; the product runs this code on the device only.
#pragma once
#include "fft_device.cuh"
#include "tma_device.cuh"

namespace impulse {

// v * exp(-2*pi*i*M/R) with the trivial roots folded at compile time
template <typename T, int R, int M> __device__ __forceinline__ cx<T> mul_root(cx<T> v) {
  constexpr int m = ((M % R) + R) % R;
  constexpr T h = (T)0.7071067811865475244008444;
  if constexpr (m == 0) return v;
  else if constexpr (4 * m == R) return mul_mi(v);
  else if constexpr (2 * m == R) return mk<T>(-v.x, -v.y);
  else if constexpr (4 * m == 3 * R) return mul_pi(v);
  else if constexpr (8 * m == R) return mk<T>((v.x + v.y) * h, (v.y - v.x) * h);
  else if constexpr (8 * m == 3 * R) return mk<T>((v.y - v.x) * h, -(v.x + v.y) * h);
  else if constexpr (8 * m == 5 * R) return mk<T>(-(v.x + v.y) * h, (v.x - v.y) * h);
  else if constexpr (8 * m == 7 * R) return mk<T>((v.x - v.y) * h, (v.x + v.y) * h);
  else {
    constexpr T c = (T)Trig<R>::c(m), s = (T)Trig<R>::s(m);
    return mk<T>(v.x * c + v.y * s, v.y * c - v.x * s);
  }
}

// v * exp(-2*pi*i*m/R) for an m that is a compile-time constant after unrolling
template <typename T, int R, int M = 0> struct RootSel {
  static __device__ __forceinline__ cx<T> run(cx<T> v, int m) { return m == M ? mul_root<T, R, M>(v) : RootSel<T, R, M + 1>::run(v, m); }
};
template <typename T, int R> struct RootSel<T, R, R> {
  static __device__ __forceinline__ cx<T> run(cx<T> v, int) { return v; }
};

// forward DFT of R points held in registers, natural order in and out
template <typename T, int R> struct RegFFT {
  static __device__ __forceinline__ void run(cx<T> (&x)[R]) { Bfly<T, R>::run(x); }
};

template <typename T, int RA, int RB> struct Composite {
  static constexpr int R = RA * RB;
  template <int J, int S> static __device__ __forceinline__ void tw_row(cx<T> (&x)[R], const cx<T> (&y)[RA]) {
    if constexpr (S < RA) {
      x[J + RB * S] = mul_root<T, R, J * S>(y[S]);
      tw_row<J, S + 1>(x, y);
    }
  }
  template <int J> static __device__ __forceinline__ void stage_a(cx<T> (&x)[R]) {
    if constexpr (J < RB) {
      cx<T> y[RA];
#pragma unroll
      for (int q = 0; q < RA; ++q) y[q] = x[J + RB * q];
      RegFFT<T, RA>::run(y);
      tw_row<J, 0>(x, y);
      stage_a<J + 1>(x);
    }
  }
  static __device__ __forceinline__ void run(cx<T> (&x)[R]) {
    stage_a<0>(x);
    cx<T> out[R];
#pragma unroll
    for (int s = 0; s < RA; ++s) {
      cx<T> z[RB];
#pragma unroll
      for (int j = 0; j < RB; ++j) z[j] = x[j + RB * s];
      RegFFT<T, RB>::run(z);
#pragma unroll
      for (int r = 0; r < RB; ++r) out[s + RA * r] = z[r];
    }
#pragma unroll
    for (int k = 0; k < R; ++k) x[k] = out[k];
  }
};
template <typename T> struct RegFFT<T, 16> { static __device__ __forceinline__ void run(cx<T> (&x)[16]) { Composite<T, 4, 4>::run(x); } };
template <typename T> struct RegFFT<T, 32> { static __device__ __forceinline__ void run(cx<T> (&x)[32]) { Composite<T, 4, 8>::run(x); } };
template <typename T> struct RegFFT<T, 9> { static __device__ __forceinline__ void run(cx<T> (&x)[9]) { Composite<T, 3, 3>::run(x); } };
template <typename T> struct RegFFT<T, 6> { static __device__ __forceinline__ void run(cx<T> (&x)[6]) { Composite<T, 2, 3>::run(x); } };
template <typename T> struct RegFFT<T, 10> { static __device__ __forceinline__ void run(cx<T> (&x)[10]) { Composite<T, 2, 5>::run(x); } };
template <typename T> struct RegFFT<T, 18> { static __device__ __forceinline__ void run(cx<T> (&x)[18]) { Composite<T, 2, 9>::run(x); } };
template <typename T> struct RegFFT<T, 20> { static __device__ __forceinline__ void run(cx<T> (&x)[20]) { Composite<T, 4, 5>::run(x); } };
template <typename T> struct RegFFT<T, 24> { static __device__ __forceinline__ void run(cx<T> (&x)[24]) { Composite<T, 3, 8>::run(x); } };
template <typename T> struct RegFFT<T, 25> { static __device__ __forceinline__ void run(cx<T> (&x)[25]) { Composite<T, 5, 5>::run(x); } };
template <typename T> struct RegFFT<T, 27> { static __device__ __forceinline__ void run(cx<T> (&x)[27]) { Composite<T, 3, 9>::run(x); } };
template <typename T> struct RegFFT<T, 30> { static __device__ __forceinline__ void run(cx<T> (&x)[30]) { Composite<T, 5, 6>::run(x); } };
template <typename T> struct RegFFT<T, 64> { static __device__ __forceinline__ void run(cx<T> (&x)[64]) { Composite<T, 8, 8>::run(x); } };

// =================================================================================================
// Three-pass register kernel: N = R1*R2*R3 complex points per row, T = N/E threads per row each
// holding E points, two shared-memory exchanges, one row per CTA at a time, rows claimed dynamically.
// (index formulas validated by tools/model_fast3.py)
//
//   load   x[t + T*q]                                                   coalesced LDG.128
//   pass 1 radix R1 over j1 (regs q = m + (E/R1)*j1), twiddle tw1[k1][i1], i1 = t + T*m
//   X1[k1*P1 + i1]            P1 = roundup(N/R1, S) + 1   -> pass-2 reads conflict-free
//   pass 2 butterflies b2 = t + T*m2: k1 = b2 % R1, i2 = b2 / R1; radix R2 over X1[k1][i2 + R3*j2];
//          twiddle tw2[k2][i2]
//   X2[i2*P2 + k1 + R1*k2]    P2 = R1*R2 (aliases X1 after a barrier)
//   pass 3 butterflies klow = t + T*m3: radix R3 over X2[j3][klow] -> X[klow + R1*R2*k3]   coalesced STG.128
//
// KIND: 0 = c2c; 1 = r2c (row of 2N reals viewed as N complex, Hermitian post-twiddle through shared
// memory, N+1 bins out); 2 = c2r (N+1 bins in, pre-twiddle through shared memory, 2N reals out).
// =================================================================================================
enum { F3_C2C = 0, F3_R2C = 1, F3_C2R = 2 };

__device__ __forceinline__ void prefetch_l2_bulk(const void *gptr, uint32_t bytes) {
#if defined(__CUDA_ARCH__)
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gptr), "r"(bytes) : "memory");
#else
  (void)gptr; (void)bytes;
#endif
}

// TMA: the NEXT claimed row is staged into shared memory by one cp.async.bulk (issued by thread 0 right after the first
// exchange, when every thread has consumed the current staged row; completion on an mbarrier) and the threads read
// their points from the staging buffer instead of global memory.  The row's HBM latency is hidden behind passes 2
// and 3 of the row before it, independent of the register budget — what took the two-pass kernel from 61 % to
// 81 % DRAM utilisation (fast2p_kernel).  Needs 16-byte aligned rows (the launcher checks and falls back).
template <typename T, int N, int KIND> struct F3Stage {
  static constexpr uint32_t ROW_BYTES = (uint32_t)((KIND == F3_C2R ? (N + 1) : N) * sizeof(cx<T>));
  static constexpr uint32_t BYTES = (ROW_BYTES + 15u) & ~15u;   // copied (a padded row stride makes the tail readable)
};

template <typename T, int R1, int R2, int R3, int E, int KIND, bool BWD, int MINB, bool PAIR = false, bool PF = false, bool DB = false,
          bool TMA = false>
__global__ void __launch_bounds__((R1 * R2 * R3) / E, MINB)
fast3_kernel(const void *__restrict__ in_v, void *__restrict__ out_v, uint64_t nrows, int64_t rs_in, int64_t rs_out,
             const cx<T> *__restrict__ tw1, const cx<T> *__restrict__ tw2, const cx<T> *__restrict__ twr, T fct,
             unsigned int *__restrict__ sched) {
  // pass-1 twiddles W_N^(t*k1) as a product A[k1>>2]*B[k1&3] of six per-thread values held in registers
  // for the whole kernel (R1 = 16): 15 global table loads per row become 9 multiplies
  constexpr bool TW1_REGS = (R1 == 16);  // with E/R1 > 1 the extra factor W_E^(m*k1) is a compile-time root
  constexpr int N = R1 * R2 * R3, TT = N / E, M1 = N / R1, S = sizeof(T) == 8 ? 8 : 16;
  constexpr int P1 = ((M1 + S - 1) / S) * S + 1, P2 = R1 * R2;
  constexpr int NB1 = E / R1, NB2 = E / R2, NB3 = E / R3;
  static_assert(E % R1 == 0 && E % R2 == 0 && E % R3 == 0 && TT % R1 == 0, "fast3 shape");
  constexpr int BUFN = (R1 * P1 > N + 1) ? R1 * P1 : N + 1;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // DB: a second exchange buffer.  Pass-2 results (X2) no longer overwrite what pass 2 is still reading (X1), and the
  // next row's pass 1 no longer overwrites what pass 3 is still reading, so two of the four barriers per row go:
  // what remains is "X1 complete" and "X2 complete".  (Direct-load variants only: the shared-memory Hermitian
  // twiddles use the buffer a third time.)
  static_assert(!DB || KIND == F3_C2C || PAIR, "the second exchange buffer needs the direct-load variants");
  cx<T> *buf = reinterpret_cast<cx<T> *>(smem_raw);
  cx<T> *buf2 = DB ? buf + BUFN : buf;
  unsigned int *s_row = reinterpret_cast<unsigned int *>(buf + (DB ? 2 : 1) * BUFN);  // [2] (+2 pad)
  cx<T> *s_tw2 = reinterpret_cast<cx<T> *>(s_row + 4);                 // [R2][R3]
  static_assert(!TMA || !PF, "staged rows replace the register prefetch");
  static_assert(!TMA || KIND != F3_C2R || PAIR, "the staged c2r needs the direct-load (pair) variant");
  // staging buffer of the next row + its mbarrier, behind the tables (128-byte aligned)
  unsigned char *stg_raw = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(s_tw2 + R2 * R3) + 127) & ~(uintptr_t)127);
  const cx<T> *stg = reinterpret_cast<const cx<T> *>(stg_raw);
  uint64_t *s_bar = reinterpret_cast<uint64_t *>(stg_raw + F3Stage<T, R1 * R2 * R3, KIND>::BYTES);
  const int t = threadIdx.x;
  griddep_launch_dependents();   // the next launch of the stream may start its own prologue while this grid runs
  if (TMA && t == 0) { mbar_init(s_bar, 1); mbar_init_fence(); }
  if (t == 0) { s_row[0] = atomicAdd(&sched[0], 1u); s_row[1] = atomicAdd(&sched[0], 1u); }
  for (int idx = t; idx < R2 * R3; idx += TT) s_tw2[idx] = tw2[idx];
  cx<T> twA[3], twB[3];
  if (TW1_REGS) {
#pragma unroll
    for (int a = 1; a < 4; ++a) { twA[a - 1] = tw1[(4 * a) * M1 + t]; twB[a - 1] = tw1[a * M1 + t]; }
  }
  __syncthreads();
  // everything above read plan-time tables and this launch's own scheduler words; rows may still be written by the
  // previous kernel of the stream (programmatic dependent launch): wait for it here
  griddep_wait();
  constexpr uint32_t ROW_BYTES_IN = (KIND == F3_C2R ? (N + 1) : N) * sizeof(cx<T>);
  const int k1 = t % R1, i2b = t / R1;  // pass-2 ownership (T % R1 == 0)
  const cx<T> wt = KIND == F3_C2C ? mk<T>((T)1, (T)0) : __ldg(twr + t);   // real kinds: W_2N^t, this thread's twiddle factor
  // PF: the NEXT claimed row is requested into registers right after the second exchange, when the row in flight
  // no longer needs them, so that its global-load latency hides behind pass 3 and the stores.
  // (c2c and paired r2c only: for the paired c2r, whose loads sit inside the pass-1 units, hoisting them measured
  // 2-13 % SLOWER on the B200 — more live registers, spills on the 18-point shapes.)
  static_assert(!PF || KIND == F3_C2C || (KIND == F3_R2C && PAIR), "register prefetch needs the direct-load variants");
  cx<T> x[E];
  auto row_ptr = [&](const uint64_t r) -> const cx<T> * {
    return KIND == F3_R2C ? reinterpret_cast<const cx<T> *>(reinterpret_cast<const T *>(in_v) + (int64_t)r * rs_in)
                          : reinterpret_cast<const cx<T> *>(in_v) + (int64_t)r * rs_in;
  };
  auto stage_row = [&](const uint64_t r) {   // thread 0 only; every reader of the staging buffer is behind a barrier
    fence_proxy_async();
    mbar_expect_tx(s_bar, F3Stage<T, R1 * R2 * R3, KIND>::BYTES);
    bulk_g2s(stg_raw, row_ptr(r), F3Stage<T, R1 * R2 * R3, KIND>::BYTES, s_bar);
  };
  if (TMA && t == 0 && s_row[0] < nrows) stage_row(s_row[0]);
  uint32_t stg_phase = 0u;
  auto load_row = [&](const uint64_t r) {
    if constexpr (KIND != F3_C2R) {
      const cx<T> *src = TMA ? stg : row_ptr(r);
#pragma unroll
      for (int q = 0; q < E; ++q) {
        x[q] = src[t + TT * q];
        if (KIND == F3_C2C && BWD) x[q].y = -x[q].y;
      }
    }
  };
  bool loaded = false;
  for (unsigned it = 0;; ++it) {
    const uint64_t row = s_row[it & 1];
    if (row >= nrows) break;
    if (TMA) { mbar_wait(s_bar, stg_phase); stg_phase ^= 1u; }   // this row has landed in the staging buffer
    if (!TMA && t == 0) {  // pull the next claimed row into L2 while this one is transformed
      const uint64_t nxt = s_row[(it + 1) & 1];
      if (nxt < nrows) {
        const char *p = KIND == F3_R2C ? reinterpret_cast<const char *>(reinterpret_cast<const T *>(in_v) + (int64_t)nxt * rs_in)
                                       : reinterpret_cast<const char *>(reinterpret_cast<const cx<T> *>(in_v) + (int64_t)nxt * rs_in);
        // the bulk prefetch wants 16-byte aligned address and size (float c2r rows are only 8-byte aligned)
        const uintptr_t lo = (reinterpret_cast<uintptr_t>(p) + 15) & ~(uintptr_t)15;
        const uintptr_t hi = (reinterpret_cast<uintptr_t>(p) + ROW_BYTES_IN) & ~(uintptr_t)15;
        if (hi > lo) prefetch_l2_bulk(reinterpret_cast<const void *>(lo), (uint32_t)(hi - lo));
      }
    }
    if (!(PF && loaded)) load_row(row);   // (no-op for c2r, which loads below)
    if constexpr (KIND == F3_C2R && PAIR) {
      // ---------------- load + c2r pre-twiddle in registers + pass 1 ----------------
      // Pass-1 butterfly i1 consumes z[i1 + M1*j]; the mirror of that point, N - n = (M1 - i1) + M1*(R1-1-j),
      // feeds butterfly M1 - i1.  A thread that owns BOTH butterflies of a unit {u, M1 - u} loads X[n] and X[N-n]
      // straight from global memory (ascending / descending runs across the lanes, both coalesced) and forms
      //   z[n] = conj(s + p),  z[N-n] = s - p,   s = X[n] + conj X[N-n],  p = i conj(W_2N^n) (X[n] - conj X[N-n]),
      // with W_2N^n = W_2N^u * W_(2 R1)^j (per-thread factor x compile-time root): no staging buffer, no barrier.
      // Unit 0 = butterflies 0 and M1/2, which mirror into themselves.  (backward = conj(FFT(conj z)).)
      static_assert(M1 % 2 == 0, "pair units need an even N/R1");
      constexpr int UNITS = M1 / 2, NU = (UNITS + TT - 1) / TT;
      const cx<T> *src = TMA ? stg : reinterpret_cast<const cx<T> *>(in_v) + (int64_t)row * rs_in;
      auto pass1 = [&](cx<T> (&y)[R1], const int i1) {
        RegFFT<T, R1>::run(y);
#pragma unroll
        for (int k = 1; k < R1; ++k) y[k] = cmul(y[k], __ldg(tw1 + k * M1 + i1));
#pragma unroll
        for (int k = 0; k < R1; ++k) buf[k * P1 + i1] = y[k];
      };
#pragma unroll
      for (int mu = 0; mu < NU; ++mu) {
        const int u = t + TT * mu;
        if (UNITS % TT != 0 && u >= UNITS) continue;
        const int ka = u, kb = u == 0 ? M1 / 2 : M1 - u;
        cx<T> A[R1], B[R1], xa[R1], xb[R1];
#pragma unroll
        for (int j = 0; j < R1; ++j) { A[j] = src[ka + M1 * j]; B[j] = src[kb + M1 * j]; }
        if (BWD) {                                  // c2r with forward=true conjugates its input
#pragma unroll
          for (int j = 0; j < R1; ++j) { A[j].y = -A[j].y; B[j].y = -B[j].y; }
        }
        if (u != 0) {
          const cx<T> wu = mu == 0 ? wt : __ldg(twr + u);
          const cx<T> cwi = mk<T>(wu.y, wu.x);      // i * conj(W_2N^u)
#pragma unroll
          for (int j = 0; j < R1; ++j) {
            const cx<T> a = A[j], b = B[R1 - 1 - j];
            const cx<T> s = mk<T>(a.x + b.x, a.y - b.y), d = mk<T>(a.x - b.x, a.y + b.y);
            const cx<T> pp = RootSel<T, 2 * R1>::run(cmul(cwi, d), (2 * R1 - j) % (2 * R1));   // * conj(W_(2 R1)^j)
            xa[j] = mk<T>(s.x + pp.x, -(s.y + pp.y));
            xb[R1 - 1 - j] = mk<T>(s.x - pp.x, s.y - pp.y);
          }
        } else {
          {                                         // n = 0 pairs with bin N; their imaginary parts are ignored
            const T a0 = A[0].x, bn = src[N].x;
            xa[0] = mk<T>(a0 + bn, -(a0 - bn));
          }
#pragma unroll
          for (int j = 1; j <= R1 / 2; ++j) {       // butterfly 0: n = M1*j <-> M1*(R1-j)
            const cx<T> a = A[j], b = A[R1 - j], w = __ldg(twr + M1 * j);
            const cx<T> s = mk<T>(a.x + b.x, a.y - b.y), d = mk<T>(a.x - b.x, a.y + b.y);
            const cx<T> pp = cmul(mk<T>(w.y, w.x), d);
            xa[j] = mk<T>(s.x + pp.x, -(s.y + pp.y));
            xa[R1 - j] = mk<T>(s.x - pp.x, s.y - pp.y);
          }
#pragma unroll
          for (int j = 0; j < (R1 + 1) / 2; ++j) {  // butterfly M1/2: n = M1/2 + M1*j <-> M1/2 + M1*(R1-1-j)
            const cx<T> a = B[j], b = B[R1 - 1 - j], w = __ldg(twr + M1 / 2 + M1 * j);
            const cx<T> s = mk<T>(a.x + b.x, a.y - b.y), d = mk<T>(a.x - b.x, a.y + b.y);
            const cx<T> pp = cmul(mk<T>(w.y, w.x), d);
            xb[j] = mk<T>(s.x + pp.x, -(s.y + pp.y));
            xb[R1 - 1 - j] = mk<T>(s.x - pp.x, s.y - pp.y);
          }
        }
        pass1(xa, ka);
        pass1(xb, kb);
      }
    } else {
      // ---------------- load (+ c2r pre-twiddle) ----------------
      if (KIND == F3_C2R) {
        const cx<T> *src = reinterpret_cast<const cx<T> *>(in_v) + (int64_t)row * rs_in;
#pragma unroll
        for (int q = 0; q < E; ++q) buf[t + TT * q] = src[t + TT * q];
        if (t == 0) buf[N] = src[N];
        __syncthreads();
        if (t == 0) s_row[it & 1] = atomicAdd(&sched[0], 1u);  // claim for iteration it+2
#pragma unroll
        for (int q = 0; q < E; ++q) {
          const int n = t + TT * q;
          cx<T> a = buf[n], b = buf[N - n];
          if (BWD) { a.y = -a.y; b.y = -b.y; }       // c2r with forward=true conjugates its input
          if (n == 0) { a.y = (T)0; b.y = (T)0; }    // imaginary parts of bins 0 and N are ignored
          const cx<T> w = cconj(q == 0 ? wt : cmul(wt, __ldg(twr + TT * q)));   // e^{+2 pi i n/(2N)}, n = t + TT*q
          const cx<T> s = cadd(a, cconj(b)), d = csub(a, cconj(b));
          const cx<T> z = cadd(s, mul_pi(cmul(w, d)));
          x[q] = cconj(z);                           // backward = conj(FFT(conj z))
        }
        __syncthreads();
      }
      // ---------------- pass 1 ----------------
#pragma unroll
      for (int m = 0; m < NB1; ++m) {
        cx<T> y[R1];
#pragma unroll
        for (int j = 0; j < R1; ++j) y[j] = x[m + NB1 * j];
        RegFFT<T, R1>::run(y);
        const int i1 = t + TT * m;
        if (TW1_REGS) {
#pragma unroll
          for (int k = 1; k < R1; ++k) {
            const int a = k >> 2, b = k & 3;
            if (a == 0) y[k] = cmul(y[k], twB[b - 1]);
            else if (b == 0) y[k] = cmul(y[k], twA[a - 1]);
            else y[k] = cmul(y[k], cmul(twA[a - 1], twB[b - 1]));
            if (NB1 > 1 && m > 0) y[k] = RootSel<T, E>::run(y[k], (m * k) % E);
          }
        } else {
#pragma unroll
          for (int k = 1; k < R1; ++k) y[k] = cmul(y[k], __ldg(tw1 + k * M1 + i1));
        }
#pragma unroll
        for (int k = 0; k < R1; ++k) buf[k * P1 + i1] = y[k];
      }
    }
    __syncthreads();
    if ((KIND != F3_C2R || PAIR) && t == 0) s_row[it & 1] = atomicAdd(&sched[0], 1u);  // everyone has read s_row[it&1]
    if (TMA && t == 0) {   // ... and its points out of the staging buffer: the next row can stream in behind passes 2 and 3
      const uint64_t nxt = s_row[(it + 1) & 1];
      if (nxt < nrows) stage_row(nxt);
    }
    // ---------------- pass 2 ----------------
#pragma unroll
    for (int m = 0; m < NB2; ++m) {
      const int i2 = i2b + (TT / R1) * m;
      cx<T> y[R2];
#pragma unroll
      for (int j = 0; j < R2; ++j) y[j] = buf[k1 * P1 + i2 + R3 * j];
      RegFFT<T, R2>::run(y);
#pragma unroll
      for (int k = 1; k < R2; ++k) y[k] = cmul(y[k], s_tw2[k * R3 + i2]);
#pragma unroll
      for (int k = 0; k < R2; ++k) x[m * R2 + k] = y[k];
    }
    if (!DB) __syncthreads();   // X2 aliases X1: every pass-2 read first
#pragma unroll
    for (int m = 0; m < NB2; ++m) {
      const int i2 = i2b + (TT / R1) * m;
#pragma unroll
      for (int k = 0; k < R2; ++k) buf2[i2 * P2 + k1 + R1 * k] = x[m * R2 + k];
    }
    __syncthreads();
    if constexpr (PF) {   // x[] (pA/pB) are free from here on: request the next claimed row
      const uint64_t nxt = s_row[(it + 1) & 1];
      loaded = nxt < nrows;
      if (loaded) load_row(nxt);
    }
    if constexpr (KIND == F3_R2C && PAIR) {
      // ---------------- pass 3 with the Hermitian post-twiddle in registers ----------------
      // Butterfly klow leaves Z[klow + P2*k3] in y[k3]; the mirror of that bin, N - k = (P2 - klow) + P2*(R3-1-k3),
      // belongs to butterfly P2 - klow.  A thread that runs BOTH butterflies of a unit u = {u, P2 - u} therefore
      // holds Z[k] and Z[N-k] together:  X[k] = E - Q,  X[N-k] = conj(E + Q),  E = (Z[k] + conj Z[N-k])/2,
      // Q = i W_2N^k (Z[k] - conj Z[N-k])/2,  W_2N^k = W_2N^u * W_(2 R3)^k3 (per-thread factor x compile-time root).
      // No exchange through shared memory, no extra barriers.  Unit 0 = butterflies 0 and P2/2, which mirror
      // into themselves.  (index algebra: tools/model_fast3.py, model_r2c_pair)
      static_assert(P2 % 2 == 0, "pair units need an even R1*R2");
      constexpr int UNITS = P2 / 2, NU = (UNITS + TT - 1) / TT;
      cx<T> *dst = reinterpret_cast<cx<T> *>(out_v) + (int64_t)row * rs_out;
      const T hf = (T)0.5 * fct;
      // one pair: a = Z[k], c = Z[N-k], wq = i * W_2N^k * fct/2
      auto emit = [&](const cx<T> a, const cx<T> c, const cx<T> q, const int k) {
        const cx<T> e = mk<T>(hf * (a.x + c.x), hf * (a.y - c.y));
        cx<T> vk = mk<T>(e.x - q.x, e.y - q.y), vm = mk<T>(e.x + q.x, -(e.y + q.y));
        if (BWD) { vk.y = -vk.y; vm.y = -vm.y; }   // r2c with forward=false returns the conjugate spectrum
        dst[k] = vk;
        dst[N - k] = vm;
      };
#pragma unroll
      for (int mu = 0; mu < NU; ++mu) {
        const int u = t + TT * mu;
        if (UNITS % TT != 0 && u >= UNITS) continue;
        const int ka = u, kb = u == 0 ? P2 / 2 : P2 - u;
        cx<T> ya[R3], yb[R3];
#pragma unroll
        for (int j = 0; j < R3; ++j) { ya[j] = buf2[j * P2 + ka]; yb[j] = buf2[j * P2 + kb]; }
        RegFFT<T, R3>::run(ya);
        RegFFT<T, R3>::run(yb);
        if (u != 0) {
          const cx<T> wu = mu == 0 ? wt : __ldg(twr + u);       // W_2N^u
          const cx<T> wh = mk<T>(-wu.y * hf, wu.x * hf);        // i * W_2N^u * fct/2
#pragma unroll
          for (int k3 = 0; k3 < R3; ++k3) {
            const cx<T> a = ya[k3], c = yb[R3 - 1 - k3];
            const cx<T> d = mk<T>(a.x - c.x, a.y + c.y);        // Z[k] - conj Z[N-k]
            emit(a, c, RootSel<T, 2 * R3>::run(cmul(wh, d), k3), ka + P2 * k3);
          }
        } else {
          {                                                     // bins 0 and N: Re Z0 +- Im Z0
            cx<T> v0 = mk<T>((ya[0].x + ya[0].y) * fct, (T)0), vn = mk<T>((ya[0].x - ya[0].y) * fct, (T)0);
            dst[0] = v0;
            dst[N] = vn;
          }
#pragma unroll
          for (int k3 = 1; k3 <= R3 / 2; ++k3) {                // butterfly 0: bin P2*k3 <-> P2*(R3-k3)
            const cx<T> a = ya[k3], c = ya[R3 - k3], w = __ldg(twr + P2 * k3);
            const cx<T> d = mk<T>(a.x - c.x, a.y + c.y);
            emit(a, c, cmul(mk<T>(-w.y * hf, w.x * hf), d), P2 * k3);
          }
#pragma unroll
          for (int k3 = 0; k3 < (R3 + 1) / 2; ++k3) {           // butterfly P2/2: bin P2/2 + P2*k3 <-> P2/2 + P2*(R3-1-k3)
            const cx<T> a = yb[k3], c = yb[R3 - 1 - k3], w = __ldg(twr + P2 / 2 + P2 * k3);
            const cx<T> d = mk<T>(a.x - c.x, a.y + c.y);
            emit(a, c, cmul(mk<T>(-w.y * hf, w.x * hf), d), P2 / 2 + P2 * k3);
          }
        }
      }
      if (!DB) __syncthreads();  // pass-3 reads done before the next row's pass-1 writes
    } else {
      // ---------------- pass 3 (+ store for c2c / c2r: straight from the butterfly's registers) ----------------
      if constexpr (KIND != F3_R2C) {
        // c2c: X[klow + R1*R2*k];  c2r: (x[2n], x[2n+1]) with the conjugation of the backward trick undone
        cx<T> *dst = KIND == F3_C2C ? reinterpret_cast<cx<T> *>(out_v) + (int64_t)row * rs_out
                                    : reinterpret_cast<cx<T> *>(reinterpret_cast<T *>(out_v) + (int64_t)row * rs_out);
        const T fy = (KIND == F3_C2R || BWD) ? -fct : fct;
#pragma unroll
        for (int m = 0; m < NB3; ++m) {
          const int klow = t + TT * m;
          cx<T> y[R3];
#pragma unroll
          for (int j = 0; j < R3; ++j) y[j] = buf2[j * P2 + klow];
          RegFFT<T, R3>::run(y);
#pragma unroll
          for (int k = 0; k < R3; ++k) dst[klow + R1 * R2 * k] = mk<T>(y[k].x * fct, y[k].y * fy);
        }
        if (!DB) __syncthreads();  // pass-3 reads done before the next row's pass-1 writes
      } else {  // r2c: Hermitian post-twiddle needs Z[k] and Z[N-k]
#pragma unroll
        for (int m = 0; m < NB3; ++m) {
          const int klow = t + TT * m;
          cx<T> y[R3];
#pragma unroll
          for (int j = 0; j < R3; ++j) y[j] = buf[j * P2 + klow];
          RegFFT<T, R3>::run(y);
#pragma unroll
          for (int k = 0; k < R3; ++k) x[m * R3 + k] = y[k];   // Z[klow + R1*R2*k]
        }
        __syncthreads();
#pragma unroll
        for (int m = 0; m < NB3; ++m)
#pragma unroll
          for (int k = 0; k < R3; ++k) buf[t + TT * m + R1 * R2 * k] = x[m * R3 + k];
        __syncthreads();
        cx<T> *dst = reinterpret_cast<cx<T> *>(out_v) + (int64_t)row * rs_out;
        const T h = (T)0.5;
#pragma unroll
        for (int q = 0; q < E; ++q) {
          const int k = t + TT * q;
          const cx<T> a = buf[k], b = cconj(buf[k == 0 ? 0 : N - k]);
          const cx<T> Ev = mk<T>((a.x + b.x) * h, (a.y + b.y) * h), Dv = mk<T>((a.x - b.x) * h, (a.y - b.y) * h);
          // W_2N^k = W^t * W^(TT*q): one per-thread factor and one warp-uniform factor (both L1-resident)
          // instead of a 16-byte table entry per output streamed from L2
          const cx<T> wk = q == 0 ? wt : cmul(wt, __ldg(twr + TT * q));
          cx<T> v = cadd(Ev, cmul(wk, mul_mi(Dv)));
          v.x *= fct; v.y *= BWD ? -fct : fct;       // r2c with forward=false returns the conjugate spectrum
          dst[k] = v;
          if (k == 0) {                               // bin N: Re Z0 - Im Z0
            cx<T> last = mk<T>((a.x - a.y) * fct, (T)0);
            dst[N] = last;
          }
        }
        __syncthreads();
      }
    }
  }
  // the last CTA to leave re-arms the scheduler words
  __syncthreads();
  if (t == 0) {
    __threadfence();
    const unsigned done = atomicAdd(&sched[1], 1u);
    if (done == gridDim.x - 1) { sched[0] = 0u; sched[1] = 0u; __threadfence(); }
  }
}

}  // namespace impulse
