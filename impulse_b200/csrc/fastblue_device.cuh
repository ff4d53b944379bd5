// Fused Bluestein kernel (fastblue_kernel) and its in-CTA three-pass transform (Fft3).  A header of its own so that
// tests/emu/emu_fastblue.cpp can compile the SAME kernel body for the host (thread-level emulation, test
// infrastructure only); the product runs it on the device only.
#pragma once
#include "tmem_device.cuh"
#include "fast3_device.cuh"

namespace impulse {

// =================================================================================================
// Fused Bluestein on the three-pass register core: chirp -> FFT(M) -> x FFT(b)/M -> inverse FFT(M)
// -> chirp, the M-point work array never leaves the SM (registers + one shared buffer).
//
// M is a power of two served by the fast3 shapes (2048/4096/8192).  Bluestein needs a cyclic length
// >= 2L-1; when M falls short by d = 2L-1-M (a few points: L = 4099 -> M = 8192, d = 5) only the lags
// |l| >= L-d alias, which touches outputs k < d through inputs n >= L-d+k.  The wrapped chirp keeps
// the positive lags; the d(d+1)/2 missing products are added back from a small table
// (corr[k][j] = b(n-k) - b(M-n+k), n = L-d+j).  Replaces fftblue_fft (pocketfft.c:1945-2008) and, for
// odd real lengths, rfftblue_* (pocketfft.c:2019-2058) with TWO rows packed into one complex line.
// =================================================================================================
template <typename T, int R1, int R2, int R3, int E>
struct Fft3 {
  static constexpr int N = R1 * R2 * R3, TT = N / E, M1 = N / R1, S = sizeof(T) == 8 ? 8 : 16;
  static constexpr int P1 = ((M1 + S - 1) / S) * S + 1, P2 = R1 * R2;
  static constexpr int NB1 = E / R1, NB2 = E / R2, NB3 = E / R3;
  static constexpr int BUFN = (R1 * P1 > N + 1) ? R1 * P1 : N + 1;
  static constexpr bool TW1_REGS = (R1 == 16);
  static_assert(E % R1 == 0 && E % R2 == 0 && E % R3 == 0 && TT % R1 == 0, "fast3 shape");
  // forward FFT of the N points x[q] <-> n = t + TT*q; result in the same layout (k = t + TT*q).
  // `buf` must be free on entry; on return other threads may still be reading it.
  // MUL: the result is multiplied by mul[k] and conjugated (the middle of a Bluestein convolution); half of each
  // pass-3 butterfly's multipliers are requested BEFORE the butterfly runs, so that their L2 latency hides behind it
  // (the registers of x[] are free at that point; all of them would not fit beside the butterfly).
  template <bool MUL = false>
  static __device__ __forceinline__ void run(cx<T> (&x)[E], cx<T> *buf, const cx<T> *__restrict__ tw1,
                                             const cx<T> *s_tw2, const cx<T> (&twA)[3], const cx<T> (&twB)[3], int t,
                                             const cx<T> *__restrict__ mul = nullptr) {
    const int k1 = t % R1, i2b = t / R1;
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
    __syncthreads();
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
    __syncthreads();
#pragma unroll
    for (int m = 0; m < NB2; ++m) {
      const int i2 = i2b + (TT / R1) * m;
#pragma unroll
      for (int k = 0; k < R2; ++k) buf[i2 * P2 + k1 + R1 * k] = x[m * R2 + k];
    }
    __syncthreads();
#pragma unroll
    for (int m = 0; m < NB3; ++m) {
      const int klow = t + TT * m;
      cx<T> y[R3];
#pragma unroll
      for (int j = 0; j < R3; ++j) y[j] = buf[j * P2 + klow];
      constexpr int PRE = MUL ? (R3 + 1) / 2 : 1;
      cx<T> w[PRE];
      if constexpr (MUL) {
#pragma unroll
        for (int k = 0; k < PRE; ++k) w[k] = __ldg(mul + t + TT * (m + NB3 * k));
      }
      RegFFT<T, R3>::run(y);
      if constexpr (MUL) {
#pragma unroll
        for (int k = 0; k < R3; ++k) {
          const cx<T> v = cmul(y[k], k < PRE ? w[k < PRE ? k : 0] : __ldg(mul + t + TT * (m + NB3 * k)));
          x[m + NB3 * k] = mk<T>(v.x, -v.y);
        }
      } else {
#pragma unroll
        for (int k = 0; k < R3; ++k) x[m + NB3 * k] = y[k];   // X[t + TT*(m + NB3*k)]  (R1*R2 = TT*NB3)
      }
    }
  }
};

// Four-pass variant of the in-CTA transform for the 8192-point work length: 512 threads x 16 points (16 warps per SM
// where Fft3<16,16,32,E32> has 8 — the fused Bluestein kernel is bound by load and shared-memory latency, not by
// arithmetic), radices 16 * 16 * 16 * 2, three exchanges through the one shared buffer.  Same contract as Fft3::run:
// x[q] <-> n = t + 512*q in, k = t + 512*q out; `buf` free on entry, possibly still being read on return.
//   pass 1  butterfly i1 = t:                      x[i1 + 512*j]           -> X1[k1*513 + i1]       * W_N^(i1*k1)
//   pass 2  (k1, i2) = (t % 16, t / 16):           X1[k1][i2 + 32*j]       -> X2[i2*256 + k1+16*k2] * W_N^(16*i2*k2)
//   pass 3  (klow2, i3) = (t % 256, t / 256):      X2[(i3 + 2*j)][klow2]   -> X3[i3*4096 + klow2+256*k3] * W_32^(i3*k3)
//   pass 4  klow3 = t + 512*m (m < 8), radix 2:    X3[0|1][klow3]          -> k = klow3 + 4096*k4 = t + 512*(m + 8*k4)
// Tables: the ones of the 16*16*32 three-pass shape (tw1[16][512], tw2[16][32]).  Every shared-memory access has
// consecutive threads on consecutive elements (pass-2 reads: stride 513 = 1 mod 8) — conflict-free.
template <typename T>
struct Fft4_8192 {
  static constexpr int N = 8192, E = 16, TT = 512, M1 = 512, P1 = 513, BUFN = 16 * P1;
  static constexpr bool TW1_REGS = true;
  // TMB: the thread's 16 multipliers mul[t + 512 q] — the same for every unit — wait in its own tensor-memory columns
  // (tm_mul; written once per CTA) instead of being streamed from L2 for every unit
  template <bool MUL = false, bool TMB = false>
  static __device__ __forceinline__ void run(cx<T> (&x)[16], cx<T> *buf, const cx<T> *__restrict__ /*tw1*/,
                                             const cx<T> *s_tw2, const cx<T> (&twA)[3], const cx<T> (&twB)[3], int t,
                                             const cx<T> *__restrict__ mul = nullptr, uint32_t tm_mul = 0) {
    // ---- pass 1
    RegFFT<T, 16>::run(x);
#pragma unroll
    for (int k = 1; k < 16; ++k) {
      const int a = k >> 2, b = k & 3;
      if (a == 0) x[k] = cmul(x[k], twB[b - 1]);
      else if (b == 0) x[k] = cmul(x[k], twA[a - 1]);
      else x[k] = cmul(x[k], cmul(twA[a - 1], twB[b - 1]));
    }
#pragma unroll
    for (int k = 0; k < 16; ++k) buf[k * P1 + t] = x[k];
    __syncthreads();
    // ---- pass 2
    const int k1 = t % 16, i2 = t / 16;
#pragma unroll
    for (int j = 0; j < 16; ++j) x[j] = buf[k1 * P1 + i2 + 32 * j];
    RegFFT<T, 16>::run(x);
#pragma unroll
    for (int k = 1; k < 16; ++k) x[k] = cmul(x[k], s_tw2[k * 32 + i2]);
    __syncthreads();   // X2 aliases X1
#pragma unroll
    for (int k = 0; k < 16; ++k) buf[i2 * 256 + k1 + 16 * k] = x[k];
    __syncthreads();
    // ---- pass 3
    const int klow2 = t % 256, i3 = t / 256;
#pragma unroll
    for (int j = 0; j < 16; ++j) x[j] = buf[(i3 + 2 * j) * 256 + klow2];
    RegFFT<T, 16>::run(x);
    if (i3) {
#pragma unroll
      for (int k = 1; k < 16; ++k) x[k] = RootSel<T, 32>::run(x[k], k);
    }
    __syncthreads();   // X3 aliases X2
#pragma unroll
    for (int k = 0; k < 16; ++k) buf[i3 * 4096 + klow2 + 256 * k] = x[k];
    // the row registers are free here: with MUL, all 16 multipliers are requested before the barrier and the radix-2 pass
    cx<T> w[MUL ? 16 : 1];
    if constexpr (MUL && TMB) {
      constexpr int WC = (int)(sizeof(cx<T>) / 4);
      uint32_t *ww = reinterpret_cast<uint32_t *>(&w[0]);
#pragma unroll
      for (int c = 0; c < 16 * WC; c += 8) cw_tmem_ld8(tm_mul + (uint32_t)c, ww + c);
      cw_tmem_wait_ld();
    } else if constexpr (MUL) {
#pragma unroll
      for (int q = 0; q < 16; ++q) w[q] = __ldg(mul + t + TT * q);
    }
    __syncthreads();
    // ---- pass 4
#pragma unroll
    for (int m = 0; m < 8; ++m) {
      const cx<T> a = buf[t + TT * m], b = buf[4096 + t + TT * m];
      cx<T> s0 = cadd(a, b), s1 = csub(a, b);
      if constexpr (MUL) {
        s0 = cmul(s0, w[m]); s1 = cmul(s1, w[MUL ? m + 8 : 0]);
        s0.y = -s0.y; s1.y = -s1.y;
      }
      x[m] = s0;
      x[m + 8] = s1;
    }
  }
};

// c2c rows of 8192 points on the four-pass core: 512 threads x 16 points, one row per CTA at a time, rows claimed
// dynamically, the next claimed row pulled into L2 while this one is transformed (as fast3_kernel).
template <typename T, bool BWD>
__global__ void __launch_bounds__(512, 1)
fast4_8192_kernel(const cx<T> *__restrict__ in, cx<T> *__restrict__ out, uint64_t nrows, int64_t rs_in, int64_t rs_out,
                  const cx<T> *__restrict__ tw1, const cx<T> *__restrict__ tw2, T fct, unsigned int *__restrict__ sched) {
  using F = Fft4_8192<T>;
  constexpr int TT = F::TT, M1 = F::M1;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cx<T> *buf = reinterpret_cast<cx<T> *>(smem_raw);
  unsigned int *s_row = reinterpret_cast<unsigned int *>(buf + F::BUFN);
  cx<T> *s_tw2 = reinterpret_cast<cx<T> *>(s_row + 4);   // [16][32]
  const int t = threadIdx.x;
  if (t == 0) { s_row[0] = atomicAdd(&sched[0], 1u); s_row[1] = atomicAdd(&sched[0], 1u); }
  for (int idx = t; idx < 16 * 32; idx += TT) s_tw2[idx] = tw2[idx];
  cx<T> twA[3], twB[3];
#pragma unroll
  for (int a = 1; a < 4; ++a) { twA[a - 1] = tw1[(4 * a) * M1 + t]; twB[a - 1] = tw1[a * M1 + t]; }
  __syncthreads();
  for (unsigned it = 0;; ++it) {
    const uint64_t row = s_row[it & 1];
    if (row >= nrows) break;
    if (t == 0) {
      const uint64_t nxt = s_row[(it + 1) & 1];
      if (nxt < nrows) prefetch_l2_bulk(in + (int64_t)nxt * rs_in, (uint32_t)(F::N * sizeof(cx<T>)));
    }
    cx<T> x[16];
    const cx<T> *src = in + (int64_t)row * rs_in;
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      x[q] = src[t + TT * q];
      if (BWD) x[q].y = -x[q].y;
    }
    F::run(x, buf, tw1, s_tw2, twA, twB, t);   // its first barrier comes after every thread has read s_row[it & 1]
    if (t == 0) s_row[it & 1] = atomicAdd(&sched[0], 1u);
    cx<T> *dst = out + (int64_t)row * rs_out;
    const T fy = BWD ? -fct : fct;
#pragma unroll
    for (int q = 0; q < 16; ++q) dst[t + TT * q] = mk<T>(x[q].x * fct, x[q].y * fy);
    __syncthreads();   // pass-4 reads done before the next row's pass-1 writes
  }
  __syncthreads();
  if (t == 0) {
    __threadfence();
    const unsigned done = atomicAdd(&sched[1], 1u);
    if (done == gridDim.x - 1) { sched[0] = 0u; sched[1] = 0u; __threadfence(); }
  }
}

template <typename T, int R1, int R2, int R3, int E, bool FOUR> struct FastBlueCore { using type = Fft3<T, R1, R2, R3, E>; };
template <typename T, int R1, int R2, int R3, int E> struct FastBlueCore<T, R1, R2, R3, E, true> { using type = Fft4_8192<T>; };

enum { BL_C2C = 0, BL_R2C_PAIR = 1, BL_C2R_PAIR = 2 };
constexpr int kBlueMaxDef = 16;  // largest supported deficiency d

// BKS: the chirp table b_k (L entries, used before the first and after the second transform of every unit) is
// copied to shared memory once per CTA instead of being streamed from L2 twice per unit.
// BFE: FFT(b)/M is multiplied in inside the first transform's last pass, half of it requested ahead of the butterfly.
// FOUR: the 8192-point work array on Fft4_8192 (instantiate with R1,R2,R3 = 16,16,32 and E = 16: 512 threads).
// TMB (four-pass core with BFE): FFT(b)/M lives in TENSOR MEMORY — every thread keeps the 16 entries it multiplies by
// in its own columns for the life of the CTA (131 KB per unit no longer streamed from L2).
template <typename T, int R1, int R2, int R3, int E, int KIND, bool BWD, bool BKS = false, bool BFE = false, bool FOUR = false,
          bool TMB = false>
__global__ void __launch_bounds__((R1 * R2 * R3) / E, 1)
fastblue_kernel(const void *__restrict__ in_v, void *__restrict__ out_v, uint64_t nrows, int64_t rs_in, int64_t rs_out,
                uint32_t L, uint32_t d, const cx<T> *__restrict__ tw1, const cx<T> *__restrict__ tw2,
                const cx<T> *__restrict__ bk, const cx<T> *__restrict__ bf, const cx<T> *__restrict__ corr, T fct,
                unsigned int *__restrict__ sched) {
  using F = typename FastBlueCore<T, R1, R2, R3, E, FOUR>::type;
  static_assert(F::TT * E == R1 * R2 * R3, "core / launch shape mismatch");
  constexpr int TT = F::TT, M1 = F::M1;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cx<T> *buf = reinterpret_cast<cx<T> *>(smem_raw);
  unsigned int *s_row = reinterpret_cast<unsigned int *>(buf + F::BUFN);
  cx<T> *s_tw2 = reinterpret_cast<cx<T> *>(s_row + 4);
  cx<T> *s_tail = s_tw2 + R2 * R3;  // a[n], n >= L-d
  cx<T> *s_bk = s_tail + kBlueMaxDef;
  const int t = threadIdx.x;
  static_assert(!TMB || (FOUR && BFE && BKS && E == 16), "tensor-memory multipliers: four-pass core with the early multiply");
  constexpr int TMB_WORDS = 16 * (int)(sizeof(cx<T>) / 4), TMB_COLS = TMB_WORDS * (TT / 128);   // 64 x 4 = 256 columns (fp64)
  uint32_t tmb_base = 0, tmb_addr = 0;
  (void)tmb_base; (void)tmb_addr;
  if constexpr (TMB) {
    static_assert(TMB_COLS <= 512 && (TMB_COLS & (TMB_COLS - 1)) == 0, "tensor-memory columns");
    tmb_base = cw_tmem_acquire<TMB_COLS>(reinterpret_cast<uint32_t *>(s_bk + L), t);
    tmb_addr = tmb_base + ((uint32_t)(((t >> 5) & 3) * 32) << 16) + (uint32_t)(t >> 7) * TMB_WORDS;
    cx<T> w[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) w[q] = __ldg(bf + t + TT * q);
    const uint32_t *ww = reinterpret_cast<const uint32_t *>(&w[0]);
#pragma unroll
    for (int c = 0; c < TMB_WORDS; c += 8) cw_tmem_st8(tmb_addr + (uint32_t)c, ww + c);
    cw_tmem_wait_st();
  }
  if (BKS) for (uint32_t idx = t; idx < L; idx += TT) s_bk[idx] = bk[idx];
  const cx<T> *bkp = BKS ? s_bk : bk;
  const uint64_t nunits = KIND == BL_C2C ? nrows : (nrows + 1) / 2;
  if (t == 0) { s_row[0] = atomicAdd(&sched[0], 1u); s_row[1] = atomicAdd(&sched[0], 1u); }
  for (int idx = t; idx < R2 * R3; idx += TT) s_tw2[idx] = tw2[idx];
  cx<T> twA[3], twB[3];
  if (F::TW1_REGS) {
#pragma unroll
    for (int a = 1; a < 4; ++a) { twA[a - 1] = tw1[(4 * a) * M1 + t]; twB[a - 1] = tw1[a * M1 + t]; }
  }
  __syncthreads();
  const uint32_t h = (L - 1) / 2;
  for (unsigned it = 0;; ++it) {
    const uint64_t unit = s_row[it & 1];
    if (unit >= nunits) break;
    const uint64_t ra = KIND == BL_C2C ? unit : 2 * unit, rb = ra + 1;
    const bool has_b = KIND != BL_C2C && rb < nrows;
    cx<T> x[E];
    // ---- load the logical complex line z[n], multiply by conj(b_n), zero-pad to M
#pragma unroll
    for (int q = 0; q < E; ++q) {
      const uint32_t n = (uint32_t)t + (uint32_t)TT * q;
      cx<T> z = mk<T>((T)0, (T)0);
      if (n < L) {
        if (KIND == BL_C2C) {
          z = reinterpret_cast<const cx<T> *>(in_v)[(int64_t)ra * rs_in + n];
          if (BWD) z.y = -z.y;
        } else if (KIND == BL_R2C_PAIR) {
          const T *p = reinterpret_cast<const T *>(in_v);
          z.x = p[(int64_t)ra * rs_in + n];
          z.y = has_b ? p[(int64_t)rb * rs_in + n] : (T)0;
        } else {
          // Hermitian extension of both half spectra, Z = Xa + i*Xb; the backward transform runs as
          // conj(FFT(conj Z)), so conj(Z) is what enters the forward Bluestein
          const cx<T> *p = reinterpret_cast<const cx<T> *>(in_v);
          const uint32_t kk = n <= h ? n : L - n;
          cx<T> xa = p[(int64_t)ra * rs_in + kk];
          cx<T> xb = has_b ? p[(int64_t)rb * rs_in + kk] : mk<T>((T)0, (T)0);
          if (BWD) { xa.y = -xa.y; xb.y = -xb.y; }   // c2r with forward=true conjugates its input
          if (kk == 0) { xa.y = (T)0; xb.y = (T)0; }
          if (n > h) { xa.y = -xa.y; xb.y = -xb.y; }
          z = mk<T>(xa.x - xb.y, -(xa.y + xb.x));     // conj(xa + i*xb)
        }
        z = cmul(z, cconj(BKS ? bkp[n] : __ldg(bk + n)));
        if (n + d >= L) s_tail[n - (L - d)] = z;
      }
      x[q] = z;
    }
    if (t == 0) {  // pull the next unit's row(s) into L2 (row by row: never past the end of the array)
      const uint64_t nxt = s_row[(it + 1) & 1];
      if (nxt < nunits) {
        const size_t esz = KIND == BL_R2C_PAIR ? sizeof(T) : sizeof(cx<T>);
        const uint32_t bytes = (uint32_t)((KIND == BL_C2R_PAIR ? (h + 1) : L) * esz);
        const uint64_t r0 = KIND == BL_C2C ? nxt : 2 * nxt;
        for (uint64_t rr = r0; rr < r0 + (KIND == BL_C2C ? 1 : 2) && rr < nrows; ++rr) {
          const char *p = reinterpret_cast<const char *>(in_v) + (int64_t)rr * rs_in * (int64_t)esz;
          const uintptr_t lo = (reinterpret_cast<uintptr_t>(p) + 15) & ~(uintptr_t)15;
          const uintptr_t hi = (reinterpret_cast<uintptr_t>(p) + bytes) & ~(uintptr_t)15;
          if (hi > lo) prefetch_l2_bulk(reinterpret_cast<const void *>(lo), (uint32_t)(hi - lo));
        }
      }
    }
    if constexpr (BFE) {
      if constexpr (TMB) F::template run<true, true>(x, buf, tw1, s_tw2, twA, twB, t, bf, tmb_addr);
      else F::template run<true>(x, buf, tw1, s_tw2, twA, twB, t, bf);
      if (t == 0) s_row[it & 1] = atomicAdd(&sched[0], 1u);
    } else {
      F::run(x, buf, tw1, s_tw2, twA, twB, t);
      if (t == 0) s_row[it & 1] = atomicAdd(&sched[0], 1u);  // every thread read s_row[it&1] before the core's barriers
      // ---- multiply by FFT(b)/M, conjugate for the inverse transform
#pragma unroll
      for (int q = 0; q < E; ++q) {
        const cx<T> v = cmul(x[q], __ldg(bf + t + TT * q));
        x[q] = mk<T>(v.x, -v.y);
      }
    }
    __syncthreads();  // pass-3 reads of the first transform are done
    F::run(x, buf, tw1, s_tw2, twA, twB, t);
    // ---- conv[k] = conj(x), add the aliased lags back, multiply by conj(b_k)
#pragma unroll
    for (int q = 0; q < E; ++q) {
      const uint32_t k = (uint32_t)t + (uint32_t)TT * q;
      cx<T> v = mk<T>(x[q].x, -x[q].y);
      if (k < d) {
        for (uint32_t j = k; j < d; ++j) v = cadd(v, cmul(s_tail[j], __ldg(corr + k * kBlueMaxDef + j)));
      }
      x[q] = k < L ? cmul(v, cconj(BKS ? bkp[k] : __ldg(bk + k))) : mk<T>((T)0, (T)0);
    }
    // ---- store
    if (KIND == BL_C2C) {
      cx<T> *dst = reinterpret_cast<cx<T> *>(out_v) + (int64_t)ra * rs_out;
#pragma unroll
      for (int q = 0; q < E; ++q) {
        const uint32_t k = (uint32_t)t + (uint32_t)TT * q;
        if (k < L) { cx<T> v = x[q]; v.x *= fct; v.y *= BWD ? -fct : fct; dst[k] = v; }
      }
      __syncthreads();
    } else if (KIND == BL_C2R_PAIR) {
      T *pa = reinterpret_cast<T *>(out_v) + (int64_t)ra * rs_out, *pb = reinterpret_cast<T *>(out_v) + (int64_t)rb * rs_out;
#pragma unroll
      for (int q = 0; q < E; ++q) {
        const uint32_t n = (uint32_t)t + (uint32_t)TT * q;
        if (n < L) {                       // z = conj(result): a = Re, b = Im
          pa[n] = x[q].x * fct;
          if (has_b) pb[n] = -x[q].y * fct;
        }
      }
      __syncthreads();
    } else {  // r2c pair: A[k] = (Z[k] + conj Z[L-k])/2, B[k] = (Z[k] - conj Z[L-k])/(2i)
      __syncthreads();
#pragma unroll
      for (int q = 0; q < E; ++q) {
        const uint32_t k = (uint32_t)t + (uint32_t)TT * q;
        if (k < L) buf[k] = x[q];
      }
      __syncthreads();
      cx<T> *pa = reinterpret_cast<cx<T> *>(out_v) + (int64_t)ra * rs_out, *pb = reinterpret_cast<cx<T> *>(out_v) + (int64_t)rb * rs_out;
      const T hf = (T)0.5 * fct;
#pragma unroll
      for (int q = 0; q < E; ++q) {
        const uint32_t k = (uint32_t)t + (uint32_t)TT * q;
        if (k <= h) {
          const cx<T> z1 = buf[k], z2 = cconj(buf[k == 0 ? 0 : L - k]);
          cx<T> A = mk<T>((z1.x + z2.x) * hf, (z1.y + z2.y) * hf);
          cx<T> B = mk<T>((z1.y - z2.y) * hf, -(z1.x - z2.x) * hf);
          if (BWD) { A.y = -A.y; B.y = -B.y; }   // r2c with forward=false returns the conjugate spectrum
          pa[k] = A;
          if (has_b) pb[k] = B;
        }
      }
      __syncthreads();
    }
  }
  __syncthreads();
  if (t == 0) {
    __threadfence();
    const unsigned done = atomicAdd(&sched[1], 1u);
    if (done == gridDim.x - 1) { sched[0] = 0u; sched[1] = 0u; __threadfence(); }
  }
  if constexpr (TMB) cw_tmem_release<TMB_COLS>(tmb_base, t);
}

}  // namespace impulse
