// Column register kernels: two-pass FFTs over adjacent STRIDED lines (colfast2_kernel), the software-pipelined
// complex128 variant (colpipe2_kernel) and the fused FFT * multiply * inverse-FFT middle pass of an axis convolution
// (colconv2_kernel).  A header of its own so that tests/emu/emu_col.cpp can compile the SAME kernel bodies for the
// host (thread-level emulation, test infrastructure only); the product runs them on the device only.
#pragma once
#include "fast3_device.cuh"

namespace impulse {

// W_N^m from the two-level four-step table of the job (m < N)
template <typename T>
__device__ __forceinline__ cx<T> four_step_w(const LineJob &J, uint32_t m) {
  return cmul(__ldg(reinterpret_cast<const cx<T> *>(J.tw4_hi) + (m >> J.tw4_shift)),
              __ldg(reinterpret_cast<const cx<T> *>(J.tw4_lo) + (m & ((1u << J.tw4_shift) - 1))));
}

// =================================================================================================
// Column kernel: two-pass register FFT over LPC adjacent STRIDED lines (element n of line l at
// base + n*es + l).  Thread u = line + LPC*i, so for every register index the lanes of a warp read /
// write LPC*16 B contiguous runs (128 B for LPC = 8 complex128 / 16 complex64).  The exchange buffer is
// laid out [k1][i][line] (line fastest): conflict-free both ways without padding.  Serves the strided
// axes of N-D transforms and both launches of the four-step split (optional W_N^(k*n2) store twiddle).
// One group of LPC lines per CTA: the hardware block scheduler balances the SMs.
// =================================================================================================
// Column kernels: CTA -> (group of adjacent lines, second and third batch index).  Launched as a 3-D grid
// whenever the two outer extents fit (no divisions at all); a 1-D grid decodes with 32-bit divisions.
struct ColGroup { uint32_t g0, i1, i2; };
__device__ __forceinline__ ColGroup col_group(const LineJob &J, uint32_t g0n) {
  ColGroup g;
  if (gridDim.y > 1 || gridDim.z > 1 || (J.bdim[1] == 1 && J.bdim[2] == 1)) {
    g.g0 = blockIdx.x; g.i1 = blockIdx.y; g.i2 = blockIdx.z;
  } else {
    const uint32_t b = blockIdx.x, r = b / g0n, d1 = (uint32_t)J.bdim[1];
    g.g0 = b - r * g0n; g.i2 = r / d1; g.i1 = r - g.i2 * d1;
  }
  return g;
}
// Short-lived CTAs cannot double-buffer; instead one lane per 128-byte run asks L2 for the input of the
// group `rows_ahead` rows further along the second batch index (about one resident wave of CTAs ahead).
__device__ __forceinline__ void col_prefetch(const LineJob &J, const ColGroup &cg, uint32_t lpc, int i, int r1, int r2) {
  const uint32_t g0n = (uint32_t)((J.bdim[0] + lpc - 1) / lpc);
  const uint32_t rows_ahead = (1776u + g0n - 1) / g0n;
  uint32_t p1 = cg.i1 + rows_ahead, p2 = cg.i2;
  if (p1 >= (uint32_t)J.bdim[1]) { p1 -= (uint32_t)J.bdim[1]; ++p2; }
  if (p1 >= (uint32_t)J.bdim[1] || p2 >= (uint32_t)J.bdim[2]) return;
  const char *pin = reinterpret_cast<const char *>(J.in) +
                    ((int64_t)(cg.g0 * lpc) + (int64_t)p1 * J.bs_in[1] + (int64_t)p2 * J.bs_in[2]) * (J.dtype == 1 ? 16 : 8);
  const int64_t step = J.es_in * (J.dtype == 1 ? 16 : 8);
#if defined(__CUDA_ARCH__)
  for (int j = 0; j < r1; ++j) asm volatile("prefetch.global.L2 [%0];" ::"l"(pin + (int64_t)(i + r2 * j) * step));
#else
  (void)pin; (void)step; (void)i; (void)r1; (void)r2;
#endif
}

struct ColYes { static constexpr bool value = true; };
struct ColNo { static constexpr bool value = false; };
// PLAIN: the job has no segmented input and no fused multiply (the launcher checks) — those code paths, their uniform
// tests inside the unrolled loops and the constant-bank loads that feed them are compiled out; the four-step twiddle is
// selected once per CTA (two copies of the second pass) and global addresses step by a 64-bit stride formed once.
// Measured on the float32 column passes of config 5 (profiles/r02_colfast2_f32_inst_mix.txt): 588 executed instructions
// per thread for 8 elements, 30 % of them floating point — constant loads 13 %, branch / reconvergence 10 %.
template <typename T, int R1, int R2, int LPC, bool BWD, bool PF, bool INROWS, bool PLAIN = false>
__global__ void __launch_bounds__(LPC * R2)
colfast2_kernel(const __grid_constant__ LineJob J) {
  constexpr int N = R1 * R2, NB2 = R1 / R2;
  static_assert(R1 % R2 == 0, "column two-pass shape");
  static_assert(!PLAIN || !INROWS, "the plain variant serves strided lines");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cx<T> *S = reinterpret_cast<cx<T> *>(smem_raw);
  const int u = threadIdx.x, line = u % LPC, i = u / LPC;
  // group -> batch indices (bdim[0] is the adjacent-lines dimension; its last group may be ragged)
  const ColGroup cg = col_group(J, (uint32_t)((J.bdim[0] + LPC - 1) / LPC));
  const uint32_t i0 = cg.g0 * LPC, i1 = cg.i1, i2 = cg.i2;
  const bool valid = i0 + line < (uint32_t)J.bdim[0];
  const int64_t off_in = (INROWS ? (int64_t)(i0 + line) * J.bs_in[0] : (int64_t)(i0 + line)) + (int64_t)i1 * J.bs_in[1] +
                         (int64_t)i2 * J.bs_in[2];
  const int64_t off_out = (int64_t)(i0 + line) + (int64_t)i1 * J.bs_out[1] + (int64_t)i2 * J.bs_out[2];
  const uint32_t twi = J.tw4_dim == 0 ? i0 + line : J.tw4_dim == 1 ? i1 : J.tw4_dim == 2 ? i2 : 0u;
  const cx<T> *in = reinterpret_cast<const cx<T> *>(J.in) + off_in;
  cx<T> *out = reinterpret_cast<cx<T> *>(J.out) + off_out;
  const cx<T> *tw = reinterpret_cast<const cx<T> *>(J.tw);   // W_N^m
  if (PF && !INROWS && line == 0 && (PLAIN || !J.seg_len)) col_prefetch(J, cg, LPC, i, R1, R2);
  cx<T> x[R1];
  if (INROWS) {
    // every line is a CONTIGUOUS row (second launch of the split on contiguous data: rows in, transposed out).
    // The group's rows are copied to shared memory with consecutive threads on consecutive elements and read
    // back in the compute layout; rows are padded by one element so that both sides are conflict-free.
    cx<T> *stage = S;   // shares the exchange buffer: one more barrier, half the shared memory
    const cx<T> *g0 = reinterpret_cast<const cx<T> *>(J.in) + (int64_t)i0 * J.bs_in[0] + (int64_t)i1 * J.bs_in[1] +
                      (int64_t)i2 * J.bs_in[2];
    const uint32_t nl = min((uint32_t)LPC, (uint32_t)J.bdim[0] - i0);
#pragma unroll
    for (int q = 0; q < R1; ++q) {
      const uint32_t f = (uint32_t)u + (uint32_t)(LPC * R2) * q, l = f / N, n = f % N;
      if (l < nl) stage[l * (N + 1) + n] = g0[(int64_t)l * J.bs_in[0] + n];
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < R1; ++j) {
      x[j] = valid ? stage[line * (N + 1) + i + R2 * j] : mk<T>((T)0, (T)0);
      if (BWD) x[j].y = -x[j].y;
    }
    __syncthreads();
  } else if (PLAIN) {
    const cx<T> *p = in + (int64_t)i * J.es_in;
    const int64_t step = (int64_t)R2 * J.es_in;
#pragma unroll
    for (int j = 0; j < R1; ++j) {
      x[j] = valid ? *p : mk<T>((T)0, (T)0);
      p += step;
      if (BWD) x[j].y = -x[j].y;
    }
  } else
#pragma unroll
  for (int j = 0; j < R1; ++j) {
    const uint32_t n = (uint32_t)(i + R2 * j);
    if (J.seg_len) {  // the line is spread over several allocations (peer slabs): segment n / seg_len
      const uint32_t sg = n / J.seg_len, w = n - sg * J.seg_len;
      x[j] = valid ? (reinterpret_cast<const cx<T> *>(J.seg_base[sg]) + off_in)[(int64_t)w * J.es_in] : mk<T>((T)0, (T)0);
    } else {
      x[j] = valid ? in[(int64_t)n * J.es_in] : mk<T>((T)0, (T)0);
    }
    if (BWD) x[j].y = -x[j].y;
  }
  RegFFT<T, R1>::run(x);
  {
    const cx<T> *pt = tw;
#pragma unroll
    for (int k = 1; k < R1; ++k) { pt += i; x[k] = cmul(x[k], __ldg(pt)); }
  }
#pragma unroll
  for (int k = 0; k < R1; ++k) S[(k * R2 + i) * LPC + line] = x[k];
  __syncthreads();
  const T f = (T)J.fct;
  if constexpr (PLAIN) {
    const int64_t ostep = (int64_t)R1 * J.es_out;
    // second pass, with (TW) or without the four-step store twiddle W_N^(k*n2): chosen once per CTA
    auto second = [&](auto TW) {
      constexpr bool tw4 = decltype(TW)::value;
      cx<T> wstep = mk<T>((T)1, (T)0);
      if constexpr (tw4) wstep = four_step_w<T>(J, (uint32_t)R1 * twi);
#pragma unroll
      for (int m = 0; m < NB2; ++m) {
        const int k1 = i + R2 * m;
        cx<T> y[R2];
#pragma unroll
        for (int j = 0; j < R2; ++j) y[j] = S[(k1 * R2 + j) * LPC + line];
        RegFFT<T, R2>::run(y);
        cx<T> w = mk<T>((T)1, (T)0);
        if constexpr (tw4) w = four_step_w<T>(J, (uint32_t)k1 * twi);
        cx<T> *po = out + (int64_t)k1 * J.es_out;
#pragma unroll
        for (int k2 = 0; k2 < R2; ++k2) {
          cx<T> v = y[k2];
          if constexpr (tw4) { v = cmul(v, w); w = cmul(w, wstep); }
          v.x *= f;
          v.y *= BWD ? -f : f;
          if (valid) *po = v;
          po += ostep;
        }
      }
    };
    if (J.tw4_n) second(ColYes{}); else second(ColNo{});
    return;
  }
  const uint64_t umul_off = J.umul_mod ? (uint64_t)off_out % J.umul_mod : 0;
  // first launch of the split: output k is multiplied by W_N^(k*n2) (n2 = twi).  k = k1 + R1*k2 walks in
  // steps of R1, so one table lookup per k1 plus the step W_N^(R1*n2) replace a lookup per element.
  cx<T> wstep = mk<T>((T)1, (T)0);
  if (J.tw4_n) wstep = four_step_w<T>(J, (uint32_t)R1 * twi);
#pragma unroll
  for (int m = 0; m < NB2; ++m) {
    const int k1 = i + R2 * m;
    cx<T> y[R2];
#pragma unroll
    for (int j = 0; j < R2; ++j) y[j] = S[(k1 * R2 + j) * LPC + line];
    RegFFT<T, R2>::run(y);
    cx<T> w = mk<T>((T)1, (T)0);
    if (J.tw4_n) w = four_step_w<T>(J, (uint32_t)k1 * twi);
#pragma unroll
    for (int k2 = 0; k2 < R2; ++k2) {
      const int k = k1 + R1 * k2;
      cx<T> v = y[k2];
      if (J.tw4_n) {  // the conjugation below turns the twiddle into its inverse for the backward transform
        v = cmul(v, w);
        w = cmul(w, wstep);
      }
      v.x *= f;
      v.y *= BWD ? -f : f;
      if (J.umul_mod && valid) {  // fused pointwise multiply (e.g. the filter spectrum of an FFT convolution)
        uint64_t o = umul_off + (uint64_t)((int64_t)k * J.es_out);
        if (o >= J.umul_mod) o %= J.umul_mod;   // one line usually spans at most one period: rarely taken
        v = cmul(v, __ldg(reinterpret_cast<const cx<T> *>(J.umul) + o));
      }
      if (valid) out[(int64_t)k * J.es_out] = v;
    }
  }
}

// Software-pipelined variant for complex128 (plain strided lines, optional four-step twiddle): a CTA walks G
// consecutive groups along the adjacent-lines dimension and requests the NEXT group's elements into a second
// register set before it transforms the current one, so the global-load latency that dominates the 64-thread
// complex128 CTAs (ncu: long-scoreboard 9-16 per issue) overlaps the arithmetic instead of preceding it.
template <typename T, int R1, int R2, int LPC, bool BWD>
__global__ void __launch_bounds__(LPC * R2)
colpipe2_kernel(const __grid_constant__ LineJob J, const uint32_t G) {
  constexpr int N = R1 * R2, NB2 = R1 / R2;
  static_assert(R1 % R2 == 0, "column two-pass shape");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cx<T> *S = reinterpret_cast<cx<T> *>(smem_raw);
  const int u = threadIdx.x, line = u % LPC, i = u / LPC;
  const uint32_t g0n = (uint32_t)((J.bdim[0] + LPC - 1) / LPC), chunks = (g0n + G - 1) / G;
  const ColGroup cg = col_group(J, chunks);
  const uint32_t i1 = cg.i1, i2 = cg.i2, gA = cg.g0 * G;
  const int64_t base_in = (int64_t)i1 * J.bs_in[1] + (int64_t)i2 * J.bs_in[2];
  const int64_t base_out = (int64_t)i1 * J.bs_out[1] + (int64_t)i2 * J.bs_out[2];
  const cx<T> *tw = reinterpret_cast<const cx<T> *>(J.tw);
  const T f = (T)J.fct;
  auto load = [&](uint32_t grp, cx<T> (&x)[R1]) {
    const uint32_t l0 = grp * LPC + line;
    const bool ok = grp < g0n && l0 < (uint32_t)J.bdim[0];
    const cx<T> *in = reinterpret_cast<const cx<T> *>(J.in) + base_in + l0;
#pragma unroll
    for (int j = 0; j < R1; ++j) x[j] = ok ? in[(int64_t)(i + R2 * j) * J.es_in] : mk<T>((T)0, (T)0);
  };
  cx<T> x[R1], xn[R1];
  load(gA, x);
  for (uint32_t g = 0; g < G; ++g) {
    const uint32_t grp = gA + g;
    if (grp >= g0n) break;
    if (g + 1 < G) load(grp + 1, xn);
    const uint32_t l0 = grp * LPC + line;
    const bool valid = l0 < (uint32_t)J.bdim[0];
    const uint32_t twi = J.tw4_dim == 0 ? l0 : J.tw4_dim == 1 ? i1 : J.tw4_dim == 2 ? i2 : 0u;
    cx<T> *out = reinterpret_cast<cx<T> *>(J.out) + base_out + l0;
    if (BWD) {
#pragma unroll
      for (int j = 0; j < R1; ++j) x[j].y = -x[j].y;
    }
    RegFFT<T, R1>::run(x);
#pragma unroll
    for (int k = 1; k < R1; ++k) x[k] = cmul(x[k], __ldg(tw + i * k));
#pragma unroll
    for (int k = 0; k < R1; ++k) S[(k * R2 + i) * LPC + line] = x[k];
    __syncthreads();
    cx<T> wstep = mk<T>((T)1, (T)0);
    if (J.tw4_n) wstep = four_step_w<T>(J, (uint32_t)R1 * twi);
#pragma unroll
    for (int m = 0; m < NB2; ++m) {
      const int k1 = i + R2 * m;
      cx<T> y[R2];
#pragma unroll
      for (int j = 0; j < R2; ++j) y[j] = S[(k1 * R2 + j) * LPC + line];
      RegFFT<T, R2>::run(y);
      cx<T> w = mk<T>((T)1, (T)0);
      if (J.tw4_n) w = four_step_w<T>(J, (uint32_t)k1 * twi);
#pragma unroll
      for (int k2 = 0; k2 < R2; ++k2) {
        cx<T> v = y[k2];
        if (J.tw4_n) { v = cmul(v, w); w = cmul(w, wstep); }
        v.x *= f;
        v.y *= BWD ? -f : f;
        if (valid) out[(int64_t)(k1 + R1 * k2) * J.es_out] = v;
      }
    }
    __syncthreads();   // the exchange buffer is reused by the next group
#pragma unroll
    for (int j = 0; j < R1; ++j) x[j] = xn[j];
  }
}

// =================================================================================================
// Convolution along a strided axis, middle pass.  A length-N = N1*N2 convolution is
//     IFFT_N( FFT_N(x) .* m ).
// With the forward transform split as n = n1*N2 + n2 -> k = k1 + N1*k2 and the inverse split the
// other way round (n = n1'*N1 + n2' -> k' = k1' + N2*k2'), bin k of the spectrum is element
// (n1' = k2, n2' = k1) of the inverse's input: the N2 outputs of the forward transform's second pass
// for one k1 ARE the inputs of one line of the inverse's first pass.  So per line (k1 fixed):
//     N2-point forward FFT -> * m[k1 + N1*k2] -> N2-point inverse FFT -> * conj(W_N^(k1'*k1))
// never leaves the SM, and the spectrum is neither written nor re-read: three passes over the data
// instead of five (FFT pass A, this kernel, inverse pass B).  Layout and thread mapping as
// colfast2_kernel; the multiplier is indexed by the element offset this kernel writes to.
// =================================================================================================
template <typename T, int R1, int R2, int LPC, bool PF>
__global__ void __launch_bounds__(LPC * R2)
colconv2_kernel(const __grid_constant__ LineJob J) {
  constexpr int N = R1 * R2, NB2 = R1 / R2;
  static_assert(R1 % R2 == 0, "column two-pass shape");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cx<T> *S0 = reinterpret_cast<cx<T> *>(smem_raw), *S1 = S0 + N * LPC;
  const int u = threadIdx.x, line = u % LPC, i = u / LPC;
  const ColGroup cg = col_group(J, (uint32_t)((J.bdim[0] + LPC - 1) / LPC));
  const uint32_t i0 = cg.g0 * LPC, i1 = cg.i1, i2 = cg.i2;
  const bool valid = i0 + line < (uint32_t)J.bdim[0];
  const int64_t off_in = (int64_t)(i0 + line) + (int64_t)i1 * J.bs_in[1] + (int64_t)i2 * J.bs_in[2];
  const int64_t off_out = (int64_t)(i0 + line) + (int64_t)i1 * J.bs_out[1] + (int64_t)i2 * J.bs_out[2];
  const uint32_t twi = J.tw4_dim == 0 ? i0 + line : J.tw4_dim == 1 ? i1 : i2;
  const cx<T> *in = reinterpret_cast<const cx<T> *>(J.in) + off_in;
  cx<T> *out = reinterpret_cast<cx<T> *>(J.out) + off_out;
  const cx<T> *tw = reinterpret_cast<const cx<T> *>(J.tw);
  const cx<T> *um = reinterpret_cast<const cx<T> *>(J.umul);
  const uint64_t umul_off = (uint64_t)off_out % J.umul_mod;
  if (PF && line == 0) col_prefetch(J, cg, LPC, i, R1, R2);
  cx<T> x[R1];
#pragma unroll
  for (int j = 0; j < R1; ++j) x[j] = valid ? in[(int64_t)(i + R2 * j) * J.es_in] : mk<T>((T)0, (T)0);
  // the multipliers this thread will need after the forward transform: requested now, so that their
  // (L2) latency overlaps the data loads instead of sitting between the two transforms
  cx<T> mreg[R1];
#pragma unroll
  for (int m = 0; m < NB2; ++m)
#pragma unroll
    for (int k2 = 0; k2 < R2; ++k2) {
      const int k = i + R2 * m + R1 * k2;
      uint64_t o = umul_off + (uint64_t)((int64_t)k * J.es_out);
      if (o >= J.umul_mod) o %= J.umul_mod;
      mreg[m * R2 + k2] = valid ? __ldg(um + o) : mk<T>((T)0, (T)0);
    }
  // ---- forward N-point FFT
  RegFFT<T, R1>::run(x);
#pragma unroll
  for (int k = 1; k < R1; ++k) x[k] = cmul(x[k], __ldg(tw + i * k));
#pragma unroll
  for (int k = 0; k < R1; ++k) S0[(k * R2 + i) * LPC + line] = x[k];
  __syncthreads();
#pragma unroll
  for (int m = 0; m < NB2; ++m) {
    const int k1 = i + R2 * m;
    cx<T> y[R2];
#pragma unroll
    for (int j = 0; j < R2; ++j) y[j] = S0[(k1 * R2 + j) * LPC + line];
    RegFFT<T, R2>::run(y);
#pragma unroll
    for (int k2 = 0; k2 < R2; ++k2) {
      const int k = k1 + R1 * k2;
      cx<T> v = cmul(y[k2], mreg[m * R2 + k2]);
      v.y = -v.y;  // inverse = conj(FFT(conj(.)))
      S1[k * LPC + line] = v;
    }
  }
  __syncthreads();
  // ---- inverse N-point FFT of the products
#pragma unroll
  for (int j = 0; j < R1; ++j) x[j] = S1[(i + R2 * j) * LPC + line];
  RegFFT<T, R1>::run(x);
#pragma unroll
  for (int k = 1; k < R1; ++k) x[k] = cmul(x[k], __ldg(tw + i * k));
#pragma unroll
  for (int k = 0; k < R1; ++k) S0[(k * R2 + i) * LPC + line] = x[k];   // S0 is free: everyone passed the second barrier
  __syncthreads();
  const T f = (T)J.fct;
  const cx<T> wstep = four_step_w<T>(J, (uint32_t)R1 * twi);   // four-step twiddle of the inverse's first pass,
#pragma unroll                                                 // stepped along k = k1 + R1*k2 (conjugated below)
  for (int m = 0; m < NB2; ++m) {
    const int k1 = i + R2 * m;
    cx<T> y[R2];
#pragma unroll
    for (int j = 0; j < R2; ++j) y[j] = S0[(k1 * R2 + j) * LPC + line];
    RegFFT<T, R2>::run(y);
    cx<T> w = four_step_w<T>(J, (uint32_t)k1 * twi);
#pragma unroll
    for (int k2 = 0; k2 < R2; ++k2) {
      const int k = k1 + R1 * k2;
      cx<T> v = cmul(y[k2], w);
      w = cmul(w, wstep);
      v.x *= f;
      v.y *= -f;
      if (valid) out[(int64_t)k * J.es_out] = v;
    }
  }
}

}  // namespace impulse
