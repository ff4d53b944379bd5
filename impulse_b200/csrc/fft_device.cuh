// Generic shared-memory line-FFT engine: device-side phase interpreter.
//
// One CTA owns a tile of C = 1<<log_c lines.  Every phase is a pure function of
// (job, tile, thread id); the kernel (fft_kernels.cu) runs them with a block barrier
// in between.  Because each phase is a plain function of its thread id, the same code
// also compiles for the host (tests/emu) where a loop over thread ids replaces the CTA —
// that is how index logic is checked in the GPU-less build container.  The emulation
// is test infrastructure only; the product library contains no CPU transform path.
//
// Algorithms (what of the reference each piece replaces):
//   * radix passes      — in-place DIF (and its transpose, DIT) on CC(i,j,k) =
//                         s[i + ido*(j + ip*k)], twiddle W_n^(i*j*l1); replaces
//                         pass2..pass11/passg + pass_all (pocketfft.c:300-929).
//                         Natural order comes back through a host-built position table
//                         at STORE time instead of pocketfft's ping-pong buffers.
//   * real transforms   — even N: N/2-point complex FFT + Hermitian post/pre-twiddle
//                         (replaces radf*/radb*, pocketfft.c:1082-1766); odd N: N-point
//                         complex FFT of the zero-imag / Hermitian-extended line
//                         (what rfftblue_* does, pocketfft.c:2019-2058).
//   * Bluestein         — chirp, DIF FFT, multiply by pre-permuted bkf, DIT FFT, chirp:
//                         no reordering pass at all (replaces fftblue_fft, pocketfft.c:1945-2008).
//   * backward          — conj(FFT(conj(x))): conjugations folded into LOAD/STORE.
#pragma once
#include "fft_types.h"
#include "trig_tables.h"

namespace impulse {

#if defined(__CUDA_ARCH__)
#define IMP_LDG(p) __ldg(p)
#else
#define IMP_LDG(p) (*(p))
#endif

template <typename T> struct Vec2;
#if defined(__CUDACC__)
template <> struct Vec2<double> { using type = double2; };
template <> struct Vec2<float> { using type = float2; };
#else
struct alignas(16) emu_double2 { double x, y; };
struct alignas(8) emu_float2 { float x, y; };
template <> struct Vec2<double> { using type = emu_double2; };
template <> struct Vec2<float> { using type = emu_float2; };
#endif

template <typename T> using cx = typename Vec2<T>::type;

template <typename T> IMP_HD cx<T> mk(T a, T b) { cx<T> r; r.x = a; r.y = b; return r; }
template <typename C> IMP_HD C cadd(C a, C b) { C r; r.x = a.x + b.x; r.y = a.y + b.y; return r; }
template <typename C> IMP_HD C csub(C a, C b) { C r; r.x = a.x - b.x; r.y = a.y - b.y; return r; }
template <typename C> IMP_HD C cmul(C a, C b) { C r; r.x = a.x * b.x - a.y * b.y; r.y = a.x * b.y + a.y * b.x; return r; }
template <typename C> IMP_HD C cconj(C a) { a.y = -a.y; return a; }
template <typename C> IMP_HD C cconj_if(C a, bool c) { if (c) a.y = -a.y; return a; }
// multiply by -i / +i
template <typename C> IMP_HD C mul_mi(C a) { C r; r.x = a.y; r.y = -a.x; return r; }
template <typename C> IMP_HD C mul_pi(C a) { C r; r.x = -a.y; r.y = a.x; return r; }

// ---------------------------------------------------------------------------
// Forward DFT butterflies, in registers: x[k] <- sum_j x[j] exp(-2*pi*i*j*k/R)
// ---------------------------------------------------------------------------
template <typename T, int R> struct Bfly;

template <typename T> struct Bfly<T, 2> {
  static IMP_HD void run(cx<T> (&x)[2]) {
    cx<T> a = x[0], b = x[1];
    x[0] = cadd(a, b);
    x[1] = csub(a, b);
  }
};

template <typename T> struct Bfly<T, 4> {
  static IMP_HD void run(cx<T> (&x)[4]) {
    cx<T> t0 = cadd(x[0], x[2]), t1 = csub(x[0], x[2]);
    cx<T> t2 = cadd(x[1], x[3]), t3 = mul_mi(csub(x[1], x[3]));
    x[0] = cadd(t0, t2);
    x[2] = csub(t0, t2);
    x[1] = cadd(t1, t3);
    x[3] = csub(t1, t3);
  }
};

template <typename T> struct Bfly<T, 8> {
  static IMP_HD void run(cx<T> (&x)[8]) {
    const T h = (T)0.7071067811865475244008444;
    cx<T> a[4], b[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) { a[j] = cadd(x[j], x[j + 4]); b[j] = csub(x[j], x[j + 4]); }
    // b[j] *= W8^j : W8 = (1-i)/sqrt2, W8^2 = -i, W8^3 = (-1-i)/sqrt2
    { cx<T> t = b[1]; b[1].x = (t.x + t.y) * h; b[1].y = (t.y - t.x) * h; }
    b[2] = mul_mi(b[2]);
    { cx<T> t = b[3]; b[3].x = (t.y - t.x) * h; b[3].y = -(t.x + t.y) * h; }
    Bfly<T, 4>::run(a);
    Bfly<T, 4>::run(b);
#pragma unroll
    for (int k = 0; k < 4; ++k) { x[2 * k] = a[k]; x[2 * k + 1] = b[k]; }
  }
};

// any odd radix (prime or not): half-sum / half-difference form, O(R^2/2)
template <typename T, int R> struct Bfly {
  static_assert(R % 2 == 1, "generic butterfly is for odd radices");
  static IMP_HD void run(cx<T> (&x)[R]) {
    constexpr int H = (R - 1) / 2;
    cx<T> a[H], b[H];
#pragma unroll
    for (int m = 0; m < H; ++m) { a[m] = cadd(x[m + 1], x[R - 1 - m]); b[m] = csub(x[m + 1], x[R - 1 - m]); }
    cx<T> x0 = x[0];
    cx<T> s0 = x0;
#pragma unroll
    for (int m = 0; m < H; ++m) s0 = cadd(s0, a[m]);
    x[0] = s0;
#pragma unroll
    for (int k = 1; k <= H; ++k) {
      cx<T> A = x0, B = mk<T>((T)0, (T)0);
#pragma unroll
      for (int m = 1; m <= H; ++m) {
        const T c = (T)Trig<R>::c((k * m) % R);
        const T s = (T)Trig<R>::s((k * m) % R);
        A.x += c * a[m - 1].x; A.y += c * a[m - 1].y;
        B.x += s * b[m - 1].x; B.y += s * b[m - 1].y;
      }
      // y_k = A - i*B ; y_{R-k} = A + i*B
      x[k] = mk<T>(A.x + B.y, A.y - B.x);
      x[R - k] = mk<T>(A.x - B.y, A.y + B.x);
    }
  }
};

// ---------------------------------------------------------------------------
// helpers
// ---------------------------------------------------------------------------
template <typename T> IMP_HD uint32_t phys(const LineJob &J, uint32_t a) {
  return sizeof(T) == 8 ? swz(a, J.swz_mask) : swz16(a, J.swz_mask);
}

struct TileCtx {
  uint64_t tile;     // tile index
  uint32_t nlines;   // valid lines in this tile (<= C)
};

IMP_HD TileCtx tile_ctx(const LineJob &J, uint64_t tile) {
  TileCtx t;
  t.tile = tile;
  uint64_t first = tile << J.log_c;
  uint64_t left = J.n_lines - first;
  uint64_t c = 1ull << J.log_c;
  t.nlines = (uint32_t)(left < c ? left : c);
  return t;
}

// PROLOG: threads < C compute the global element offsets of their line.
IMP_HD void phase_prolog(const LineJob &J, const TileCtx &tc, uint32_t tid, int64_t *offs) {
  if (tid < tc.nlines) {
    uint64_t l = (tc.tile << J.log_c) + tid;
    uint64_t i0 = l % J.bdim[0];
    uint64_t r = l / J.bdim[0];
    uint64_t i1 = r % J.bdim[1];
    uint64_t i2 = r / J.bdim[1];
    offs[3 * tid] = (int64_t)i0 * J.bs_in[0] + (int64_t)i1 * J.bs_in[1] + (int64_t)i2 * J.bs_in[2];
    offs[3 * tid + 1] = (int64_t)i0 * J.bs_out[0] + (int64_t)i1 * J.bs_out[1] + (int64_t)i2 * J.bs_out[2];
    offs[3 * tid + 2] = (int64_t)(J.tw4_dim == 0 ? i0 : J.tw4_dim == 1 ? i1 : J.tw4_dim == 2 ? i2 : 0);
  }
}

// ---------------------------------------------------------------------------
// LOAD
// ---------------------------------------------------------------------------
// DCT / DST embeddings (shared by the line kernel's LOAD phase and the elementwise pass of long lines): element e of the
// M-point complex line built from the real input line (LD_X_* of fft_types.h).  JB = LineJob or AuxJob (same field names).
template <typename T, typename JB>
IMP_HD cx<T> x_embed(const JB &J, int mode, const T *inr, int64_t off, int64_t es, uint32_t e, uint32_t N, uint32_t M) {
  switch (mode) {
    case LD_X_ZPAD:
      return e < N ? mk<T>(inr[off + (int64_t)e * es], (T)0) : mk<T>((T)0, (T)0);
    case LD_X_TW: {
      if (e >= N) return mk<T>((T)0, (T)0);
      const T f = (T)(e == 0 ? J.x_f0 : J.x_f);
      const cx<T> w = IMP_LDG((const cx<T> *)J.x_tw + 2 * e);
      const T xv = inr[off + (int64_t)e * es] * f;
      return mk<T>(xv * w.x, xv * w.y);
    }
    case LD_X_TW_SHIFT: {
      if (e < 1 || e > N) return mk<T>((T)0, (T)0);
      const T f = (T)(e == N ? J.x_fl : J.x_f) * (T)(e == 1 ? J.x_f0 : 1.0);
      const cx<T> w = IMP_LDG((const cx<T> *)J.x_tw + 2 * e);
      const T xv = inr[off + (int64_t)(e - 1) * es] * f;
      return mk<T>(xv * w.x, xv * w.y);
    }
    case LD_X_SYM: {
      const uint32_t j = e < N ? e : M - e;
      const T f = (T)((j == 0 || j == N - 1) ? J.x_f0 : 1.0);
      return mk<T>(inr[off + (int64_t)j * es] * f, (T)0);
    }
    default: {  // LD_X_ASYM
      T xv = (T)0;
      if (e >= 1 && e <= N) xv = inr[off + (int64_t)(e - 1) * es];
      else if (e > N + 1) xv = -inr[off + (int64_t)(M - e - 1) * es];
      return mk<T>(xv, (T)0);
    }
  }
}
// ... and real output e from bin e + x_shift of the transformed line (ST_X); f = the caller's scale factor
template <typename T, typename JB>
IMP_HD T x_extract(const JB &J, cx<T> v, uint32_t e, uint32_t N, T f) {
  if (J.x_wadd != 0xffffffffu) v = cmul(v, IMP_LDG((const cx<T> *)J.x_tw + (2 * e + J.x_wadd)));
  T y = (J.x_im ? -v.y : v.x) * (T)J.x_s * f;
  if (e == 0) y *= (T)J.x_s0;
  if (e == N - 1) y *= (T)J.x_sn;
  return y;
}

template <typename T>
IMP_HD void load_one(const LineJob &J, cx<T> *S /*line base*/, int64_t off, uint32_t e) {
  const bool cin = J.flags & F_CONJ_IN, cseq = J.flags & F_CONJ_SEQ;
  const T *inr = (const T *)J.in;
  const cx<T> *inc = (const cx<T> *)J.in;
  const int64_t es = J.es_in;
  switch (J.load_mode) {
    case LD_C: {
      cx<T> v;
      if (J.seg_len) {
        const uint32_t sg = e / J.seg_len, w = e - sg * J.seg_len;
        v = ((const cx<T> *)J.seg_base[sg])[off + (int64_t)w * es];
      } else {
        v = inc[off + (int64_t)e * es];
      }
      S[phys<T>(J, e)] = cconj_if(v, cin != cseq);
    } break;
    case LD_R_PAIRS: {
      cx<T> v;
      if (J.flags & F_VEC_IN) {
        v = *(const cx<T> *)(inr + off + 2 * (int64_t)e);
      } else {
        v.x = inr[off + (2 * (int64_t)e) * es];
        v.y = inr[off + (2 * (int64_t)e + 1) * es];
      }
      if ((J.flags & F_NEG_EVEN_IN) && e > 0) v.x = -v.x;
      S[phys<T>(J, e)] = v;
    } break;
    case LD_R_ZEROIM: {
      T xv = inr[off + (int64_t)e * es];
      if ((J.flags & F_NEG_EVEN_IN) && e > 0 && !(e & 1)) xv = -xv;
      S[phys<T>(J, e)] = mk<T>(xv, (T)0);
    } break;
    case LD_HERM_EVEN: {
      cx<T> v = inc[off + (int64_t)e * es];
      if (e == 0 || e == J.n_seq) v.y = (T)0;
      S[phys<T>(J, e)] = cconj_if(v, cin);  // F_CONJ_SEQ is applied by OP_C2R_PRE_EVEN
    } break;
    case LD_HERM_FULL: {
      cx<T> v = inc[off + (int64_t)e * es];
      if (e == 0) v.y = (T)0;
      v = cconj_if(v, cin != cseq);
      S[phys<T>(J, e)] = v;
      if (e > 0) S[phys<T>(J, J.n_real - e)] = cconj(v);
    } break;
    case LD_HC_EVEN: {
      cx<T> v;
      v.x = inr[off + (e == 0 ? 0 : 2 * (int64_t)e - 1) * es];
      v.y = (e == 0 || e == J.n_seq) ? (T)0 : inr[off + (2 * (int64_t)e) * es];
      S[phys<T>(J, e)] = cconj_if(v, cin);
    } break;
    case LD_HC_FULL: {
      cx<T> v;
      v.x = inr[off + (e == 0 ? 0 : 2 * (int64_t)e - 1) * es];
      v.y = (e == 0) ? (T)0 : inr[off + (2 * (int64_t)e) * es];
      v = cconj_if(v, cin != cseq);
      S[phys<T>(J, e)] = v;
      if (e > 0) S[phys<T>(J, J.n_real - e)] = cconj(v);
    } break;
    case LD_X_ZPAD: case LD_X_TW: case LD_X_TW_SHIFT: case LD_X_SYM: case LD_X_ASYM:
      S[phys<T>(J, e)] = x_embed<T>(J, J.load_mode, inr, off, es, e, J.n_real, J.n_seq);
      break;
    default: break;
  }
}

template <typename T>
IMP_HD void phase_load(const LineJob &J, const TileCtx &tc, uint32_t tid, uint32_t nthr,
                       const int64_t *offs, cx<T> *smem) {
  const uint32_t n = J.n_load;
  if (J.flags & F_IN_LINES_FAST) {
    const uint32_t cmask = (1u << J.log_c) - 1;
    const uint64_t total = (uint64_t)n << J.log_c;
    for (uint64_t g = tid; g < total; g += nthr) {
      uint32_t c = (uint32_t)g & cmask, e = (uint32_t)(g >> J.log_c);
      if (c < tc.nlines) load_one<T>(J, smem + (size_t)c * J.pitch, offs[3 * c], e);
    }
  } else {
    for (uint32_t c = 0; c < tc.nlines; ++c) {
      cx<T> *S = smem + (size_t)c * J.pitch;
      const int64_t off = offs[3 * c];
      for (uint32_t e = tid; e < n; e += nthr) load_one<T>(J, S, off, e);
    }
  }
}

// ---------------------------------------------------------------------------
// radix passes
// ---------------------------------------------------------------------------
template <typename T, int IP>
IMP_HD void pass_radix(const LineJob &J, const Phase &P, uint32_t tid, uint32_t nthr, cx<T> *smem) {
  const uint32_t ido = P.ido, l1 = P.l1;
  const uint32_t nb = J.n_fft / IP;
  const uint32_t cmask = (1u << J.log_c) - 1;
  const uint32_t total = nb << J.log_c;
  const cx<T> *tw = (const cx<T> *)J.tw;
  const bool dit = P.op == OP_PASS_DIT;
  const bool bfast = J.swz_mask != 0;  // few long lines per CTA: threads walk one line (XOR-swizzled)
  for (uint32_t g = tid; g < total; g += nthr) {
    uint32_t c, b;
    if (bfast) { c = g / nb; b = g - c * nb; } else { c = g & cmask; b = g >> J.log_c; }
    const uint32_t k = b / ido, i = b - k * ido;
    const uint32_t base = i + ido * IP * k;
    cx<T> *S = smem + (size_t)c * J.pitch;
    cx<T> x[IP];
#pragma unroll
    for (int j = 0; j < IP; ++j) x[j] = S[phys<T>(J, base + j * ido)];
    if (ido > 1) {
      const uint32_t t1 = i * l1;
      if (dit) {
#pragma unroll
        for (int j = 1; j < IP; ++j) x[j] = cmul(x[j], IMP_LDG(tw + t1 * j));
        Bfly<T, IP>::run(x);
      } else {
        Bfly<T, IP>::run(x);
#pragma unroll
        for (int j = 1; j < IP; ++j) x[j] = cmul(x[j], IMP_LDG(tw + t1 * j));
      }
    } else {
      Bfly<T, IP>::run(x);
    }
#pragma unroll
    for (int j = 0; j < IP; ++j) S[phys<T>(J, base + j * ido)] = x[j];
  }
}

// Large odd radices keep up to 31 complex values live; compiled as separate functions so that their
// register demand (and spills) stay out of the hot radix-2..9 path of the kernel.
#if defined(__CUDACC__)
#define IMP_NOINLINE __device__ __noinline__
#else
#define IMP_NOINLINE
#endif
template <typename T, int IP>
IMP_NOINLINE void pass_radix_big(const LineJob &J, const Phase &P, uint32_t tid, uint32_t nthr, cx<T> *smem) {
  pass_radix<T, IP>(J, P, tid, nthr, smem);
}

template <typename T>
IMP_HD void phase_pass(const LineJob &J, const Phase &P, uint32_t tid, uint32_t nthr, cx<T> *smem) {
  switch (P.radix) {
    case 2: pass_radix<T, 2>(J, P, tid, nthr, smem); break;
    case 3: pass_radix<T, 3>(J, P, tid, nthr, smem); break;
    case 4: pass_radix<T, 4>(J, P, tid, nthr, smem); break;
    case 5: pass_radix<T, 5>(J, P, tid, nthr, smem); break;
    case 7: pass_radix<T, 7>(J, P, tid, nthr, smem); break;
    case 8: pass_radix<T, 8>(J, P, tid, nthr, smem); break;
    case 9: pass_radix<T, 9>(J, P, tid, nthr, smem); break;
    case 11: pass_radix_big<T, 11>(J, P, tid, nthr, smem); break;
    case 13: pass_radix_big<T, 13>(J, P, tid, nthr, smem); break;
    case 17: pass_radix_big<T, 17>(J, P, tid, nthr, smem); break;
    case 19: pass_radix_big<T, 19>(J, P, tid, nthr, smem); break;
    case 23: pass_radix_big<T, 23>(J, P, tid, nthr, smem); break;
    case 29: pass_radix_big<T, 29>(J, P, tid, nthr, smem); break;
    case 31: pass_radix_big<T, 31>(J, P, tid, nthr, smem); break;
    default: break;
  }
}

// ---------------------------------------------------------------------------
// element-wise phases (Bluestein, c2r pre-twiddle)
// ---------------------------------------------------------------------------
template <typename T>
IMP_HD void phase_elementwise(const LineJob &J, const Phase &P, uint32_t tid, uint32_t nthr, cx<T> *smem) {
  const uint32_t cmask = (1u << J.log_c) - 1;
  const uint32_t L = J.n_seq;
  uint32_t n;
  switch (P.op) {
    case OP_BLUE_PRE: n = J.n_fft; break;
    case OP_BLUE_MUL: n = J.n_fft; break;
    case OP_BLUE_POST: n = L; break;
    case OP_MUL_CONJ_BK: n = L; break;
    case OP_C2R_PRE_EVEN: n = L / 2 + 1; break;
    default: return;
  }
  const uint32_t total = n << J.log_c;
  const cx<T> *bk = (const cx<T> *)J.bk;
  const cx<T> *bkf = (const cx<T> *)J.bkf;
  const cx<T> *twr = (const cx<T> *)J.tw_r;
  const bool bfast = J.swz_mask != 0;
  for (uint32_t g = tid; g < total; g += nthr) {
    uint32_t c, e;
    if (bfast) { c = g / n; e = g - c * n; } else { c = g & cmask; e = g >> J.log_c; }
    cx<T> *S = smem + (size_t)c * J.pitch;
    switch (P.op) {
      case OP_BLUE_PRE: {
        const uint32_t p = phys<T>(J, e);
        S[p] = e < L ? cmul(S[p], cconj(IMP_LDG(bk + e))) : mk<T>((T)0, (T)0);
      } break;
      case OP_BLUE_MUL: {
        const uint32_t p = phys<T>(J, e);
        S[p] = cconj(cmul(S[p], IMP_LDG(bkf + e)));
      } break;
      case OP_BLUE_POST: {
        const uint32_t p = phys<T>(J, e);
        S[p] = cconj(cmul(S[p], IMP_LDG(bk + e)));
      } break;
      case OP_MUL_CONJ_BK: {
        const uint32_t p = phys<T>(J, e);
        S[p] = cmul(S[p], cconj(IMP_LDG(bk + e)));
      } break;
      case OP_C2R_PRE_EVEN: {
        // X[0..M] -> Z[k] = (X[k]+conj X[M-k]) + i e^{+2 pi i k/N} (X[k]-conj X[M-k])
        const uint32_t M = L, k = e, km = M - k;
        const uint32_t pk = phys<T>(J, k), pm = phys<T>(J, km);
        const cx<T> a = S[pk], b = S[pm];
        const cx<T> w = cconj(IMP_LDG(twr + k));  // e^{+2 pi i k/N}
        const bool cs = J.flags & F_CONJ_SEQ;
        {
          const cx<T> s = cadd(a, cconj(b)), d = csub(a, cconj(b));
          const cx<T> z = cadd(s, mul_pi(cmul(w, d)));
          S[pk] = cconj_if(z, cs);
        }
        if (k != 0 && k != km) {
          const cx<T> s = cadd(b, cconj(a)), d = csub(b, cconj(a));
          // e^{+2 pi i (M-k)/N} = -conj(w)
          const cx<T> wm = mk<T>(-w.x, w.y);
          const cx<T> z = cadd(s, mul_pi(cmul(wm, d)));
          S[pm] = cconj_if(z, cs);
        }
      } break;
      default: break;
    }
  }
}

// ---------------------------------------------------------------------------
// STORE
// ---------------------------------------------------------------------------
template <typename T>
IMP_HD cx<T> read_bin(const LineJob &J, const cx<T> *S, uint32_t k) {
  const uint32_t p = J.perm ? IMP_LDG(J.perm + k) : k;
  return cconj_if(S[phys<T>(J, p)], (J.flags & F_CONJ_OUT) != 0);
}

// bin k (0..M) of the even-N real transform from the M-point FFT Z of the packed line
template <typename T>
IMP_HD cx<T> r2c_even_bin(const LineJob &J, const cx<T> *S, uint32_t k) {
  const uint32_t M = J.n_seq;
  const cx<T> a = read_bin<T>(J, S, k == M ? 0 : k);
  const cx<T> b = cconj(read_bin<T>(J, S, k == 0 ? 0 : M - k));
  const T h = (T)0.5;
  const cx<T> E = mk<T>((a.x + b.x) * h, (a.y + b.y) * h);
  const cx<T> D = mk<T>((a.x - b.x) * h, (a.y - b.y) * h);
  const cx<T> O = mul_mi(D);
  const cx<T> w = IMP_LDG((const cx<T> *)J.tw_r + k);
  return cadd(E, cmul(w, O));
}

template <typename T>
IMP_HD void store_one(const LineJob &J, const cx<T> *S, int64_t off, uint32_t e, uint32_t tw_idx) {
  T *outr = (T *)J.out;
  cx<T> *outc = (cx<T> *)J.out;
  const int64_t es = J.es_out;
  const T f = (T)J.fct;
  const bool cres = J.flags & F_CONJ_RESULT;
  switch (J.store_mode) {
    case ST_C:
    case ST_HERM_HALF: {
      cx<T> v;
      if (J.zero_pad_from && e >= J.zero_pad_from) {
        v = mk<T>((T)0, (T)0);
      } else {
        v = read_bin<T>(J, S, e);
        if (J.tw4_n) {  // four-step twiddle W_N^(e*tw_idx), conjugated for the backward transform
          const uint32_t m = e * tw_idx;
          const cx<T> w = cmul(IMP_LDG((const cx<T> *)J.tw4_hi + (m >> J.tw4_shift)),
                               IMP_LDG((const cx<T> *)J.tw4_lo + (m & ((1u << J.tw4_shift) - 1))));
          v = cmul(v, cconj_if(w, (J.flags & F_CONJ_OUT) != 0));
        }
        if (J.mul_tab) v = cmul(v, IMP_LDG((const cx<T> *)J.mul_tab + (tw_idx + J.mul_stride * e)));
        if (J.umul_mod) {
          uint64_t o = (uint64_t)(off + (int64_t)e * es);
          if (o >= J.umul_mod) o %= J.umul_mod;
          v = cmul(v, IMP_LDG((const cx<T> *)J.umul + o));
        }
        v.x *= f; v.y *= f;
      }
      outc[off + (int64_t)e * es] = cconj_if(v, cres);
    } break;
    case ST_HERM_SYM: {
      cx<T> v = read_bin<T>(J, S, e);
      v.x *= f; v.y *= f;
      v = cconj_if(v, cres);
      outc[off + (int64_t)e * es] = v;
      if (e > 0) outc[off + (int64_t)(J.n_real - e) * es] = cconj(v);
    } break;
    case ST_R2C_EVEN: {
      cx<T> v = r2c_even_bin<T>(J, S, e);
      v.x *= f; v.y *= f;
      outc[off + (int64_t)e * es] = cconj_if(v, cres);
    } break;
    case ST_R2C_EVEN_SYM: {
      cx<T> v = r2c_even_bin<T>(J, S, e);
      v.x *= f; v.y *= f;
      v = cconj_if(v, cres);
      outc[off + (int64_t)e * es] = v;
      if (e > 0 && e < J.n_seq) outc[off + (int64_t)(J.n_real - e) * es] = cconj(v);
    } break;
    case ST_R_PAIRS: {
      cx<T> v = read_bin<T>(J, S, e);
      v.x *= f; v.y *= f;
      if ((J.flags & F_NEG_EVEN_OUT) && e > 0) v.x = -v.x;
      if (J.flags & F_VEC_OUT) {
        *(cx<T> *)(outr + off + 2 * (int64_t)e) = v;
      } else {
        outr[off + (2 * (int64_t)e) * es] = v.x;
        outr[off + (2 * (int64_t)e + 1) * es] = v.y;
      }
    } break;
    case ST_R_REALPART: {
      T y = read_bin<T>(J, S, e).x * f;
      if ((J.flags & F_NEG_EVEN_OUT) && e > 0 && !(e & 1)) y = -y;
      outr[off + (int64_t)e * es] = y;
    } break;
    case ST_HARTLEY_EVEN:
    case ST_HARTLEY_FULL: {
      cx<T> v = J.store_mode == ST_HARTLEY_EVEN ? r2c_even_bin<T>(J, S, e) : read_bin<T>(J, S, e);
      v.x *= f; v.y *= f;
      if (e == 0 || 2 * e == J.n_real) {
        outr[off + (int64_t)e * es] = v.x;
      } else {
        outr[off + (int64_t)e * es] = v.x + v.y;
        outr[off + (int64_t)(J.n_real - e) * es] = v.x - v.y;
      }
    } break;
    case ST_HC_EVEN: {
      cx<T> v = r2c_even_bin<T>(J, S, e);
      v.x *= f; v.y *= cres ? -f : f;   // r2c with forward=false packs the conjugate spectrum
      if (e == 0) outr[off] = v.x;
      else if (e == J.n_seq) outr[off + (int64_t)(J.n_real - 1) * es] = v.x;
      else { outr[off + (2 * (int64_t)e - 1) * es] = v.x; outr[off + (2 * (int64_t)e) * es] = v.y; }
    } break;
    case ST_X:
      outr[off + (int64_t)e * es] = x_extract<T>(J, read_bin<T>(J, S, e + J.x_shift), e, J.n_real, f);
      break;
    case ST_HC_FULL: {
      cx<T> v = read_bin<T>(J, S, e);
      v.x *= f; v.y *= cres ? -f : f;   // r2c with forward=false packs the conjugate spectrum
      if (e == 0) outr[off] = v.x;
      else { outr[off + (2 * (int64_t)e - 1) * es] = v.x; outr[off + (2 * (int64_t)e) * es] = v.y; }
    } break;
    default: break;
  }
}

template <typename T>
IMP_HD void phase_store(const LineJob &J, const TileCtx &tc, uint32_t tid, uint32_t nthr,
                        const int64_t *offs, const cx<T> *smem) {
  const uint32_t n = J.n_store;
  if (J.flags & F_OUT_LINES_FAST) {
    const uint32_t cmask = (1u << J.log_c) - 1;
    const uint64_t total = (uint64_t)n << J.log_c;
    for (uint64_t g = tid; g < total; g += nthr) {
      uint32_t c = (uint32_t)g & cmask, e = (uint32_t)(g >> J.log_c);
      if (c < tc.nlines) store_one<T>(J, smem + (size_t)c * J.pitch, offs[3 * c + 1], e, (uint32_t)offs[3 * c + 2]);
    }
  } else {
    for (uint32_t c = 0; c < tc.nlines; ++c) {
      const cx<T> *S = smem + (size_t)c * J.pitch;
      const int64_t off = offs[3 * c + 1];
      const uint32_t twi = (uint32_t)offs[3 * c + 2];
      for (uint32_t e = tid; e < n; e += nthr) store_one<T>(J, S, off, e, twi);
    }
  }
}

// One middle phase (everything between LOAD and STORE).
template <typename T>
IMP_HD void phase_mid(const LineJob &J, const Phase &P, uint32_t tid, uint32_t nthr, cx<T> *smem) {
  if (P.op == OP_PASS_DIF || P.op == OP_PASS_DIT) phase_pass<T>(J, P, tid, nthr, smem);
  else phase_elementwise<T>(J, P, tid, nthr, smem);
}

// Genuine Hartley fold of element `idx` of the half spectrum (see CombineJob).  Elements whose mirror image
// lies inside the half spectrum as well (index 0 or N/2 along the halved axis) write only themselves: the
// mirror is written by its own element, so no output is written twice.
template <typename T>
IMP_HD void hartley_combine_one(const CombineJob &C, uint64_t idx) {
  const cx<T> v = ((const cx<T> *)C.in)[idx];
  int64_t off = 0, roff = 0;
  uint32_t kh = 0;
  uint64_t r = idx;
  for (int d = C.ndim - 1; d >= 0; --d) {
    const uint32_t k = (uint32_t)(r % C.hshape[d]);
    r /= C.hshape[d];
    const uint32_t rk = C.rev[d] ? (k == 0 ? 0 : C.full[d] - k) : k;
    off += (int64_t)k * C.so[d];
    roff += (int64_t)rk * C.so[d];
    if ((uint32_t)d == C.half_axis) kh = k;
  }
  T *out = (T *)C.out;
  out[off] = v.x + v.y;
  if (kh != 0 && 2 * kh != C.full[C.half_axis]) out[roff] = v.x - v.y;
}

// ---------------------------------------------------------------------------
// AuxJob: one element of the iteration space (see fft_types.h)
// ---------------------------------------------------------------------------
// store spectrum bin k of a length-N real transform in the user layout (RealLayout values of planner.h:
// 0 Hermitian complex, 1 FFTPACK halfcomplex reals, 2 full symmetric complex, 4 Hartley reals)
template <typename T>
IMP_HD void aux_store_bin(const AuxJob &A, int64_t off, int64_t es, uint32_t k, cx<T> v) {
  const T f = (T)A.fct;
  v.x *= f; v.y *= f;
  if (A.flags & F_CONJ_RESULT) v.y = -v.y;
  const uint32_t N = A.N;
  cx<T> *oc = (cx<T> *)A.out;
  T *orl = (T *)A.out;
  switch (A.layout) {
    case 1:
      if (k == 0) orl[off] = v.x;
      else if (2 * k == N) orl[off + (int64_t)(N - 1) * es] = v.x;
      else { orl[off + (2 * (int64_t)k - 1) * es] = v.x; orl[off + (2 * (int64_t)k) * es] = v.y; }
      break;
    case 2:
      oc[off + (int64_t)k * es] = v;
      if (k > 0 && 2 * k != N) oc[off + (int64_t)(N - k) * es] = cconj(v);
      break;
    case 4:
      if (k == 0 || 2 * k == N) orl[off + (int64_t)k * es] = v.x;
      else { orl[off + (int64_t)k * es] = v.x + v.y; orl[off + (int64_t)(N - k) * es] = v.x - v.y; }
      break;
    default:
      oc[off + (int64_t)k * es] = v;
      break;
  }
}

// load spectrum bin k (0 <= k <= N/2) of a Hermitian input in the user layout (0 Hermitian, 1 halfcomplex)
template <typename T>
IMP_HD cx<T> aux_load_bin(const AuxJob &A, int64_t off, int64_t es, uint32_t k) {
  const uint32_t N = A.N;
  cx<T> v;
  if (A.layout == 1) {
    const T *ir = (const T *)A.in;
    if (k == 0) v = mk<T>(ir[off], (T)0);
    else if (2 * k == N) v = mk<T>(ir[off + (int64_t)(N - 1) * es], (T)0);
    else v = mk<T>(ir[off + (2 * (int64_t)k - 1) * es], ir[off + (2 * (int64_t)k) * es]);
  } else {
    v = ((const cx<T> *)A.in)[off + (int64_t)k * es];
    if (k == 0 || 2 * k == N) v.y = (T)0;   // pocketfft ignores these imaginary parts
  }
  if (A.flags & F_CONJ_IN) v.y = -v.y;
  return v;
}

template <typename T>
IMP_HD void aux_one(const AuxJob &A, uint64_t idx) {
  // decode: last dimension = position along the transform axis
  const int ax = A.ndim - 1;
  uint32_t e;
  int64_t ou = 0, ow = 0;
  if (A.total <= 0xffffffffull) {   // the usual case: 32-bit index arithmetic
    uint32_t r = (uint32_t)idx / A.shape[ax];
    e = (uint32_t)idx - r * A.shape[ax];
    for (int d = ax - 1; d >= 0; --d) {
      const uint32_t q = r / A.shape[d], i = r - q * A.shape[d];
      r = q;
      ou += (int64_t)i * A.s_user[d];
      ow += (int64_t)i * A.s_work[d];
    }
  } else {
    e = (uint32_t)(idx % A.shape[ax]);
    uint64_t r = idx / A.shape[ax];
    for (int d = ax - 1; d >= 0; --d) {
      const uint32_t i = (uint32_t)(r % A.shape[d]);
      r /= A.shape[d];
      ou += (int64_t)i * A.s_user[d];
      ow += (int64_t)i * A.s_work[d];
    }
  }
  const int64_t eu = A.s_user[ax], ew = A.s_work[ax];
  const uint32_t N = A.N, M = A.M;
  const cx<T> *tab = (const cx<T> *)A.tab;
  switch (A.mode) {
    case AUX_R2C_POST_EVEN: {   // e = k in 0..M
      const cx<T> *Z = (const cx<T> *)A.in + ow;
      const cx<T> a = Z[(int64_t)(e == M ? 0 : e) * ew], b = cconj(Z[(int64_t)(e == 0 ? 0 : M - e) * ew]);
      const T h = (T)0.5;
      const cx<T> E = mk<T>((a.x + b.x) * h, (a.y + b.y) * h), D = mk<T>((a.x - b.x) * h, (a.y - b.y) * h);
      aux_store_bin<T>(A, ou, eu, e, cadd(E, cmul(IMP_LDG(tab + e), mul_mi(D))));
    } break;
    case AUX_C2R_PRE_EVEN: {    // e = k in 0..M/2: writes Z[k] and Z[M-k]
      cx<T> *Z = (cx<T> *)A.out + ow;
      const uint32_t k = e, km = M - k;
      const cx<T> a = aux_load_bin<T>(A, ou, eu, k), b = aux_load_bin<T>(A, ou, eu, km);
      const cx<T> w = cconj(IMP_LDG(tab + k));   // e^{+2 pi i k/N}
      {
        const cx<T> s = cadd(a, cconj(b)), d = csub(a, cconj(b));
        Z[(int64_t)(k == M ? 0 : k) * ew] = cadd(s, mul_pi(cmul(w, d)));
      }
      if (k != 0 && k != km) {
        const cx<T> s = cadd(b, cconj(a)), d = csub(b, cconj(a));
        const cx<T> wm = mk<T>(-w.x, w.y);
        Z[(int64_t)km * ew] = cadd(s, mul_pi(cmul(wm, d)));
      }
    } break;
    case AUX_R2C_PRE_ODD: {
      T xv = ((const T *)A.in)[ou + (int64_t)e * eu];
      if ((A.flags & F_NEG_EVEN_IN) && e > 0 && !(e & 1)) xv = -xv;
      ((cx<T> *)A.out)[ow + (int64_t)e * ew] = mk<T>(xv, (T)0);
    } break;
    case AUX_R2C_PACK_EVEN: {   // e = m in 0..M
      const T *x = (const T *)A.in + ou;
      cx<T> v = mk<T>(x[(2 * (int64_t)e) * eu], x[(2 * (int64_t)e + 1) * eu]);
      if ((A.flags & F_NEG_EVEN_IN) && e > 0) v.x = -v.x;
      ((cx<T> *)A.out)[ow + (int64_t)e * ew] = v;
    } break;
    case AUX_C2R_UNPACK_EVEN: { // e = m in 0..M
      cx<T> v = ((const cx<T> *)A.in)[ow + (int64_t)e * ew];
      const T f = (T)A.fct;
      v.x *= f; v.y *= f;
      if ((A.flags & F_NEG_EVEN_OUT) && e > 0) v.x = -v.x;
      T *x = (T *)A.out + ou;
      x[(2 * (int64_t)e) * eu] = v.x;
      x[(2 * (int64_t)e + 1) * eu] = v.y;
    } break;
    case AUX_R2C_POST_ODD:
      aux_store_bin<T>(A, ou, eu, e, ((const cx<T> *)A.in)[ow + (int64_t)e * ew]);
      break;
    case AUX_C2R_PRE_ODD: {     // e = k in 0..(N-1)/2
      const cx<T> v = aux_load_bin<T>(A, ou, eu, e);
      cx<T> *Z = (cx<T> *)A.out + ow;
      Z[(int64_t)e * ew] = v;
      if (e > 0) Z[(int64_t)(N - e) * ew] = cconj(v);
    } break;
    case AUX_C2R_POST_ODD: {
      T y = ((const cx<T> *)A.in)[ow + (int64_t)e * ew].x * (T)A.fct;
      if ((A.flags & F_NEG_EVEN_OUT) && e > 0 && !(e & 1)) y = -y;
      ((T *)A.out)[ou + (int64_t)e * eu] = y;
    } break;
    case AUX_X_EMBED:           // e in 0..M
      ((cx<T> *)A.out)[ow + (int64_t)e * ew] = x_embed<T>(A, A.x_load, (const T *)A.in, ou, eu, e, N, M);
      break;
    case AUX_X_EXTRACT:         // e in 0..N
      ((T *)A.out)[ou + (int64_t)e * eu] = x_extract<T>(A, ((const cx<T> *)A.in)[ow + (int64_t)(e + A.x_shift) * ew], e, N, (T)A.fct);
      break;
    case AUX_BLUE_PRE: {        // e = n in 0..n2
      cx<T> v = mk<T>((T)0, (T)0);
      if (e < N) {
        v = ((const cx<T> *)A.in)[ou + (int64_t)e * eu];
        if (A.flags & F_CONJ_IN) v.y = -v.y;
        v = cmul(v, cconj(IMP_LDG(tab + e)));
      }
      ((cx<T> *)A.out)[ow + (int64_t)e * ew] = v;
    } break;
    case AUX_BLUE_POST: {       // e = k in 0..L
      cx<T> v = cmul(((const cx<T> *)A.in)[ow + (int64_t)e * ew], cconj(IMP_LDG(tab + e)));
      const T f = (T)A.fct;
      v.x *= f; v.y *= f;
      if (A.flags & F_CONJ_RESULT) v.y = -v.y;
      ((cx<T> *)A.out)[ou + (int64_t)e * eu] = v;
    } break;
    default: break;
  }
}

}  // namespace impulse
