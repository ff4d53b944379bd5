/* TEST INFRASTRUCTURE — NOT PRODUCT CODE.
 *
 * CPU restatement ("port") of the algorithm the reference's FFT path runs, written
 * from scratch for this repo as the parity oracle.  Only tests/, the smoke test and
 * bench.py's cpu_baseline / --impl reference legs may load it; impulse_b200/ never does.
 *
 * Reference = /root/reference/impulse/fft/c_pocketfft/pocketfft.c (cited as C:line).
 * What is restated, and where it deliberately differs:
 *   - plan choice fftpack-vs-Bluestein            C:2066-2093 (complex), C:2126-2153 (real)
 *   - largest_prime_factor / cost_guess / good_size  C:205-260
 *   - factor order: 4s, one 2 moved to the front, odd divisors ascending   C:953-983
 *   - pass driver: one out-of-place Stockham pass per factor, ping-pong, final scale  C:871-929
 *   - each pass computes  CH(i,k,jo) = W^(i*jo*l1) * sum_j CC(i,j,k) w_ip^(j*jo)
 *     with CC/CH indexed as at C:286-287.  The reference hard-codes radix 2/3/4/5/7/11
 *     (C:300-761) and uses a symmetric O(ip^2) form for others (C:767-865); here ONE
 *     generic O(ip^2) pass serves every radix.  Same mathematics, rounding differs in
 *     the last bits (measured <= 1e-15 rel-L2, tests/test_oracle.py).
 *   - twiddles: exact-argument long double sin/cos instead of the octant polynomial
 *     scheme of C:36-203 (both are correctly rounded to ~0.5 ulp).
 *   - Bluestein plan and execution                 C:1889-2008
 *   - real transforms: the reference runs FFTPACK real passes radf/radb (C:1082-1766)
 *     for fftpack lengths; this port always goes through the complex transform and
 *     packs/unpacks the halfcomplex layout exactly like the reference's own Bluestein
 *     wrappers rfftblue_forward/backward do (C:2019-2058).  Output layout identical.
 *
 * Pinned against: README known answers, tests/test_fft2.nim vectors, and the compiled
 * reference itself (oracle/_ref) over all lengths 1..8191 — see tests/test_oracle.py.
 *
 * Compile twice: default (double) and -DPORT_FLOAT (float arithmetic, float I/O) so
 * fp32 paths have a same-precision checker.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifdef PORT_FLOAT
typedef float real;
#define SYM(x) x##_f32
#else
typedef double real;
#define SYM(x) x##_f64
#endif

typedef struct { real r, i; } cpx;

/* ---- plan heuristics (C:205-260) ---------------------------------------- */
static size_t lpf(size_t n) {
  size_t res = 1;
  while ((n & 1) == 0) { res = 2; n >>= 1; }
  for (size_t x = 3; x * x <= n; x += 2)
    while (n % x == 0) { res = x; n /= x; }
  if (n > 1) res = n;
  return res;
}

static double cost_guess(size_t n) {
  const double lfp = 1.1;
  size_t ni = n;
  double result = 0.;
  while ((n & 1) == 0) { result += 2; n >>= 1; }
  for (size_t x = 3; x * x <= n; x += 2)
    while (n % x == 0) { result += (x <= 5) ? (double)x : lfp * (double)x; n /= x; }
  if (n > 1) result += (n <= 5) ? (double)n : lfp * (double)n;
  return result * (double)ni;
}

size_t SYM(port_good_size)(size_t n) {
  if (n <= 6) return n;
  size_t best = 2 * n;
  for (size_t f2 = 1; f2 < best; f2 *= 2)
    for (size_t f23 = f2; f23 < best; f23 *= 3)
      for (size_t f235 = f23; f235 < best; f235 *= 5)
        for (size_t f2357 = f235; f2357 < best; f2357 *= 7)
          for (size_t f = f2357; f < best; f *= 11)
            if (f >= n) best = f;
  return best;
}

/* returns 1 when the reference would pick Bluestein for a complex (real=0) or
 * real (real=1) transform of this length: C:2073-2091 / C:2133-2151 */
int SYM(port_uses_bluestein)(size_t n, int real_input) {
  if (n < 50) return 0;
  size_t l = lpf(n);
  if ((double)l <= sqrt((double)n)) return 0;
  double comp1 = (real_input ? 0.5 : 1.0) * cost_guess(n);
  double comp2 = 2 * cost_guess(SYM(port_good_size)(2 * n - 1)) * 1.5;
  return comp2 < comp1;
}

/* ---- factorisation (C:953-983) ------------------------------------------ */
#define NFCT 25
static int factorize(size_t n, size_t *fct) {
  int nf = 0;
  while (n % 4 == 0) { if (nf >= NFCT) return -1; fct[nf++] = 4; n >>= 2; }
  if (n % 2 == 0) {
    n >>= 1;
    if (nf >= NFCT) return -1;
    fct[nf++] = 2;
    size_t t = fct[0]; fct[0] = fct[nf - 1]; fct[nf - 1] = t;
  }
  for (size_t d = 3; n > 1 && d * d <= n; d += 2)
    while (n % d == 0) { if (nf >= NFCT) return -1; fct[nf++] = d; n /= d; }
  if (n > 1) { if (nf >= NFCT) return -1; fct[nf++] = n; }
  return nf;
}

int SYM(port_factors)(size_t n, size_t *out) { return factorize(n, out); }

/* ---- twiddles: w[m] = exp(-2*pi*i*m/n), exact argument reduction -------- */
static void unit_root(size_t m, size_t n, long double *c, long double *s) {
  /* reduce m/n to the first octant so the argument handed to cosl/sinl is small */
  static const long double PI = 3.141592653589793238462643383279502884L;
  m %= n;
  size_t m8 = 8 * m; /* position in units of 1/(8n) turns */
  size_t oct = m8 / n;
  size_t rem = m8 - oct * n; /* 0 <= rem < n, angle within octant = 2*pi*rem/(8n) */
  long double x, cr, sr;
  switch (oct) {
    default:
    case 0: x = 2 * PI * (long double)rem / (8 * (long double)n); cr = cosl(x); sr = sinl(x); break;
    case 1: x = 2 * PI * (long double)(n - rem) / (8 * (long double)n); cr = sinl(x); sr = cosl(x); break;
    case 2: x = 2 * PI * (long double)rem / (8 * (long double)n); cr = -sinl(x); sr = cosl(x); break;
    case 3: x = 2 * PI * (long double)(n - rem) / (8 * (long double)n); cr = -cosl(x); sr = sinl(x); break;
    case 4: x = 2 * PI * (long double)rem / (8 * (long double)n); cr = -cosl(x); sr = -sinl(x); break;
    case 5: x = 2 * PI * (long double)(n - rem) / (8 * (long double)n); cr = -sinl(x); sr = -cosl(x); break;
    case 6: x = 2 * PI * (long double)rem / (8 * (long double)n); cr = sinl(x); sr = -cosl(x); break;
    case 7: x = 2 * PI * (long double)(n - rem) / (8 * (long double)n); cr = cosl(x); sr = -sinl(x); break;
  }
  *c = cr;  /* cos(2*pi*m/n) */
  *s = sr;  /* sin(2*pi*m/n) */
}

static cpx *root_table(size_t n) {
  cpx *w = (cpx *)malloc(n * sizeof(cpx));
  if (!w) return NULL;
  for (size_t m = 0; m < n; ++m) {
    long double c, s;
    unit_root(m, n, &c, &s);
    w[m].r = (real)c;
    w[m].i = (real)(-s); /* forward sign */
  }
  return w;
}

/* ---- one Stockham pass, any radix (index convention of C:286-287) ------- */
static void pass_any(size_t ido, size_t ip, size_t l1, const cpx *cc, cpx *ch, const cpx *w,
                     size_t n, int sign) {
  const size_t step = n / ip; /* w_ip^m = w[m*step] */
  for (size_t k = 0; k < l1; ++k)
    for (size_t i = 0; i < ido; ++i)
      for (size_t jo = 0; jo < ip; ++jo) {
        real sr = 0, si = 0;
        for (size_t j = 0; j < ip; ++j) {
          cpx x = cc[i + ido * (j + ip * k)];
          cpx t = w[((j * jo) % ip) * step];
          real ti = sign < 0 ? t.i : -t.i;
          sr += x.r * t.r - x.i * ti;
          si += x.r * ti + x.i * t.r;
        }
        cpx t = w[i * jo * l1];
        real ti = sign < 0 ? t.i : -t.i;
        cpx *o = &ch[i + ido * (k + l1 * jo)];
        o->r = sr * t.r - si * ti;
        o->i = sr * ti + si * t.r;
      }
}

/* pass driver: C:871-929 */
static int cfftp(size_t n, cpx *c, real fct, int sign) {
  if (n == 1) return 0; /* C:875 / C:1704: length-1 plans return BEFORE scaling by fct */
  size_t fctr[NFCT];
  int nf = factorize(n, fctr);
  if (nf < 0) return -1;
  cpx *w = root_table(n);
  cpx *ch = (cpx *)malloc(n * sizeof(cpx));
  if (!w || !ch) { free(w); free(ch); return -1; }
  cpx *p1 = c, *p2 = ch;
  size_t l1 = 1;
  for (int f = 0; f < nf; ++f) {
    size_t ip = fctr[f], ido = n / (l1 * ip);
    pass_any(ido, ip, l1, p1, p2, w, n, sign);
    cpx *t = p1; p1 = p2; p2 = t;
    l1 *= ip;
  }
  for (size_t m = 0; m < n; ++m) { c[m].r = p1[m].r * fct; c[m].i = p1[m].i * fct; }
  free(ch);
  free(w);
  return 0;
}

/* Bluestein: C:1889-2008 */
static int fftblue(size_t n, cpx *c, int sign, real fct) {
  size_t n2 = SYM(port_good_size)(2 * n - 1);
  cpx *bk = (cpx *)malloc(n * sizeof(cpx));
  cpx *bkf = (cpx *)calloc(n2, sizeof(cpx));
  cpx *akf = (cpx *)calloc(n2, sizeof(cpx));
  if (!bk || !bkf || !akf) { free(bk); free(bkf); free(akf); return -1; }
  /* b_k = exp(i*pi*k^2/n); k^2 mod 2n by the recurrence of C:1907-1914 */
  size_t coeff = 0;
  for (size_t m = 0; m < n; ++m) {
    if (m > 0) { coeff += 2 * m - 1; if (coeff >= 2 * n) coeff -= 2 * n; }
    long double cs, sn;
    unit_root(coeff, 2 * n, &cs, &sn);
    bk[m].r = (real)cs;
    bk[m].i = (real)sn;
  }
  real xn2 = (real)1 / (real)n2;
  bkf[0].r = bk[0].r * xn2; bkf[0].i = bk[0].i * xn2;
  for (size_t m = 1; m < n; ++m) {
    bkf[m].r = bkf[n2 - m].r = bk[m].r * xn2;
    bkf[m].i = bkf[n2 - m].i = bk[m].i * xn2;
  }
  int rc = cfftp(n2, bkf, 1, -1);
  /* a_k = c_k * (sign>0 ? b_k : conj b_k), zero padded, FFT */
  for (size_t m = 0; m < n && !rc; ++m) {
    real bi = sign > 0 ? bk[m].i : -bk[m].i;
    akf[m].r = c[m].r * bk[m].r - c[m].i * bi;
    akf[m].i = c[m].r * bi + c[m].i * bk[m].r;
  }
  if (!rc) rc = cfftp(n2, akf, fct, -1);
  /* convolution: multiply by (sign>0 ? conj bkf : bkf) */
  for (size_t m = 0; m < n2 && !rc; ++m) {
    real fi = sign > 0 ? -bkf[m].i : bkf[m].i;
    real re = akf[m].r * bkf[m].r - akf[m].i * fi;
    real im = akf[m].r * fi + akf[m].i * bkf[m].r;
    akf[m].r = re; akf[m].i = im;
  }
  if (!rc) rc = cfftp(n2, akf, 1, +1);
  for (size_t m = 0; m < n && !rc; ++m) {
    real bi = sign > 0 ? bk[m].i : -bk[m].i;
    c[m].r = bk[m].r * akf[m].r - bi * akf[m].i;
    c[m].i = bi * akf[m].r + bk[m].r * akf[m].i;
  }
  free(bk); free(bkf); free(akf);
  return rc;
}

/* ---- public: one complex row in place (C:2104-2118) --------------------- */
int SYM(port_cfft)(real *c, size_t n, int forward, double fct) {
  if (n == 0) return -1;
  int sign = forward ? -1 : 1;
  if (SYM(port_uses_bluestein)(n, 0)) return fftblue(n, (cpx *)c, sign, (real)fct);
  return cfftp(n, (cpx *)c, (real)fct, sign);
}

/* real forward, FFTPACK halfcomplex result in place (layout: pocketfft.nim:228-238) */
int SYM(port_rfft_forward)(real *c, size_t n, double fct) {
  if (n == 0) return -1;
  cpx *tmp = (cpx *)malloc(n * sizeof(cpx));
  if (!tmp) return -1;
  for (size_t m = 0; m < n; ++m) { tmp[m].r = c[m]; tmp[m].i = 0; }
  int rc = SYM(port_uses_bluestein)(n, 1) ? fftblue(n, tmp, -1, (real)fct)
                                          : cfftp(n, tmp, (real)fct, -1);
  if (!rc) {
    c[0] = tmp[0].r;
    memcpy(c + 1, (real *)tmp + 2, (n - 1) * sizeof(real)); /* as C:2053-2054 */
  }
  free(tmp);
  return rc;
}

/* real backward from halfcomplex, in place (C:2019-2040) */
int SYM(port_rfft_backward)(real *c, size_t n, double fct) {
  if (n == 0) return -1;
  cpx *tmp = (cpx *)calloc(n + 1, sizeof(cpx));
  if (!tmp) return -1;
  real *t = (real *)tmp;
  t[0] = c[0]; t[1] = 0;
  memcpy(t + 2, c + 1, (n - 1) * sizeof(real));
  if ((n & 1) == 0) t[n + 1] = 0;
  for (size_t m = 2; m < n; m += 2) { t[2 * n - m] = t[m]; t[2 * n - m + 1] = -t[m + 1]; }
  int rc = SYM(port_uses_bluestein)(n, 1) ? fftblue(n, tmp, +1, (real)fct)
                                          : cfftp(n, tmp, (real)fct, +1);
  if (!rc) for (size_t m = 0; m < n; ++m) c[m] = tmp[m].r;
  free(tmp);
  return rc;
}

/* batched contiguous rows, in place */
int SYM(port_cfft_rows)(real *data, size_t nrows, size_t n, int forward, double fct) {
  for (size_t r = 0; r < nrows; ++r)
    if (SYM(port_cfft)(data + 2 * r * n, n, forward, fct)) return -1;
  return 0;
}
int SYM(port_rfft_rows)(real *data, size_t nrows, size_t n, int forward, double fct) {
  for (size_t r = 0; r < nrows; ++r)
    if (forward ? SYM(port_rfft_forward)(data + r * n, n, fct)
                : SYM(port_rfft_backward)(data + r * n, n, fct)) return -1;
  return 0;
}
