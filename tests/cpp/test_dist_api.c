/* A plain C host driving the multi-GPU entry points of include/impulse_fft_b200.h — what the Nim veneer's importc
 * lines bind.  usage: test_dist_api <ndev>.  Checks against a direct O(N^2) DFT in long double. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "impulse_fft_b200.h"

static int failures = 0;
#define CHECK(cond, ...) do { if (!(cond)) { printf("FAIL %s:%d: ", __FILE__, __LINE__); printf(__VA_ARGS__); printf("\n"); ++failures; } } while (0)

static void dft_rows(const double *x, double *y, size_t rows, size_t n, int forward) {   /* interleaved complex */
  const long double tau = 6.283185307179586476925286766559L * (forward ? -1.0L : 1.0L);
  for (size_t r = 0; r < rows; ++r)
    for (size_t k = 0; k < n; ++k) {
      long double sr = 0, si = 0;
      for (size_t j = 0; j < n; ++j) {
        const long double a = tau * (long double)((j * k) % n) / (long double)n, c = cosl(a), s = sinl(a);
        const long double xr = x[2 * (r * n + j)], xi = x[2 * (r * n + j) + 1];
        sr += xr * c - xi * s; si += xr * s + xi * c;
      }
      y[2 * (r * n + k)] = (double)sr; y[2 * (r * n + k) + 1] = (double)si;
    }
}
static double rel_l2(const double *a, const double *b, size_t n) {
  long double num = 0, den = 0;
  for (size_t i = 0; i < n; ++i) { num += (long double)(a[i] - b[i]) * (a[i] - b[i]); den += (long double)b[i] * b[i]; }
  return (double)sqrtl(num / den);
}

int main(int argc, char **argv) {
  int ndev = argc > 1 ? atoi(argv[1]) : 1;
  if (ndev < 1) ndev = 1;
  if (ndev > 8) ndev = 8;
  int devs[8];
  for (int i = 0; i < ndev; ++i) devs[i] = i;
  srand(7);

  { /* BATCH_SHARD: 37 rows x 96 points complex128, host arrays */
    const size_t rows = 37, n = 96;
    double *x = malloc(16 * rows * n), *y = malloc(16 * rows * n), *w = malloc(16 * rows * n);
    for (size_t i = 0; i < 2 * rows * n; ++i) x[i] = rand() / (double)RAND_MAX - 0.5;
    impulse_fft_desc d;
    memset(&d, 0, sizeof d);
    d.kind = IMPULSE_FFT_C2C; d.dtype = IMPULSE_FFT_F64; d.forward = 1; d.ndim = 2; d.naxes = 1;
    d.shape[0] = rows; d.shape[1] = n; d.axes[0] = 1;
    d.stride_in[0] = d.stride_out[0] = (ptrdiff_t)(16 * n); d.stride_in[1] = d.stride_out[1] = 16;
    impulse_fft_dist h = NULL;
    int rc = impulse_fft_dist_create(&h, IMPULSE_FFT_DIST_BATCH_SHARD, &d, ndev, devs);
    CHECK(rc == 0, "dist_create: %s", impulse_fft_last_error());
    if (!rc) {
      size_t covered = 0;
      for (int i = 0; i < ndev; ++i) { size_t lo, hi; impulse_fft_dist_shard(h, i, &lo, &hi); CHECK(lo == covered, "shard %d starts at %zu", i, lo); covered = hi; }
      CHECK(covered == rows, "shards cover %zu of %zu rows", covered, rows);
      rc = impulse_fft_dist_execute(h, x, y, 1.0);
      CHECK(rc == 0, "dist_execute: %s", impulse_fft_last_error());
      dft_rows(x, w, rows, n, 1);
      CHECK(rel_l2(y, w, 2 * rows * n) < 1e-13, "batch shard rel-L2 %g", rel_l2(y, w, 2 * rows * n));
      impulse_fft_dist_destroy(h);
    }
    free(x); free(y); free(w);
  }
  { /* SLAB_2D: 64 x 64 complex128, host arrays, natural layout back */
    const size_t n = 64;
    double *x = malloc(16 * n * n), *y = malloc(16 * n * n), *t = malloc(16 * n * n), *w = malloc(16 * n * n), *tt = malloc(16 * n * n);
    for (size_t i = 0; i < 2 * n * n; ++i) x[i] = rand() / (double)RAND_MAX - 0.5;
    impulse_fft_desc d;
    memset(&d, 0, sizeof d);
    d.kind = IMPULSE_FFT_C2C; d.dtype = IMPULSE_FFT_F64; d.forward = 1; d.ndim = 2; d.naxes = 2;
    d.shape[0] = d.shape[1] = n; d.axes[0] = 0; d.axes[1] = 1;
    d.stride_in[0] = d.stride_out[0] = (ptrdiff_t)(16 * n); d.stride_in[1] = d.stride_out[1] = 16;
    impulse_fft_dist h = NULL;
    int rc = impulse_fft_dist_create(&h, IMPULSE_FFT_DIST_SLAB_2D, &d, ndev, devs);
    CHECK(rc == 0, "dist_create slab: %s", impulse_fft_last_error());
    if (!rc) {
      rc = impulse_fft_dist_execute(h, x, y, 1.0);
      CHECK(rc == 0, "dist_execute slab: %s", impulse_fft_last_error());
      dft_rows(x, t, n, n, 1);                                  /* rows, transpose, rows, transpose */
      for (size_t i = 0; i < n; ++i) for (size_t j = 0; j < n; ++j) { tt[2 * (j * n + i)] = t[2 * (i * n + j)]; tt[2 * (j * n + i) + 1] = t[2 * (i * n + j) + 1]; }
      dft_rows(tt, t, n, n, 1);
      for (size_t i = 0; i < n; ++i) for (size_t j = 0; j < n; ++j) { w[2 * (j * n + i)] = t[2 * (i * n + j)]; w[2 * (j * n + i) + 1] = t[2 * (i * n + j) + 1]; }
      CHECK(rel_l2(y, w, 2 * n * n) < 1e-13, "slab fft2 rel-L2 %g", rel_l2(y, w, 2 * n * n));
      impulse_fft_dist_destroy(h);
    }
    free(x); free(y); free(t); free(w); free(tt);
  }
  { /* argument errors come back as status codes */
    impulse_fft_desc d;
    memset(&d, 0, sizeof d);
    d.kind = IMPULSE_FFT_C2C; d.dtype = IMPULSE_FFT_F64; d.forward = 1; d.ndim = 2; d.naxes = 1;
    d.shape[0] = 4; d.shape[1] = 8; d.axes[0] = 0;
    d.stride_in[0] = d.stride_out[0] = 128; d.stride_in[1] = d.stride_out[1] = 16;
    impulse_fft_dist h = NULL;
    CHECK(impulse_fft_dist_create(&h, IMPULSE_FFT_DIST_BATCH_SHARD, &d, 1, devs) == IMPULSE_FFT_ERR_INVALID, "axis 0 accepted");
    CHECK(impulse_fft_dist_create(&h, 7, &d, 1, devs) == IMPULSE_FFT_ERR_INVALID, "bad mode accepted");
    CHECK(impulse_fft_dist_create(&h, IMPULSE_FFT_DIST_BATCH_SHARD, &d, 0, devs) == IMPULSE_FFT_ERR_INVALID, "ndev 0 accepted");
  }
  if (failures) { printf("%d check(s) failed\n", failures); return 1; }
  printf("all checks passed on %d device(s)\n", ndev);
  return 0;
}
