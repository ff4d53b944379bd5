// TEST INFRASTRUCTURE — NOT PRODUCT CODE.
//
// extern "C" shim around the *reference's own* FFT engines so that tests and
// bench.py's cpu_baseline / --impl reference legs can call them through ctypes.
// Nothing in impulse_b200/ may link, load or call this file.
//
// It is compiled by oracle/Makefile together with the reference sources where
// they lie under /root/reference (never copied into this repo):
//   * impulse/fft/c_pocketfft/pocketfft.c            (C engine, the 10 symbols of
//                                                     c_pocketfft/pocketfft.h:18-32)
//   * impulse/fft/cpp_pocketfft/pocketfft_hdronly.h  (C++ engine, c2c/r2c/c2r at
//                                                     pocketfft_hdronly.h:3272-3390)
// Output goes to oracle/_ref/ only.
//
// Everything below is glue written for this repo: argument marshalling,
// exception -> return-code mapping and a pthread row driver that shares one
// read-only plan between threads (legal per c_pocketfft/README.md:32-36).

#include <cstddef>
#include <cstring>
#include <complex>
#include <string>
#include <thread>
#include <vector>

#include "pocketfft_hdronly.h"

extern "C" {
#include "pocketfft.h"
}

namespace {
thread_local std::string g_err;

pocketfft::shape_t mk_shape(const size_t *p, size_t n) { return pocketfft::shape_t(p, p + n); }
pocketfft::stride_t mk_stride(const ptrdiff_t *p, size_t n) { return pocketfft::stride_t(p, p + n); }

template <typename F> int guarded(F &&f) {
  try {
    f();
    return 0;
  } catch (const std::exception &e) {
    g_err = e.what();
    return -1;
  } catch (...) {
    g_err = "unknown exception";
    return -1;
  }
}
}  // namespace

extern "C" {

const char *ref_last_error(void) { return g_err.c_str(); }

unsigned ref_hardware_threads(void) { return std::thread::hardware_concurrency(); }

// dtype: 0 = float32, 1 = float64.  Strides in BYTES (pocketfft convention).
int ref_c2c(int dtype, size_t ndim, const size_t *shape, const ptrdiff_t *stride_in,
            const ptrdiff_t *stride_out, size_t naxes, const size_t *axes, int forward,
            const void *in, void *out, double fct, size_t nthreads) {
  return guarded([&] {
    auto sh = mk_shape(shape, ndim);
    auto si = mk_stride(stride_in, ndim), so = mk_stride(stride_out, ndim);
    auto ax = mk_shape(axes, naxes);
    if (dtype == 1)
      pocketfft::c2c<double>(sh, si, so, ax, forward != 0, (const std::complex<double> *)in,
                             (std::complex<double> *)out, fct, nthreads);
    else
      pocketfft::c2c<float>(sh, si, so, ax, forward != 0, (const std::complex<float> *)in,
                            (std::complex<float> *)out, (float)fct, nthreads);
  });
}

// shape = shape of the REAL array (pocketfft_hdronly.h:3320-3349).
int ref_r2c(int dtype, size_t ndim, const size_t *shape, const ptrdiff_t *stride_in,
            const ptrdiff_t *stride_out, size_t naxes, const size_t *axes, int forward,
            const void *in, void *out, double fct, size_t nthreads) {
  return guarded([&] {
    auto sh = mk_shape(shape, ndim);
    auto si = mk_stride(stride_in, ndim), so = mk_stride(stride_out, ndim);
    auto ax = mk_shape(axes, naxes);
    if (dtype == 1)
      pocketfft::r2c<double>(sh, si, so, ax, forward != 0, (const double *)in,
                             (std::complex<double> *)out, fct, nthreads);
    else
      pocketfft::r2c<float>(sh, si, so, ax, forward != 0, (const float *)in,
                            (std::complex<float> *)out, (float)fct, nthreads);
  });
}

// shape = shape of the REAL (output) array (pocketfft_hdronly.h:3352-3390).
int ref_c2r(int dtype, size_t ndim, const size_t *shape, const ptrdiff_t *stride_in,
            const ptrdiff_t *stride_out, size_t naxes, const size_t *axes, int forward,
            const void *in, void *out, double fct, size_t nthreads) {
  return guarded([&] {
    auto sh = mk_shape(shape, ndim);
    auto si = mk_stride(stride_in, ndim), so = mk_stride(stride_out, ndim);
    auto ax = mk_shape(axes, naxes);
    if (dtype == 1)
      pocketfft::c2r<double>(sh, si, so, ax, forward != 0, (const std::complex<double> *)in,
                             (double *)out, fct, nthreads);
    else
      pocketfft::c2r<float>(sh, si, so, ax, forward != 0, (const std::complex<float> *)in,
                            (float *)out, (float)fct, nthreads);
  });
}

// DCT (cosine != 0) / DST types 1..4 (pocketfft_hdronly.h:3284-3318)
int ref_r2r(int cosine, int type, int ortho, int dtype, size_t ndim, const size_t *shape, const ptrdiff_t *stride_in,
            const ptrdiff_t *stride_out, size_t naxes, const size_t *axes, const void *in, void *out, double fct,
            size_t nthreads) {
  return guarded([&] {
    auto sh = mk_shape(shape, ndim);
    auto si = mk_stride(stride_in, ndim), so = mk_stride(stride_out, ndim);
    auto ax = mk_shape(axes, naxes);
    if (dtype == 1) {
      if (cosine) pocketfft::dct<double>(sh, si, so, ax, type, (const double *)in, (double *)out, fct, ortho != 0, nthreads);
      else pocketfft::dst<double>(sh, si, so, ax, type, (const double *)in, (double *)out, fct, ortho != 0, nthreads);
    } else {
      if (cosine) pocketfft::dct<float>(sh, si, so, ax, type, (const float *)in, (float *)out, (float)fct, ortho != 0, nthreads);
      else pocketfft::dst<float>(sh, si, so, ax, type, (const float *)in, (float *)out, (float)fct, ortho != 0, nthreads);
    }
  });
}

// r2r_fftpack (which = 0), r2r_separable_hartley (1), r2r_genuine_hartley (2)  (pocketfft_hdronly.h:3392-3445)
int ref_r2r_real(int which, int dtype, size_t ndim, const size_t *shape, const ptrdiff_t *stride_in,
                 const ptrdiff_t *stride_out, size_t naxes, const size_t *axes, int real2hermitian, int forward,
                 const void *in, void *out, double fct, size_t nthreads) {
  return guarded([&] {
    auto sh = mk_shape(shape, ndim);
    auto si = mk_stride(stride_in, ndim), so = mk_stride(stride_out, ndim);
    auto ax = mk_shape(axes, naxes);
    if (dtype == 1) {
      const double *i = (const double *)in; double *o = (double *)out;
      if (which == 0) pocketfft::r2r_fftpack<double>(sh, si, so, ax, real2hermitian != 0, forward != 0, i, o, fct, nthreads);
      else if (which == 1) pocketfft::r2r_separable_hartley<double>(sh, si, so, ax, i, o, fct, nthreads);
      else pocketfft::r2r_genuine_hartley<double>(sh, si, so, ax, i, o, fct, nthreads);
    } else {
      const float *i = (const float *)in; float *o = (float *)out;
      if (which == 0) pocketfft::r2r_fftpack<float>(sh, si, so, ax, real2hermitian != 0, forward != 0, i, o, (float)fct, nthreads);
      else if (which == 1) pocketfft::r2r_separable_hartley<float>(sh, si, so, ax, i, o, (float)fct, nthreads);
      else pocketfft::r2r_genuine_hartley<float>(sh, si, so, ax, i, o, (float)fct, nthreads);
    }
  });
}

// Row drivers over the reference C engine: `nrows` contiguous rows transformed in
// place, one plan shared by `nthreads` threads.  plan_per_row != 0 reproduces what
// the Nim wrapper does (a fresh plan per call, c_pocketfft/pocketfft.nim:285,300).
static int rows_driver(int real, double *data, size_t nrows, size_t n, int forward, double fct,
                       int nthreads, int plan_per_row) {
  if (n == 0) return -1;
  if (nthreads <= 0) nthreads = (int)std::thread::hardware_concurrency();
  if (nthreads < 1) nthreads = 1;
  if ((size_t)nthreads > nrows) nthreads = nrows ? (int)nrows : 1;
  const size_t row_doubles = real ? n : 2 * n;
  cfft_plan cplan = nullptr;
  rfft_plan rplan = nullptr;
  if (!plan_per_row) {
    if (real) rplan = make_rfft_plan(n); else cplan = make_cfft_plan(n);
    if (!rplan && !cplan) return -1;
  }
  std::vector<int> rc((size_t)nthreads, 0);
  auto work = [&](int t) {
    size_t lo = nrows * (size_t)t / (size_t)nthreads, hi = nrows * (size_t)(t + 1) / (size_t)nthreads;
    for (size_t r = lo; r < hi; ++r) {
      double *row = data + r * row_doubles;
      int e;
      if (real) {
        rfft_plan p = plan_per_row ? make_rfft_plan(n) : rplan;
        e = forward ? rfft_forward(p, row, fct) : rfft_backward(p, row, fct);
        if (plan_per_row) destroy_rfft_plan(p);
      } else {
        cfft_plan p = plan_per_row ? make_cfft_plan(n) : cplan;
        e = forward ? cfft_forward(p, row, fct) : cfft_backward(p, row, fct);
        if (plan_per_row) destroy_cfft_plan(p);
      }
      if (e) rc[(size_t)t] = e;
    }
  };
  if (nthreads == 1) {
    work(0);
  } else {
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; ++t) th.emplace_back(work, t);
    for (auto &x : th) x.join();
  }
  if (rplan) destroy_rfft_plan(rplan);
  if (cplan) destroy_cfft_plan(cplan);
  for (int e : rc)
    if (e) return e;
  return 0;
}

int ref_c_cfft_rows(double *data, size_t nrows, size_t n, int forward, double fct, int nthreads,
                    int plan_per_row) {
  return rows_driver(0, data, nrows, n, forward, fct, nthreads, plan_per_row);
}

int ref_c_rfft_rows(double *data, size_t nrows, size_t n, int forward, double fct, int nthreads,
                    int plan_per_row) {
  return rows_driver(1, data, nrows, n, forward, fct, nthreads, plan_per_row);
}

}  // extern "C"
