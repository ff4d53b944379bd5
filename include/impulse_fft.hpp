// impulse_fft.hpp — C++ host mirror of Impulse's FFT API over the C ABI of libimpulse_fft_b200.so.
//
// The reference's host language (Nim) has no toolchain in this build environment, and Nim compiles
// through C/C++: this header is the compiled-language host side.  It mirrors, name for name,
//   * the C backend of impulse/fft/c_pocketfft/pocketfft.nim (cited NC:line):
//       NormalizeKind / initNormalize (NC:189-204), FFTPlanReal / FFTPlanComplex RAII (NC:85-114),
//       unpackFFT (NC:126-158), symmetrize / symmTargetSize (NC:160-187), rfft (NC:216-288),
//       fft in place (NC:305-310), rfft_packed (NC:312-319), rfft (NC:321-332), fft (NC:334-348),
//       ifft (NC:350-360);
//   * the C++ backend of impulse/fft/cpp_pocketfft/pocketfft.nim (NX:line):
//       DataDesc (NX:137-142,158-199), FFTDesc (NX:144-149,201-215), apply (NX:235-277).
// Errors: the Nim wrapper raises Exception on a non-zero return (NC:206-214); here std::runtime_error.
// Deviation: normalize/normValue are honoured by every overload.
#pragma once
#include <cmath>
#include <complex>
#include <cstddef>
#include <limits>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

#include "impulse_fft_b200.h"
#include "pocketfft.h"

namespace impulse {

using Complex64 = std::complex<double>;  // Nim's Complex64 = two packed float64 (NC:18-20)

inline bool isOdd(std::size_t i) { return (i & 1) == 1; }

enum NormalizeKind { nkBackward, nkOrtho, nkForward, nkCustom };
struct Normalize { NormalizeKind kind; double value; };

inline Normalize initNormalize(NormalizeKind kind, bool forward, double value, std::size_t length) {
  switch (kind) {
    case nkBackward: return {kind, forward ? 1.0 : 1.0 / double(length)};
    case nkForward: return {kind, forward ? 1.0 / double(length) : 1.0};
    case nkOrtho: return {kind, 1.0 / std::sqrt(double(length))};
    default: return {kind, value};
  }
}

// RAII plans (the `=destroy` hooks of NC:93-108); non-copyable, which the Nim objects are not
struct FFTPlanReal {
  rfft_plan pocket;
  explicit FFTPlanReal(std::size_t length) : pocket(make_rfft_plan(length)) {
    if (!pocket) throw std::runtime_error(std::string("make_rfft_plan failed: ") + impulse_fft_last_error());
  }
  ~FFTPlanReal() { destroy_rfft_plan(pocket); }
  FFTPlanReal(const FFTPlanReal &) = delete;
  FFTPlanReal &operator=(const FFTPlanReal &) = delete;
};
struct FFTPlanComplex {
  cfft_plan pocket;
  explicit FFTPlanComplex(std::size_t length) : pocket(make_cfft_plan(length)) {
    if (!pocket) throw std::runtime_error(std::string("make_cfft_plan failed: ") + impulse_fft_last_error());
  }
  ~FFTPlanComplex() { destroy_cfft_plan(pocket); }
  FFTPlanComplex(const FFTPlanComplex &) = delete;
  FFTPlanComplex &operator=(const FFTPlanComplex &) = delete;
};

// ---- packing helpers -----------------------------------------------------------------------------
inline void unpackFFT(const double *data, Complex64 *outDat, std::size_t inLen) {  // NC:126-150
  std::size_t k = 0;
  Complex64 cmpl(data[0], 0.0);
  outDat[k++] = cmpl;
  for (std::size_t i = 1; i < inLen; ++i) {
    if (isOdd(i)) cmpl = Complex64(data[i], cmpl.imag());
    else { cmpl = Complex64(cmpl.real(), data[i]); outDat[k++] = cmpl; }
  }
  if (!isOdd(inLen)) outDat[k] = Complex64(cmpl.real(), 0.0);
}
inline std::vector<Complex64> unpackFFT(const std::vector<double> &data) {  // NC:152-158
  const std::size_t n = data.size();
  std::vector<Complex64> out(isOdd(n) ? (n + 1) / 2 : (n + 2) / 2);
  unpackFFT(data.data(), out.data(), n);
  return out;
}
inline std::size_t symmTargetSize(const std::vector<double> &data) { return data.size(); }
inline std::size_t symmTargetSize(const std::vector<Complex64> &data) {  // NC:173-183
  return data.size() * 2 - (data.back().imag() == 0.0 ? 2 : 1);
}
inline void fillConjugates(std::vector<Complex64> &res) {  // NC:168-171
  std::size_t k = res.size() - 1;
  for (std::size_t i = 1; i < (res.size() + 1) / 2; ++i) res[k--] = std::conj(res[i]);
}
inline std::vector<Complex64> symmetrize(const std::vector<double> &data) {  // NC:160-187
  std::vector<Complex64> res(symmTargetSize(data));
  unpackFFT(data.data(), res.data(), data.size());
  fillConjugates(res);
  return res;
}
inline std::vector<Complex64> symmetrize(const std::vector<Complex64> &data) {
  std::vector<Complex64> res(symmTargetSize(data));
  for (std::size_t i = 0; i < data.size(); ++i) res[i] = data[i];
  fillConjugates(res);
  return res;
}

// ---- transforms ------------------------------------------------------------------------------------
namespace detail {
inline void check(int err, bool forward) {  // callFFT, NC:206-214
  if (err != 0) throw std::runtime_error(forward ? "Forward FFT calculation failed." : "Backward FFT calculation failed.");
}
}  // namespace detail

// rfft(MemoryView[float], length, forward, ...) — in place, packed (NC:216-288)
inline void rfft(double *data, std::size_t length, bool forward, NormalizeKind normalize = nkBackward,
                 double normValue = std::numeric_limits<double>::infinity()) {
  const Normalize norm = initNormalize(normalize, forward, normValue, length);
  FFTPlanReal plan(length);
  detail::check(forward ? rfft_forward(plan.pocket, data, norm.value) : rfft_backward(plan.pocket, data, norm.value), forward);
}
// fft_impl (NC:290-303)
inline void fft_impl(double *data, std::size_t length, bool forward, NormalizeKind normalize = nkBackward,
                     double normValue = std::numeric_limits<double>::infinity()) {
  rfft(data, length, forward, normalize, normValue);
}
inline void fft_impl(Complex64 *data, std::size_t length, bool forward, NormalizeKind normalize = nkBackward,
                     double normValue = std::numeric_limits<double>::infinity()) {
  const Normalize norm = initNormalize(normalize, forward, normValue, length);
  FFTPlanComplex plan(length);
  double *p = reinterpret_cast<double *>(data);
  detail::check(forward ? cfft_forward(plan.pocket, p, norm.value) : cfft_backward(plan.pocket, p, norm.value), forward);
}
// fft(var data) in place (NC:305-310)
template <typename T>
inline void fft_inplace(std::vector<T> &data, bool forward = true, NormalizeKind normalize = nkBackward,
                        double normValue = std::numeric_limits<double>::infinity()) {
  fft_impl(data.data(), data.size(), forward, normalize, normValue);
}
inline std::vector<double> rfft_packed(const std::vector<double> &data, bool forward = true, NormalizeKind normalize = nkBackward,
                                       double normValue = std::numeric_limits<double>::infinity()) {  // NC:312-319
  std::vector<double> result(data);
  rfft(result.data(), result.size(), forward, normalize, normValue);
  return result;
}
inline std::vector<Complex64> rfft(const std::vector<double> &data, bool forward = true, NormalizeKind normalize = nkBackward,
                                   double normValue = std::numeric_limits<double>::infinity()) {  // NC:321-332
  return unpackFFT(rfft_packed(data, forward, normalize, normValue));
}
inline std::vector<Complex64> fft(const std::vector<double> &data, bool forward = true, NormalizeKind normalize = nkBackward,
                                  double normValue = std::numeric_limits<double>::infinity()) {  // NC:334-345
  return symmetrize(rfft_packed(data, forward, normalize, normValue));
}
inline std::vector<Complex64> fft(const std::vector<Complex64> &data, bool forward = true, NormalizeKind normalize = nkBackward,
                                  double normValue = std::numeric_limits<double>::infinity()) {  // NC:346-348
  std::vector<Complex64> result(data);
  fft_impl(result.data(), result.size(), forward, normalize, normValue);
  return result;
}
template <typename T>
inline std::vector<Complex64> ifft(const std::vector<T> &data, bool backward = true, NormalizeKind normalize = nkBackward,
                                   double normValue = std::numeric_limits<double>::infinity()) {  // NC:350-360
  return fft(data, !backward, normalize, normValue);
}

// ---- C++-backend API: DataDesc / FFTDesc / apply ------------------------------------------------------
template <typename T> struct is_complex : std::false_type {};
template <typename T> struct is_complex<std::complex<T>> : std::true_type { using real = T; };

template <typename T> struct DataDesc {
  std::vector<std::size_t> shape;
  std::vector<std::ptrdiff_t> stride;  // bytes (NX:174,191-194)
  T *buf = nullptr;
  // stride in elements of T (NX:158-176)
  static DataDesc init(T *buffer, const std::vector<std::size_t> &shape, const std::vector<std::ptrdiff_t> &stride) {
    if (shape.size() != stride.size()) throw std::invalid_argument("shape.len == stride.len");
    if (!buffer) throw std::invalid_argument("buffer is nil");
    DataDesc d;
    d.shape = shape;
    for (auto s : stride) d.stride.push_back(s * (std::ptrdiff_t)sizeof(T));
    d.buf = buffer;
    return d;
  }
  // C-contiguous (NX:178-199)
  static DataDesc init(T *buffer, const std::vector<std::size_t> &shape) {
    if (!buffer) throw std::invalid_argument("buffer is nil");
    DataDesc d;
    d.shape = shape;
    d.stride.assign(shape.size(), 0);
    std::ptrdiff_t accum = sizeof(T);
    for (std::size_t i = shape.size(); i-- > 0;) { d.stride[i] = accum; accum *= (std::ptrdiff_t)shape[i]; }
    d.buf = buffer;
    return d;
  }
};

template <typename T> struct FFTDesc {
  std::vector<std::size_t> axes;
  T scalingFactor = 1;
  unsigned nthreads = 1;
  bool forward = true;
  static FFTDesc init(const std::vector<std::size_t> &axes, bool forward, T scalingFactor = 1, unsigned nthreads = 1) {  // NX:201-215
    FFTDesc f;
    f.axes = axes; f.forward = forward; f.scalingFactor = scalingFactor; f.nthreads = nthreads;
    return f;
  }
  template <typename In, typename Out> void apply(DataDesc<Out> &descOut, const DataDesc<In> &descIn, void *stream = nullptr) const {  // NX:235-277
    constexpr int dtype = std::is_same<T, float>::value ? IMPULSE_FFT_F32 : IMPULSE_FFT_F64;
    int rc;
    if constexpr (is_complex<In>::value && is_complex<Out>::value) {
      rc = impulse_fft_c2c(dtype, descIn.shape.size(), descIn.shape.data(), descIn.stride.data(), descOut.stride.data(), axes.size(),
                           axes.data(), forward, descIn.buf, descOut.buf, double(scalingFactor), nthreads, stream);
    } else if constexpr (is_complex<Out>::value) {
      rc = impulse_fft_r2c(dtype, descIn.shape.size(), descIn.shape.data(), descIn.stride.data(), descOut.stride.data(), axes.size(),
                           axes.data(), forward, descIn.buf, descOut.buf, double(scalingFactor), nthreads, stream);
    } else if constexpr (is_complex<In>::value) {
      // the REAL (output) shape, as pocketfft::c2r requires
      rc = impulse_fft_c2r(dtype, descOut.shape.size(), descOut.shape.data(), descIn.stride.data(), descOut.stride.data(), axes.size(),
                           axes.data(), forward, descIn.buf, descOut.buf, double(scalingFactor), nthreads, stream);
    } else {
      static_assert(is_complex<In>::value || is_complex<Out>::value, "Not implemented");
      rc = IMPULSE_FFT_ERR_INVALID;
    }
    if (rc != 0) throw std::runtime_error(std::string("impulse_fft_b200: ") + impulse_fft_last_error());
  }
};


// DCTDesc (NX:151-156, 217-233) and its apply (NX:279-295); `sine` selects pocketfft::dst
template <typename T> struct DCTDesc {
  std::vector<std::size_t> axes;
  int dctType = 2;  // 1..4
  T scalingFactor = 1;
  unsigned nthreads = 1;
  bool ortho = false;
  bool sine = false;
  static DCTDesc init(const std::vector<std::size_t> &axes, int dctType = 2, bool ortho = false, T scalingFactor = 1,
                      unsigned nthreads = 1, bool sine = false) {
    if (dctType < 1 || dctType > 4) throw std::invalid_argument("dctType must be in 1..4");
    DCTDesc d;
    d.axes = axes; d.dctType = dctType; d.ortho = ortho; d.scalingFactor = scalingFactor; d.nthreads = nthreads; d.sine = sine;
    return d;
  }
  void apply(DataDesc<T> &descOut, const DataDesc<T> &descIn, void *stream = nullptr) const {
    constexpr int dtype = std::is_same<T, float>::value ? IMPULSE_FFT_F32 : IMPULSE_FFT_F64;
    auto fn = sine ? impulse_fft_dst : impulse_fft_dct;
    const int rc = fn(dtype, descIn.shape.size(), descIn.shape.data(), descIn.stride.data(), descOut.stride.data(), axes.size(),
                      axes.data(), dctType, descIn.buf, descOut.buf, double(scalingFactor), ortho, nthreads, stream);
    if (rc != 0) throw std::runtime_error(std::string("impulse_fft_b200: ") + impulse_fft_last_error());
  }
};

// r2r_fftpack / r2r_separable_hartley / r2r_genuine_hartley (NX:71-106): the reference imports them but
// no descriptor reaches them; free functions over DataDesc here
template <typename T>
void r2r_fftpack(DataDesc<T> &descOut, const DataDesc<T> &descIn, const std::vector<std::size_t> &axes, bool real2hermitian,
                 bool forward, T fct = 1, unsigned nthreads = 1, void *stream = nullptr) {
  constexpr int dtype = std::is_same<T, float>::value ? IMPULSE_FFT_F32 : IMPULSE_FFT_F64;
  const int rc = impulse_fft_r2r_fftpack(dtype, descIn.shape.size(), descIn.shape.data(), descIn.stride.data(),
                                         descOut.stride.data(), axes.size(), axes.data(), real2hermitian, forward, descIn.buf,
                                         descOut.buf, double(fct), nthreads, stream);
  if (rc != 0) throw std::runtime_error(std::string("impulse_fft_b200: ") + impulse_fft_last_error());
}
template <typename T>
void r2r_hartley(DataDesc<T> &descOut, const DataDesc<T> &descIn, const std::vector<std::size_t> &axes, bool genuine, T fct = 1,
                 unsigned nthreads = 1, void *stream = nullptr) {
  constexpr int dtype = std::is_same<T, float>::value ? IMPULSE_FFT_F32 : IMPULSE_FFT_F64;
  auto fn = genuine ? impulse_fft_r2r_genuine_hartley : impulse_fft_r2r_separable_hartley;
  const int rc = fn(dtype, descIn.shape.size(), descIn.shape.data(), descIn.stride.data(), descOut.stride.data(), axes.size(),
                    axes.data(), descIn.buf, descOut.buf, double(fct), nthreads, stream);
  if (rc != 0) throw std::runtime_error(std::string("impulse_fft_b200: ") + impulse_fft_last_error());
}

}  // namespace impulse
