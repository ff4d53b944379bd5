// Host planner implementation.  See planner.h for the reference counterparts.
#include "planner.h"

#include <algorithm>
#include <cmath>
#include <complex>
#include <cstdlib>
#include <cstring>
#include <functional>

namespace impulse {

namespace {

using ld = long double;
using cld = std::complex<long double>;

// exp(+2*pi*i*m/n) in long double with exact octant reduction (the job of
// sincos_2pibyn, pocketfft.c:36-203 / pocketfft_hdronly.h:277-349).
cld unit_root(uint64_t m, uint64_t n) {
  static const ld PI = 3.141592653589793238462643383279502884L;
  m %= n;
  const uint64_t m8 = 8 * m, oct = m8 / n, rem = m8 - oct * n;
  const ld xa = 2 * PI * (ld)rem / (8 * (ld)n);
  const ld xb = 2 * PI * (ld)(n - rem) / (8 * (ld)n);
  switch (oct) {
    default:
    case 0: return cld(cosl(xa), sinl(xa));
    case 1: return cld(sinl(xb), cosl(xb));
    case 2: return cld(-sinl(xa), cosl(xa));
    case 3: return cld(-cosl(xb), sinl(xb));
    case 4: return cld(-cosl(xa), -sinl(xa));
    case 5: return cld(-sinl(xb), -cosl(xb));
    case 6: return cld(sinl(xa), -cosl(xa));
    case 7: return cld(cosl(xb), -sinl(xb));
  }
}

template <typename T> void *upload_cplx(TableAlloc *a, const std::vector<cld> &v) {
  std::vector<T> h(2 * v.size());
  for (size_t i = 0; i < v.size(); ++i) { h[2 * i] = (T)v[i].real(); h[2 * i + 1] = (T)v[i].imag(); }
  return a->upload(h.data(), h.size() * sizeof(T));
}

void *upload_cplx(TableAlloc *a, const std::vector<cld> &v, int dtype) {
  return dtype == DT_F64 ? upload_cplx<double>(a, v) : upload_cplx<float>(a, v);
}

// plan-time forward FFT in long double (Stockham, generic radix) — used once per Bluestein
// plan to transform the chirp (pocketfft.c:1916-1931 does this with its own cfftp_forward).
void host_fft(std::vector<cld> &x, const std::vector<uint32_t> &radices) {
  const size_t n = x.size();
  std::vector<cld> w(n), y(n);
  for (size_t m = 0; m < n; ++m) w[m] = std::conj(unit_root(m, n));
  size_t l1 = 1;
  cld *p1 = x.data(), *p2 = y.data();
  for (uint32_t ip : radices) {
    const size_t ido = n / (l1 * ip), step = n / ip;
    for (size_t k = 0; k < l1; ++k)
      for (size_t i = 0; i < ido; ++i)
        for (size_t jo = 0; jo < ip; ++jo) {
          cld s = 0;
          for (size_t j = 0; j < ip; ++j) s += p1[i + ido * (j + ip * k)] * w[((j * jo) % ip) * step];
          p2[i + ido * (k + l1 * jo)] = s * w[i * jo * l1];
        }
    std::swap(p1, p2);
    l1 *= ip;
  }
  if (p1 != x.data()) std::copy(p1, p1 + n, x.data());
}

uint32_t pow2floor(uint64_t v) {
  uint32_t r = 1;
  while ((uint64_t)r * 2 <= v) r *= 2;
  return r;
}
uint32_t pow2ceil(uint64_t v) {
  uint32_t r = 1;
  while (r < v) r *= 2;
  return r;
}
uint32_t ilog2(uint32_t v) {
  uint32_t l = 0;
  while ((1u << l) < v) ++l;
  return l;
}
int env_int(const char *name, int dflt) {
  const char *s = std::getenv(name);
  return s && *s ? std::atoi(s) : dflt;
}

}  // namespace

// Radix schedule for the GPU: 9s and 3s, 5s, 7s, other odd primes <= 31, then the power of
// two as 8s with a 4 / 4,4 / 2 tail.  pocketfft factors into 4s, a 2, odd primes
// (pocketfft.c:953-983); only results have to agree, the schedule is free.
std::vector<uint32_t> choose_radices(uint32_t L) {
  std::vector<uint32_t> odd, two;
  uint32_t n = L, e = 0;
  while (n % 2 == 0) { n /= 2; ++e; }
  uint32_t threes = 0;
  while (n % 3 == 0) { n /= 3; ++threes; }
  for (; threes >= 2; threes -= 2) odd.push_back(9);
  if (threes) odd.push_back(3);
  for (uint32_t p = 5; p <= kMaxGenericRadix; p += 2) {
    bool prime = true;
    for (uint32_t q = 3; q * q <= p; q += 2) if (p % q == 0) prime = false;
    if (!prime) continue;
    while (n % p == 0) { n /= p; odd.push_back(p); }
  }
  if (n != 1) return {};  // large prime factor -> Bluestein
  if (e == 1) two = {2};
  else if (e == 2) two = {4};
  else if (e >= 3) {
    uint32_t n8 = e / 3, rem = e % 3;
    if (rem == 1) { n8 -= 1; two.assign(n8, 8); two.push_back(4); two.push_back(4); }
    else { two.assign(n8, 8); if (rem == 2) two.push_back(4); }
  }
  std::vector<uint32_t> r = odd;
  r.insert(r.end(), two.begin(), two.end());
  return r;
}

uint32_t bluestein_size(uint32_t L, uint64_t one_cta_limit) {
  const uint64_t need = 2 * (uint64_t)L - 1;
  uint64_t best = ~0ull;
  for (uint64_t f2 = 1; f2 < 4 * need; f2 *= 2)
    for (uint64_t f3 = f2; f3 < 4 * need; f3 *= 3)
      for (uint64_t f5 = f3; f5 < 4 * need; f5 *= 5)
        for (uint64_t f7 = f5; f7 < 4 * need; f7 *= 7)
          if (f7 >= need && f7 < best) best = f7;
  // A work array that does not fit one CTA is transformed by two column-kernel launches, which exist for
  // power-of-two factors: there the next power of two beats a smaller 7-smooth length (measured 2-3x).
  if (best > one_cta_limit) {
    uint64_t p2 = 1;
    while (p2 < need) p2 *= 2;
    if (p2 <= (1ull << 28)) best = p2;
  }
  return (uint32_t)best;
}

// pos_of_k: shared-memory position of frequency k after the in-place DIF passes.
// With k = k1 + R1*(k2 + R2*(...)) the position is k1*(n/R1) + k2*(n/(R1*R2)) + ...
std::vector<uint32_t> dif_positions(uint32_t n, const std::vector<uint32_t> &radices) {
  std::vector<uint32_t> pos(n);
  for (uint32_t k = 0; k < n; ++k) {
    uint32_t kk = k, span = n, p = 0;
    for (uint32_t r : radices) {
      span /= r;
      p += (kk % r) * span;
      kk /= r;
    }
    pos[k] = p;
  }
  return pos;
}

// the plan under construction on this thread collects references to every table it looks up
static thread_local std::vector<std::shared_ptr<void>> *tl_keep = nullptr;

PlanCache::PlanCache(TableAlloc *alloc) : alloc_(alloc) {
  const int c = env_int("IMPULSE_FFT_TABLE_CACHE", -1);
  if (c >= 0) capacity = (size_t)c;
}
PlanCache::~PlanCache() {}   // entries release their tables when the last reference goes

size_t PlanCache::entries() const {
  std::lock_guard<std::mutex> lk(mu_);
  return engines_.m.size() + real_tw_.m.size() + r2r_tw_.m.size() + tw4_.m.size() + f3_.m.size() + fb_.m.size();
}
TableRef PlanCache::own(void *dev) {
  TableAlloc *a = alloc_;
  return TableRef(dev, [a](void *p) { if (p) a->release(p); });
}
void PlanCache::pin(const TableRef &r) { if (tl_keep && r) tl_keep->push_back(r); }

int PlanCache::status_engine(uint32_t L, int dtype, const Engine1D **out, std::string *err) {
  std::lock_guard<std::mutex> lk(mu_);
  auto key = std::make_pair(L, dtype);
  if (auto *hit = engines_.find(key)) { pin(*hit); *out = hit->get(); return ST_OK; }
  if (L == 0) { *err = "zero-length transform"; return ERR_INVALID; }
  std::shared_ptr<Engine1D> e(new Engine1D);
  e->owner = alloc_;
  e->L = L;
  e->radices = choose_radices(L);
  e->blue = (L > 1 && e->radices.empty());
  e->n_fft = L;
  if (e->blue) {
    e->n_fft = bluestein_size(L, (max_smem - kSmemHeaderBytes) / (dtype == DT_F64 ? 16 : 8) - (dtype == DT_F64 ? 8 : 16));
    e->radices = choose_radices(e->n_fft);
  }
  const uint32_t n = e->n_fft;
  std::vector<cld> tw(n);
  for (uint32_t m = 0; m < n; ++m) tw[m] = std::conj(unit_root(m, n));
  e->d_tw = upload_cplx(alloc_, tw, dtype);
  if (!e->d_tw) { *err = "table upload failed"; return ERR_NOMEM; }
  std::vector<uint32_t> pos = dif_positions(n, e->radices);
  if (!e->blue) {
    e->d_perm = alloc_->upload(pos.data(), pos.size() * sizeof(uint32_t));
    if (!e->d_perm) { *err = "table upload failed"; return ERR_NOMEM; }
  } else {
    // b_k = exp(i*pi*k^2/L): k^2 mod 2L by the recurrence of pocketfft.c:1907-1914
    std::vector<cld> bk(L), bw(n, cld(0, 0));
    uint64_t coeff = 0;
    for (uint32_t m = 0; m < L; ++m) {
      if (m > 0) { coeff += 2 * (uint64_t)m - 1; if (coeff >= 2 * (uint64_t)L) coeff -= 2 * (uint64_t)L; }
      bk[m] = unit_root(coeff, 2 * (uint64_t)L);
    }
    const ld xn2 = 1.0L / (ld)n;
    bw[0] = bk[0] * xn2;
    for (uint32_t m = 1; m < L; ++m) bw[m] = bw[n - m] = bk[m] * xn2;
    host_fft(bw, e->radices);
    std::vector<cld> bkf(n);
    for (uint32_t k = 0; k < n; ++k) bkf[pos[k]] = bw[k];
    e->bkf_nat_host.resize(2 * (size_t)n);
    for (uint32_t k = 0; k < n; ++k) { e->bkf_nat_host[2 * k] = (double)bw[k].real(); e->bkf_nat_host[2 * k + 1] = (double)bw[k].imag(); }
    e->d_bk = upload_cplx(alloc_, bk, dtype);
    e->d_bkf = upload_cplx(alloc_, bkf, dtype);
    if (!e->d_bk || !e->d_bkf) { *err = "table upload failed"; return ERR_NOMEM; }
  }
  *out = e.get();
  pin(e);
  engines_.insert(key, std::move(e), capacity);
  return ST_OK;
}

int PlanCache::bluestein_natural_table(uint32_t L, int dtype, const void **out, std::string *err) {
  const Engine1D *ce = nullptr;
  int rc = status_engine(L, dtype, &ce, err);
  if (rc) return rc;
  std::lock_guard<std::mutex> lk(mu_);
  // (the entry found by status_engine is pinned by the plan under construction, or still cached: const_cast only
  // to fill in the on-demand table under the lock)
  Engine1D *e = const_cast<Engine1D *>(ce);
  if (!e->blue) { *err = "not a Bluestein length"; return ERR_INVALID; }
  if (!e->d_bkf_nat) {
    if (dtype == DT_F64) {
      e->d_bkf_nat = alloc_->upload(e->bkf_nat_host.data(), e->bkf_nat_host.size() * sizeof(double));
    } else {
      std::vector<float> f(e->bkf_nat_host.begin(), e->bkf_nat_host.end());
      e->d_bkf_nat = alloc_->upload(f.data(), f.size() * sizeof(float));
    }
    if (!e->d_bkf_nat) { *err = "table upload failed"; return ERR_NOMEM; }
  }
  *out = e->d_bkf_nat;
  return ST_OK;
}

// Fused Bluestein with a power-of-two work length M that may fall short of 2L-1 by d points:
// bf = FFT_M(wrapped chirp, positive lags kept)/M in natural order; corr[k][j] = b(n-k) - b(M-n+k) with
// n = L-d+j, the products lost to aliasing (outputs k < d only).  See fastblue_kernel.
int PlanCache::fastblue_tables(uint32_t L, uint32_t M, int dtype, const void **bf, const void **corr, uint32_t *dd, std::string *err) {
  const uint32_t need = 2 * L - 1;
  const uint32_t d = need > M ? need - M : 0;
  *dd = d;
  if (d > 16 || L > M) { *err = "work length too short for the fused Bluestein"; return ERR_UNSUPPORTED; }
  std::lock_guard<std::mutex> lk(mu_);
  auto key = std::make_pair(((uint64_t)L << 32) | M, dtype);
  if (auto *hit = fb_.find(key)) { pin(hit->first); pin(hit->second); *bf = hit->first.get(); *corr = hit->second.get(); return ST_OK; }
  std::vector<cld> b(L), bw(M, cld(0, 0));
  uint64_t coeff = 0;
  for (uint32_t m = 0; m < L; ++m) {
    if (m > 0) { coeff += 2 * (uint64_t)m - 1; if (coeff >= 2 * (uint64_t)L) coeff -= 2 * (uint64_t)L; }
    b[m] = unit_root(coeff, 2 * (uint64_t)L);
  }
  const ld xn = 1.0L / (ld)M;
  for (uint32_t m = 1; m < L; ++m) bw[M - m] = b[m] * xn;   // negative lags first ...
  for (uint32_t m = 0; m < L; ++m) bw[m] = b[m] * xn;       // ... positive lags win the aliased residues
  host_fft(bw, choose_radices(M));
  std::vector<cld> cr(16 * 16, cld(0, 0));
  for (uint32_t k = 0; k < d; ++k)
    for (uint32_t j = k; j < d; ++j) {
      const uint32_t n = L - d + j;              // n >= L-d+k
      cr[k * 16 + j] = b[n - k] - b[M - n + k];  // true lag n-k vs the positive lag stored at residue M-(n-k)
    }
  void *dbf = upload_cplx(alloc_, bw, dtype), *dcr = upload_cplx(alloc_, cr, dtype);
  if (!dbf || !dcr) { *err = "table upload failed"; return ERR_NOMEM; }
  auto &ent = fb_.insert(key, std::make_pair(own(dbf), own(dcr)), capacity);
  pin(ent.first); pin(ent.second);
  *bf = dbf; *corr = dcr;
  return ST_OK;
}

// tw1[k1][i1] = W_N^(i1*k1) (i1 < N/R1), tw2[k2][i2] = W_N^(R1*i2*k2) (i2 < R3): fast3_kernel
int PlanCache::fast3_tables(uint32_t N, uint32_t R1, uint32_t R2, uint32_t R3, int dtype, const void **tw1, const void **tw2,
                            std::string *err) {
  std::lock_guard<std::mutex> lk(mu_);
  auto key = std::make_pair(((uint64_t)N << 32) | (R1 << 16) | (R2 << 8) | R3, dtype);
  if (auto *hit = f3_.find(key)) { pin(hit->first); pin(hit->second); *tw1 = hit->first.get(); *tw2 = hit->second.get(); return ST_OK; }
  const uint32_t M1 = N / R1;
  std::vector<cld> a((size_t)R1 * M1), b((size_t)R2 * R3);
  for (uint32_t k1 = 0; k1 < R1; ++k1)
    for (uint32_t i1 = 0; i1 < M1; ++i1) a[(size_t)k1 * M1 + i1] = std::conj(unit_root((uint64_t)i1 * k1, N));
  for (uint32_t k2 = 0; k2 < R2; ++k2)
    for (uint32_t i2 = 0; i2 < R3; ++i2) b[(size_t)k2 * R3 + i2] = std::conj(unit_root((uint64_t)R1 * i2 * k2, N));
  void *da = upload_cplx(alloc_, a, dtype), *db = upload_cplx(alloc_, b, dtype);
  if (!da || !db) { *err = "table upload failed"; return ERR_NOMEM; }
  auto &ent = f3_.insert(key, std::make_pair(own(da), own(db)), capacity);
  pin(ent.first); pin(ent.second);
  *tw1 = da; *tw2 = db;
  return ST_OK;
}

int PlanCache::four_step_tables(uint32_t N, int dtype, const void **hi, const void **lo, uint32_t *shift, std::string *err) {
  std::lock_guard<std::mutex> lk(mu_);
  uint32_t sh = 0;
  while ((1ull << (2 * sh)) < N) ++sh;  // B = 2^sh >= sqrt(N)
  *shift = sh;
  auto key = std::make_pair(N, dtype);
  if (auto *hit = tw4_.find(key)) { pin(hit->first); pin(hit->second); *hi = hit->first.get(); *lo = hit->second.get(); return ST_OK; }
  const uint32_t B = 1u << sh, nhi = (N + B - 1) / B;
  std::vector<cld> vh(nhi), vl(B);
  for (uint32_t j = 0; j < nhi; ++j) vh[j] = std::conj(unit_root((uint64_t)j * B, N));
  for (uint32_t j = 0; j < B; ++j) vl[j] = std::conj(unit_root(j, N));
  void *dh = upload_cplx(alloc_, vh, dtype), *dl = upload_cplx(alloc_, vl, dtype);
  if (!dh || !dl) { *err = "table upload failed"; return ERR_NOMEM; }
  auto &ent = tw4_.insert(key, std::make_pair(own(dh), own(dl)), capacity);
  pin(ent.first); pin(ent.second);
  *hi = dh; *lo = dl;
  return ST_OK;
}

// W_8N^m = exp(-2*pi*i*m/(8N)), m < 2N+2: W_4N^n = tab[2n], W_8N^(2k+1) = tab[2k+1]  (DCT/DST II-IV)
int PlanCache::r2r_twiddle(uint32_t N, int dtype, const void **out, std::string *err) {
  std::lock_guard<std::mutex> lk(mu_);
  auto key = std::make_pair(N, dtype);
  if (auto *hit = r2r_tw_.find(key)) { pin(*hit); *out = hit->get(); return ST_OK; }
  std::vector<cld> w(2 * (size_t)N + 2);
  for (uint32_t m = 0; m < 2 * N + 2; ++m) w[m] = std::conj(unit_root(m, 8 * (uint64_t)N));
  void *d = upload_cplx(alloc_, w, dtype);
  if (!d) { *err = "table upload failed"; return ERR_NOMEM; }
  pin(r2r_tw_.insert(key, own(d), capacity));
  *out = d;
  return ST_OK;
}

int PlanCache::real_twiddle(uint32_t N, int dtype, const void **out, std::string *err) {
  std::lock_guard<std::mutex> lk(mu_);
  auto key = std::make_pair(N, dtype);
  if (auto *hit = real_tw_.find(key)) { pin(*hit); *out = hit->get(); return ST_OK; }
  std::vector<cld> w(N / 2 + 1);
  for (uint32_t k = 0; k <= N / 2; ++k) w[k] = std::conj(unit_root(k, N));
  void *d = upload_cplx(alloc_, w, dtype);
  if (!d) { *err = "table upload failed"; return ERR_NOMEM; }
  pin(real_tw_.insert(key, own(d), capacity));
  *out = d;
  return ST_OK;
}

int PlanCache::build_line_job(const LineSpec &s, LineJob *J, LaunchCfg *cfg, std::string *err) {
  std::memset(J, 0, sizeof(*J));
  const uint32_t N = s.N;
  if (N == 0) { *err = "zero-length transform"; return ERR_INVALID; }
  const bool f64 = s.dtype == DT_F64;
  const bool even = (N % 2) == 0;
  const bool r2r = s.kind == KIND_DCT || s.kind == KIND_DST;
  uint32_t L = N;
  if (r2r) {  // complex embedding length M (see LD_X_* in fft_types.h)
    if (s.r2r_type < 1 || s.r2r_type > 4) { *err = s.kind == KIND_DCT ? "invalid DCT type" : "invalid DST type"; return ERR_INVALID; }
    if (s.r2r_type == 1 && s.kind == KIND_DCT && N < 2) { *err = "DCT-I needs at least two points"; return ERR_INVALID; }
    L = s.r2r_type == 1 ? (s.kind == KIND_DCT ? 2 * (N - 1) : 2 * (N + 1)) : 2 * N;
  } else if (s.kind != KIND_C2C && even) {
    L = N / 2;
  }
  const Engine1D *E = nullptr;
  int rc = status_engine(L, s.dtype, &E, err);
  if (rc) return rc;

  J->n_fft = E->n_fft;
  J->n_seq = L;
  J->n_real = N;
  J->dtype = (uint8_t)s.dtype;
  J->tw = E->d_tw;
  J->perm = E->blue ? nullptr : (const uint32_t *)E->d_perm;
  J->bk = E->d_bk;
  J->bkf = E->d_bkf;
  J->fct = 1.0;
  J->es_in = s.es_in;
  J->es_out = s.es_out;
  uint64_t n_lines = 1;
  for (int d = 0; d < kMaxBatchDims; ++d) {
    J->bdim[d] = s.bdim[d] ? s.bdim[d] : 1;
    J->bs_in[d] = s.bs_in[d];
    J->bs_out[d] = s.bs_out[d];
    n_lines *= J->bdim[d];
  }
  J->n_lines = n_lines;
  J->tw4_dim = s.tw4_dim;
  if (s.tw4_n) {
    rc = four_step_tables(s.tw4_n, s.dtype, &J->tw4_hi, &J->tw4_lo, &J->tw4_shift, err);
    if (rc) return rc;
    J->tw4_n = s.tw4_n;
  }
  J->zero_pad_from = s.zero_pad_from;
  J->mul_tab = s.mul_tab;
  J->mul_stride = s.mul_stride;
  J->umul_mod = s.umul_mod;
  if (s.blue_stage) {  // the line itself (L points) lives in shared memory; the n2-point work array is global
    if (!E->blue) { *err = "internal: Bluestein staging on a direct length"; return ERR_INVALID; }
    J->n_fft = L;
    J->perm = nullptr;
  }

  uint32_t need = J->n_fft;  // shared-memory element slots per line
  uint32_t flags = 0;
  const bool hc = s.layout == RL_HALFCOMPLEX || s.layout == RL_HALFCOMPLEX_NEG;
  std::vector<Phase> pre;
  switch (s.kind) {
    case KIND_C2C:
      J->load_mode = LD_C; J->store_mode = ST_C;
      J->n_load = N; J->n_store = N;
      if (!s.forward) flags |= F_CONJ_SEQ | F_CONJ_OUT;
      break;
    case KIND_R2C:
      if (even) {
        J->load_mode = LD_R_PAIRS; J->n_load = L; J->n_store = L + 1;
        J->store_mode = hc ? ST_HC_EVEN : s.layout == RL_FULLSYM ? ST_R2C_EVEN_SYM : s.layout == RL_HARTLEY ? ST_HARTLEY_EVEN : ST_R2C_EVEN;
        rc = real_twiddle(N, s.dtype, &J->tw_r, err);
        if (rc) return rc;
      } else {
        J->load_mode = LD_R_ZEROIM; J->n_load = N; J->n_store = (N + 1) / 2;
        J->store_mode = hc ? ST_HC_FULL : s.layout == RL_FULLSYM ? ST_HERM_SYM : s.layout == RL_HARTLEY ? ST_HARTLEY_FULL : ST_HERM_HALF;
      }
      if (!s.forward) flags |= F_CONJ_RESULT;
      if (s.layout == RL_HALFCOMPLEX_NEG) flags |= F_NEG_EVEN_IN;
      break;
    case KIND_C2R:
      flags |= F_CONJ_SEQ | F_CONJ_OUT;
      if (s.forward) flags |= F_CONJ_IN;
      if (even) {
        J->load_mode = hc ? LD_HC_EVEN : LD_HERM_EVEN;
        J->n_load = L + 1; J->n_store = L; J->store_mode = ST_R_PAIRS;
        need = std::max(need, L + 1);
        rc = real_twiddle(N, s.dtype, &J->tw_r, err);
        if (rc) return rc;
        Phase p{}; p.op = OP_C2R_PRE_EVEN; pre.push_back(p);
      } else {
        J->load_mode = hc ? LD_HC_FULL : LD_HERM_FULL;
        J->n_load = (N + 1) / 2; J->n_store = N; J->store_mode = ST_R_REALPART;
      }
      if (s.layout == RL_FULLSYM || s.layout == RL_HARTLEY) { *err = "FULLSYM / Hartley are r2c output layouts"; return ERR_INVALID; }
      if (s.layout == RL_HALFCOMPLEX_NEG) flags |= F_NEG_EVEN_OUT;
      break;
    case KIND_DCT:
    case KIND_DST: {
      const bool cosine = s.kind == KIND_DCT;
      const double r2 = 1.4142135623730950488016887242097, ir2 = 0.70710678118654752440084436210485;
      J->n_load = L; J->n_store = N; J->store_mode = ST_X;
      J->x_f0 = J->x_f = J->x_fl = J->x_s = J->x_s0 = J->x_sn = 1.0;
      J->x_shift = 0; J->x_wadd = 0xffffffffu; J->x_im = cosine ? 0 : 1;
      if (s.r2r_type != 1) { rc = r2r_twiddle(N, s.dtype, &J->x_tw, err); if (rc) return rc; }
      switch (s.r2r_type) {
        case 1:
          J->load_mode = cosine ? LD_X_SYM : LD_X_ASYM;
          if (cosine) { if (s.ortho) { J->x_f0 = r2; J->x_s0 = J->x_sn = ir2; } }
          else J->x_shift = 1;
          break;
        case 2:
          J->load_mode = LD_X_ZPAD; J->x_s = 2.0; J->x_wadd = cosine ? 0 : 2; J->x_shift = cosine ? 0 : 1;
          if (s.ortho) J->x_s0 = ir2;
          break;
        case 3:
          J->load_mode = cosine ? LD_X_TW : LD_X_TW_SHIFT; J->x_f = 2.0; J->x_fl = 1.0;
          J->x_f0 = s.ortho ? r2 : 1.0;
          break;
        default:
          J->load_mode = LD_X_TW; J->x_s = 2.0; J->x_wadd = 1;
          break;
      }
    } break;
    default: *err = "bad transform kind"; return ERR_INVALID;
  }

  // strided-axis detection: walk adjacent lines with consecutive threads when the fastest
  // batch dimension is closer in memory than consecutive elements of a line
  auto lines_fast = [&](int64_t es, int64_t bs0, uint64_t b0) {
    if (b0 <= 1) return false;
    return std::llabs(bs0) < std::llabs(es) || es == 0;
  };
  const bool in_lf = lines_fast(s.es_in, s.bs_in[0], J->bdim[0]);
  const bool out_lf = lines_fast(s.es_out, s.bs_out[0], J->bdim[0]);
  if (in_lf) flags |= F_IN_LINES_FAST;
  if (out_lf) flags |= F_OUT_LINES_FAST;

  // phase program
  std::vector<Phase> prog = pre;
  auto add_passes = [&](bool dit) {
    const uint32_t n = E->n_fft;
    std::vector<Phase> ps;
    uint32_t l1 = 1;
    for (uint32_t r : E->radices) {
      Phase p{}; p.op = dit ? OP_PASS_DIT : OP_PASS_DIF; p.radix = (uint8_t)r; p.l1 = l1; p.ido = n / (l1 * r);
      ps.push_back(p);
      l1 *= r;
    }
    if (dit) std::reverse(ps.begin(), ps.end());
    prog.insert(prog.end(), ps.begin(), ps.end());
  };
  if (s.blue_stage == 1) {
    Phase p{}; p.op = OP_BLUE_PRE; prog.push_back(p);
    J->store_mode = ST_C; J->n_store = E->n_fft; J->zero_pad_from = L;
    flags &= ~(uint32_t)(F_CONJ_OUT | F_CONJ_RESULT);
  } else if (s.blue_stage == 2) {
    prog.clear();
    Phase p{}; p.op = OP_MUL_CONJ_BK; prog.push_back(p);
    J->load_mode = LD_C; J->n_load = L;
    flags &= ~(uint32_t)(F_CONJ_IN | F_CONJ_SEQ);
    need = L;
  } else if (E->blue) {
    Phase p{}; p.op = OP_BLUE_PRE; prog.push_back(p);
    add_passes(false);
    p.op = OP_BLUE_MUL; prog.push_back(p);
    add_passes(true);
    p.op = OP_BLUE_POST; prog.push_back(p);
  } else {
    add_passes(false);
  }
  if (prog.size() > (size_t)kMaxPhases) { *err = "phase program too long"; return ERR_UNSUPPORTED; }
  J->nphases = (uint8_t)prog.size();
  for (size_t i = 0; i < prog.size(); ++i) J->ph[i] = prog[i];

  // tile geometry
  const uint32_t S = f64 ? 8 : 16;               // elements per 128-byte shared-memory row
  const size_t esz = f64 ? 16 : 8;
  const size_t budget = max_smem - kSmemHeaderBytes;
  auto pitch_for = [&](uint32_t C) { return ((need + S - 1) / S) * S + (C > 1 ? 1 : 0); };
  uint32_t target = f64 ? 4096 : 8192;
  uint32_t C = pow2floor(std::max<uint64_t>(1, target / need));
  // strided axes: S adjacent lines give full 128-byte segments; small tiles keep many CTAs per SM in
  // flight, which is what hides the load->compute->store serialisation of one CTA (measured: 8 lines
  // beat 64 on the two launches of the 8192-point column split)
  if (in_lf || out_lf) C = S;
  int envC = env_int("IMPULSE_FFT_LINES", 0);
  if (envC > 0) C = pow2floor((uint64_t)envC);
  C = std::min<uint32_t>(C, kMaxLinesPerCta);
  while (C > 1 && (size_t)C * pitch_for(C) * esz > budget) C /= 2;
  if ((size_t)C * pitch_for(C) * esz > budget) {
    *err = "line of " + std::to_string(J->n_fft) + " points does not fit in shared memory";
    return ERR_UNSUPPORTED;
  }
  C = std::min<uint32_t>(C, pow2ceil(n_lines));
  J->log_c = ilog2(C);
  J->pitch = pitch_for(C);
  J->swz_mask = (C < S) ? (S - 1) : 0;
  const size_t smem_bytes = kSmemHeaderBytes + (size_t)C * J->pitch * esz;
  const int tmax = smem_bytes > max_smem / 2 ? kMaxThreadsBig : kMaxThreads;
  int threads = (int)std::min<uint64_t>(tmax, std::max<uint64_t>(64, (((uint64_t)C * J->n_fft / 8) + 31) / 32 * 32));
  int envT = env_int("IMPULSE_FFT_THREADS", 0);
  if (envT > 0) threads = std::min(tmax, (envT + 31) / 32 * 32);
  cfg->threads = threads;
  cfg->smem_bytes = smem_bytes;
  cfg->n_tiles = (n_lines + C - 1) / C;

  // vector access for the packed-real sides (pointer alignment is re-checked at execute time)
  auto all_even = [&](const int64_t *bs) {
    for (int d = 0; d < kMaxBatchDims; ++d) if (J->bdim[d] > 1 && (bs[d] % 2) != 0) return false;
    return true;
  };
  if (J->load_mode == LD_R_PAIRS && s.es_in == 1 && all_even(s.bs_in)) flags |= F_VEC_IN;
  if (J->store_mode == ST_R_PAIRS && s.es_out == 1 && all_even(s.bs_out)) flags |= F_VEC_OUT;
  J->flags = flags;

  // specialised kernels: contiguous complex rows, one batch dimension, headline lengths
  J->fast_id = FAST_NONE;
  if (s.kind == KIND_C2C && !E->blue && !s.tw4_n && !s.zero_pad_from && !s.mul_tab && !s.umul_mod && s.es_in == 1 && s.es_out == 1 && J->bdim[1] == 1 && J->bdim[2] == 1 &&
      !env_int("IMPULSE_FFT_NO_FAST", 0)) {
    switch (N) {
      case 4: J->fast_id = f64 ? FAST2_4_F64 : FAST2_4_F32; break;
      case 8: J->fast_id = f64 ? FAST2_8_F64 : FAST2_8_F32; break;
      case 16: J->fast_id = f64 ? FAST2_16_F64 : FAST2_16_F32; break;
      case 32: J->fast_id = f64 ? FAST2_32_F64 : FAST2_32_F32; break;
      case 64: J->fast_id = f64 ? FAST2_64_F64 : FAST2_64_F32; break;
      case 128: J->fast_id = f64 ? FAST2_128_F64 : FAST2_128_F32; break;
      case 256: J->fast_id = f64 ? FAST2_256_F64 : FAST2_256_F32; break;
      case 512: J->fast_id = f64 ? FAST2_512_F64 : FAST2_512_F32; break;
      case 1024: J->fast_id = f64 ? FAST2_1024_F64 : FAST2_1024_F32; break;
      case 50: if (env_int("IMPULSE_FFT_MORE_SHAPES", 1)) J->fast_id = f64 ? FAST2_50_F64 : FAST2_50_F32; break;
      case 72: if (env_int("IMPULSE_FFT_MORE_SHAPES", 1)) J->fast_id = f64 ? FAST2_72_F64 : FAST2_72_F32; break;
      case 81: if (env_int("IMPULSE_FFT_MORE_SHAPES", 1)) J->fast_id = f64 ? FAST2_81_F64 : FAST2_81_F32; break;
      case 96: if (env_int("IMPULSE_FFT_MORE_SHAPES", 1)) J->fast_id = f64 ? FAST2_96_F64 : FAST2_96_F32; break;
      case 192: if (env_int("IMPULSE_FFT_MORE_SHAPES", 1)) J->fast_id = f64 ? FAST2_192_F64 : FAST2_192_F32; break;
      case 200: if (env_int("IMPULSE_FFT_MORE_SHAPES", 1)) J->fast_id = f64 ? FAST2_200_F64 : FAST2_200_F32; break;
      case 400: if (env_int("IMPULSE_FFT_MORE_SHAPES", 1)) J->fast_id = f64 ? FAST2_400_F64 : FAST2_400_F32; break;
      case 576: if (env_int("IMPULSE_FFT_MORE_SHAPES", 1)) J->fast_id = f64 ? FAST2_576_F64 : FAST2_576_F32; break;
      case 729: if (env_int("IMPULSE_FFT_MORE_SHAPES", 1)) J->fast_id = f64 ? FAST2_729_F64 : FAST2_729_F32; break;
      case 900: if (env_int("IMPULSE_FFT_MORE_SHAPES", 1)) J->fast_id = f64 ? FAST2_900_F64 : FAST2_900_F32; break;
      case 100: if (env_int("IMPULSE_FFT_MORE_SHAPES", 1)) J->fast_id = f64 ? FAST2_100_F64 : FAST2_100_F32; break;
      case 243: if (env_int("IMPULSE_FFT_MORE_SHAPES", 1)) J->fast_id = f64 ? FAST2_243_F64 : FAST2_243_F32; break;
      case 625: if (env_int("IMPULSE_FFT_MORE_SHAPES", 1)) J->fast_id = f64 ? FAST2_625_F64 : FAST2_625_F32; break;
      default: break;
    }
  }
  // column kernels: complex sub-transforms of 64/128/256 points over adjacent strided lines (both sides
  // lines-fastest with unit line stride, whole groups of S lines) — the strided axes of N-D transforms and
  // the two launches of the four-step split
  // (also: contiguous rows in, lines-fastest out — the second launch of the split on contiguous data)
  const bool rows_in = !in_lf && s.es_in == 1 && J->bdim[0] > 1 && s.bs_in[0] >= (int64_t)N && !s.tw4_n && !s.umul_mod &&
                       !env_int("IMPULSE_FFT_NO_COLROWS", 0);
  J->col_in_rows = 0;
  if (s.kind == KIND_C2C && !E->blue && !s.zero_pad_from && !s.mul_tab && !s.blue_stage && (in_lf || rows_in) && out_lf &&
      (rows_in || s.bs_in[0] == 1) && s.bs_out[0] == 1 && J->bdim[0] >= S &&
      (N == 32 || N == 64 || N == 128 || N == 256 || N == 512) &&
      !env_int("IMPULSE_FFT_NO_FAST", 0) && !env_int("IMPULSE_FFT_NO_COLFAST", 0)) {
    switch (N) {
      case 32: J->fast_id = f64 ? COL2_32_F64 : COL2_32_F32; break;
      case 64: J->fast_id = f64 ? COL2_64_F64 : COL2_64_F32; break;
      case 128: J->fast_id = f64 ? COL2_128_F64 : COL2_128_F32; break;
      case 256: J->fast_id = f64 ? COL2_256_F64 : COL2_256_F32; break;
      default: J->fast_id = f64 ? COL2_512_F64 : COL2_512_F32; break;
    }
    J->col_in_rows = (rows_in && !in_lf) ? 1u : 0u;
    if (s.conv_mid) {
      J->fast_id = FAST_NONE;
      if (s.tw4_n && s.umul_mod && J->tw4_dim < 3) {
        if (N == 32) J->fast_id = f64 ? COLCONV_32_F64 : COLCONV_32_F32;
        else if (N == 64) J->fast_id = f64 ? COLCONV_64_F64 : COLCONV_64_F32;
        else if (N == 128) J->fast_id = f64 ? COLCONV_128_F64 : COLCONV_128_F32;
      }
    }
  }
  if (s.conv_mid && (J->fast_id < COLCONV_32_F64 || J->fast_id > COLCONV_128_F32)) {
    *err = "no fused convolution kernel for this shape";
    return ERR_UNSUPPORTED;
  }
  // plain c2c of a strided axis with the whole axis in shared memory: one launch instead of the two of the split
  if (s.col_whole) {
    uint32_t id = FAST_NONE, r2 = 0;
    if (s.kind == KIND_C2C && !s.umul_mod && !s.tw4_n && !s.mul_tab && !s.zero_pad_from && !s.blue_stage && in_lf && out_lf &&
        s.bs_in[0] == 1 && s.bs_out[0] == 1 && n_lines / J->bdim[0] * ((J->bdim[0] + 1) / 2) < (1ull << 31)) {
      // measured against the two-launch split (profiles/r02_ab_colw.txt, r02_ab_convw_tmem.txt): complex128, with the
      // next tile staged in tensor memory, +6...13 % at 1024 points and +4 % at 2048; complex64 -3 % at 1024 and
      // -15...-30 % at 2048 points (one resident CTA per SM and no registers to spare for the staging: load,
      // transform and store do not overlap) — those only under IMPULSE_FFT_COL_WHOLE=2
      const bool all = env_int("IMPULSE_FFT_COL_WHOLE", 1) >= 2;
      if (N == 1024 && (f64 || all)) { id = f64 ? COLW_1024_F64 : COLW_1024_F32; r2 = 8; }
      else if (N == 2048 && (f64 || all)) { id = f64 ? COLW_2048_F64 : COLW_2048_F32; r2 = 16; }
    }
    if (id == FAST_NONE) { *err = "no whole-axis kernel for this shape"; return ERR_UNSUPPORTED; }
    rc = fast3_tables(N, 16, r2, 8, s.dtype, &J->f3_tw1, &J->f3_tw2, err);
    if (rc) return rc;
    J->fast_id = id;
    return ST_OK;
  }
  // whole-axis convolution: FFT -> multiply -> inverse FFT of W adjacent strided lines with the full axis in shared
  // memory (one pass over the data instead of three): power-of-two axes of 512 ... 4096 points
  if (s.conv_whole) {
    uint32_t id = FAST_NONE, r1 = 0, r2 = 0, r3 = 0;
    if (s.kind == KIND_C2C && s.umul_mod && in_lf && out_lf && s.bs_in[0] == 1 && s.bs_out[0] == 1 &&
        n_lines / J->bdim[0] * ((J->bdim[0] + 1) / 2) < (1ull << 31)) {
      if (N == 512) { id = f64 ? COLCONVW_512_F64 : COLCONVW_512_F32; r1 = 8; r2 = 8; r3 = 8; }
      else if (N == 1024) { id = f64 ? COLCONVW_1024_F64 : COLCONVW_1024_F32; r1 = 16; r2 = 8; r3 = 8; }
      // 2048 points leave room for 64-byte runs (measured 7-9 % ahead of the three-launch scheme), 4096 points for
      // 32-byte runs only, which HBM serves at 2.0 TB/s (tools/micro/strided_runs.cu; profiles/r02_ab_convw.txt):
      // slower than the three-launch scheme, so only under IMPULSE_FFT_CONV_WHOLE=2
      else if (N == 2048) { id = f64 ? COLCONVW_2048_F64 : COLCONVW_2048_F32; r1 = 16; r2 = 16; r3 = 8; }
      else if (N == 4096 && env_int("IMPULSE_FFT_CONV_WHOLE", 1) >= 2) { id = f64 ? COLCONVW_4096_F64 : COLCONVW_4096_F32; r1 = 16; r2 = 16; r3 = 16; }
    }
    if (id == FAST_NONE) { *err = "no whole-axis convolution kernel for this shape"; return ERR_UNSUPPORTED; }
    rc = fast3_tables(N, r1, r2, r3, s.dtype, &J->f3_tw1, &J->f3_tw2, err);
    if (rc) return rc;
    J->fast_id = id;
    return ST_OK;
  }
  // fused Bluestein on the register core: complex Bluestein lengths up to 4104 points, and odd real
  // lengths in that range with two rows packed per complex line (Hermitian layout); contiguous rows
  // (float32 measured on the B200: 2.0-3.5x the generic engine, profiles/r02_ab_round2.txt; IMPULSE_FFT_BLUE_F32=0: off)
  if (E->blue && (f64 || env_int("IMPULSE_FFT_BLUE_F32", 1)) && !s.tw4_n && !s.zero_pad_from && !s.mul_tab && !s.umul_mod && !s.blue_stage && s.es_in == 1 && s.es_out == 1 &&
      J->bdim[1] == 1 && J->bdim[2] == 1 && !env_int("IMPULSE_FFT_NO_FAST", 0) && !env_int("IMPULSE_FFT_NO_FASTBLUE", 0)) {
    const bool okc = s.kind == KIND_C2C;
    const bool okr = (s.kind == KIND_R2C || s.kind == KIND_C2R) && !even && s.layout == RL_HERMITIAN;
    uint32_t M = 0, r1 = 16, r2 = 16, r3 = 0, id = FAST_NONE;
    if (2 * L - 1 <= 2048 + 8 && L > 256) { M = 2048; r3 = 8; id = f64 ? FASTBLUE_2048_F64 : FASTBLUE_2048_F32; }
    else if (2 * L - 1 <= 4096 + 8 && L > 256) { M = 4096; r3 = 16; id = f64 ? FASTBLUE_4096_F64 : FASTBLUE_4096_F32; }
    else if (2 * L - 1 <= 8192 + 8 && L > 256) { M = 8192; r3 = 32; id = f64 ? FASTBLUE_8192_F64 : FASTBLUE_8192_F32; }
    if ((okc || okr) && id != FAST_NONE) {
      rc = fast3_tables(M, r1, r2, r3, s.dtype, &J->f3_tw1, &J->f3_tw2, err);
      if (rc) return rc;
      rc = fastblue_tables(L, M, s.dtype, &J->fb_bf, &J->fb_corr, &J->fb_d, err);
      if (rc) return rc;
      J->fast_id = id;
    }
  }
  // three-pass register kernels: c2c of 2048/4096/8192 points, and even-N r2c/c2r (Hermitian layout)
  // whose half-length complex transform is one of those; contiguous rows, one batch dimension
  if (!E->blue && !s.tw4_n && !s.zero_pad_from && !s.mul_tab && !s.umul_mod && !s.blue_stage && s.es_in == 1 && s.es_out == 1 &&
      J->bdim[1] == 1 && J->bdim[2] == 1 && !env_int("IMPULSE_FFT_NO_FAST", 0) && !env_int("IMPULSE_FFT_NO_FAST3", 0)) {
    const bool c2c = s.kind == KIND_C2C;
    const bool r2c = s.kind == KIND_R2C && even && s.layout == RL_HERMITIAN && (J->bdim[0] == 1 || s.bs_in[0] % 2 == 0);
    const bool c2r = s.kind == KIND_C2R && even && s.layout == RL_HERMITIAN && (J->bdim[0] == 1 || s.bs_out[0] % 2 == 0);
    if (c2c || r2c || c2r) {
      uint32_t id = FAST_NONE, r1 = 0, r2 = 0, r3 = 0;
      // c2r of 4096 / 2048 points: shapes with a SHORT first radix, whose pass 1 pairs point n with point N-n in
      // registers (fast3_kernel, PAIR); IMPULSE_FFT_C2R_PAIR=0 restores the shapes shared with c2c / r2c
      if (L == 2048 && c2r && env_int("IMPULSE_FFT_C2R_PAIR", 1)) { id = f64 ? FAST3C_2048_F64 : FAST3C_2048_F32; r1 = 8; r2 = 16; r3 = 16; }
      else if (L == 1024 && c2r && env_int("IMPULSE_FFT_C2R_PAIR", 1)) { id = f64 ? FAST3C_1024_F64 : FAST3C_1024_F32; r1 = 8; r2 = 8; r3 = 16; }
      else if (L == 2048) { id = f64 ? FAST3_2048_F64 : FAST3_2048_F32; r1 = 16; r2 = 16; r3 = 8; }
      else if (L == 4096) { id = f64 ? FAST3_4096_F64 : FAST3_4096_F32; r1 = 16; r2 = 16; r3 = 16; }
      else if (L == 8192) { id = f64 ? FAST3_8192_F64 : FAST3_8192_F32; r1 = 16; r2 = 16; r3 = 32; }
      // r2c of 1000 / 3888 points: shapes with a SHORT last radix, whose pass 3 pairs bin k with bin N-k in
      // registers (fast3_kernel, PAIR); IMPULSE_FFT_R2C_PAIR=0 restores the shapes shared with c2c / c2r
      else if (L == 500 && r2c && env_int("IMPULSE_FFT_R2C_PAIR", 1)) { id = f64 ? FAST3R_500_F64 : FAST3R_500_F32; r1 = 10; r2 = 10; r3 = 5; }
      else if (L == 1944 && r2c && env_int("IMPULSE_FFT_R2C_PAIR", 1)) { id = f64 ? FAST3R_1944_F64 : FAST3R_1944_F32; r1 = 18; r2 = 18; r3 = 6; }
      else if (L == 500) { id = f64 ? FAST3_500_F64 : FAST3_500_F32; r1 = 5; r2 = 10; r3 = 10; }
      else if (L == 1944) { id = f64 ? FAST3_1944_F64 : FAST3_1944_F32; r1 = 6; r2 = 18; r3 = 18; }
      else if (L == 1000) { id = f64 ? FAST3_1000_F64 : FAST3_1000_F32; r1 = 10; r2 = 10; r3 = 10; }
      // more mixed-radix shapes (measured 2.6-3.2x the generic engine, profiles/r02_ab_round2.txt; IMPULSE_FFT_MORE_SHAPES=0: off)
      else if (L == 1536 && f64 && env_int("IMPULSE_FFT_MORE_SHAPES", 1)) { id = FAST3_1536_F64; r1 = 8; r2 = 24; r3 = 8; }
      else if (L == 2000 && f64 && env_int("IMPULSE_FFT_MORE_SHAPES", 1)) { id = FAST3_2000_F64; r1 = 10; r2 = 20; r3 = 10; }
      else if (L == 4000 && f64 && env_int("IMPULSE_FFT_MORE_SHAPES", 1)) { id = FAST3_4000_F64; r1 = 10; r2 = 20; r3 = 20; }
      else if (L == 2187 && env_int("IMPULSE_FFT_MORE_SHAPES", 1)) { id = f64 ? FAST3_2187_F64 : FAST3_2187_F32; r1 = 27; r2 = 9; r3 = 9; }
      else if (L == 3000 && env_int("IMPULSE_FFT_MORE_SHAPES", 1)) { id = f64 ? FAST3_3000_F64 : FAST3_3000_F32; r1 = 10; r2 = 30; r3 = 10; }
      else if (L == 6561 && env_int("IMPULSE_FFT_MORE_SHAPES", 1)) { id = f64 ? FAST3_6561_F64 : FAST3_6561_F32; r1 = 27; r2 = 27; r3 = 9; }
      else if (L == 1536 && !f64 && env_int("IMPULSE_FFT_MORE_SHAPES", 1)) { id = FAST3_1536_F32; r1 = 8; r2 = 24; r3 = 8; }
      else if (L == 2000 && !f64 && env_int("IMPULSE_FFT_MORE_SHAPES", 1)) { id = FAST3_2000_F32; r1 = 10; r2 = 20; r3 = 10; }
      else if (L == 4000 && !f64 && env_int("IMPULSE_FFT_MORE_SHAPES", 1)) { id = FAST3_4000_F32; r1 = 10; r2 = 20; r3 = 20; }
      else if (!c2c && (L == 4 || L == 8) && J->tw_r) {
        // tiny real rows (8 / 16 points): the same warp kernel, 16 rows per warp
        J->fast_id = L == 8 ? (f64 ? FAST2R_8_F64 : FAST2R_8_F32) : (f64 ? FAST2R_4_F64 : FAST2R_4_F32);
      }
      else if (!c2c && (L == 16 || L == 32 || L == 64 || L == 128) && J->tw_r) {
        // short real rows: two-pass warp kernel (fast2r_kernel), tables = the engine's own W_L^m and W_N^k
        J->fast_id = (L == 16 ? FAST2R_16_F64 : L == 32 ? FAST2R_32_F64 : L == 64 ? FAST2R_64_F64 : FAST2R_128_F64) + (f64 ? 0 : 4);
      }
      else if (!c2c && L == 256) { id = f64 ? FAST3R_256_F64 : FAST3R_256_F32; r1 = 8; r2 = 8; r3 = 4; }
      // real rows of 1024 points: 8*8*8 with 16 points per thread pairs both the r2c and the c2r twiddle in registers
      // (measured +9...17 % over the 8-points-per-thread shape; IMPULSE_FFT_F3_512P=0 restores that one)
      else if (!c2c && L == 512 && env_int("IMPULSE_FFT_F3_512P", 1)) { id = f64 ? FAST3P_512_F64 : FAST3P_512_F32; r1 = 8; r2 = 8; r3 = 8; }
      else if (!c2c && L == 512) { id = f64 ? FAST3R_512_F64 : FAST3R_512_F32; r1 = 8; r2 = 8; r3 = 8; }
      else if (!c2c && L == 1024) { id = f64 ? FAST3R_1024_F64 : FAST3R_1024_F32; r1 = 16; r2 = 8; r3 = 8; }
      if (id != FAST_NONE) {
        rc = fast3_tables(L, r1, r2, r3, s.dtype, &J->f3_tw1, &J->f3_tw2, err);
        if (rc) return rc;
        J->fast_id = id;
      }
    }
  }
  return ST_OK;
}

// ---------------------------------------------------------------------------
// N-D driver
// ---------------------------------------------------------------------------
namespace {

struct Dim { uint64_t n; int64_t sin, sout; };

// byte span [lo, hi) touched by an array relative to its base pointer
void span_of(const std::vector<size_t> &shape, const std::vector<ptrdiff_t> &stride, size_t esz,
             ptrdiff_t *lo, ptrdiff_t *hi) {
  ptrdiff_t l = 0, h = 0;
  for (size_t d = 0; d < shape.size(); ++d) {
    ptrdiff_t ext = (ptrdiff_t)(shape[d] - 1) * stride[d];
    if (ext < 0) l += ext; else h += ext;
  }
  *lo = l;
  *hi = h + (ptrdiff_t)esz;
}

}  // namespace

int PlanCache::build_nd(const NdDesc &d, NdPlan *plan, std::string *err) {
  plan->desc = d;
  plan->steps.clear();
  plan->keep.clear();
  struct KeepScope {   // every table looked up while this plan is built is pinned by the plan
    std::vector<std::shared_ptr<void>> *prev;
    explicit KeepScope(std::vector<std::shared_ptr<void>> *k) : prev(tl_keep) { tl_keep = k; }
    ~KeepScope() { tl_keep = prev; }
  } keep_scope(&plan->keep);
  plan->tmp_bytes = 0;
  plan->tmp2_bytes = 0;
  plan->tmp3_bytes = 0;
  const size_t nd = d.shape.size();
  // sanity_check (pocketfft_hdronly.h:446-476)
  if (nd < 1) { *err = "ndim must be >= 1"; return ERR_INVALID; }
  if (d.stride_in.size() != nd || d.stride_out.size() != nd) { *err = "stride dimension mismatch"; return ERR_STRIDE; }
  if (d.axes.empty()) { *err = "no axes given"; return ERR_INVALID; }
  {
    std::vector<int> seen(nd, 0);
    for (size_t ax : d.axes) {
      if (ax >= nd) { *err = "bad axis number"; return ERR_INVALID; }
      if (++seen[ax] > 1) { *err = "axis specified repeatedly"; return ERR_INVALID; }
    }
  }
  if (d.dtype != DT_F32 && d.dtype != DT_F64) { *err = "bad dtype"; return ERR_INVALID; }
  if (d.umul_mod && d.kind != KIND_C2C && d.kind != KIND_CONV_AXIS) { *err = "fused multiply is a c2c option"; return ERR_INVALID; }
  if (d.layout != RL_HERMITIAN && d.axes.size() != 1) { *err = "packed/symmetric real layouts are 1-axis only"; return ERR_INVALID; }
  size_t total = 1;
  for (size_t s : d.shape) total *= s;
  plan->empty = (total == 0);

  const size_t rsz = d.dtype == DT_F64 ? 8 : 4, csz = 2 * rsz;
  const size_t last = d.axes.back();
  std::vector<size_t> cshape = d.shape;  // shape of the complex (half-spectrum) array for real transforms
  if (d.kind == KIND_R2C || d.kind == KIND_C2R) {
    if (d.layout == RL_HERMITIAN) cshape[last] = d.shape[last] / 2 + 1;
    else if (d.layout == RL_FULLSYM) cshape[last] = d.shape[last];
  }
  const bool nd_r2r = d.kind == KIND_DCT || d.kind == KIND_DST || d.kind == KIND_FFTPACK || d.kind == KIND_HARTLEY_SEP ||
                      d.kind == KIND_HARTLEY_GEN;
  const size_t in_esz = (nd_r2r || d.kind == KIND_R2C || (d.kind == KIND_C2R && d.layout == RL_HALFCOMPLEX)) ? rsz : csz;
  const size_t out_esz = (nd_r2r || d.kind == KIND_C2R || (d.kind == KIND_R2C && d.layout == RL_HALFCOMPLEX)) ? rsz : csz;
  const std::vector<size_t> &in_shape = (d.kind == KIND_C2R && d.layout == RL_HERMITIAN) ? cshape : d.shape;
  const std::vector<size_t> &out_shape = (d.kind == KIND_R2C && d.layout != RL_HALFCOMPLEX) ? cshape : d.shape;
  for (size_t i = 0; i < nd; ++i) {
    if (d.stride_in[i] % (ptrdiff_t)in_esz || d.stride_out[i] % (ptrdiff_t)out_esz) {
      *err = "strides must be multiples of the element size";
      return ERR_STRIDE;
    }
  }
  span_of(in_shape, d.stride_in, in_esz, &plan->in_lo, &plan->in_hi);
  span_of(out_shape, d.stride_out, out_esz, &plan->out_lo, &plan->out_hi);
  {
    size_t nout = 1;
    for (size_t s : out_shape) nout *= s;
    plan->out_dense = (size_t)(plan->out_hi - plan->out_lo) == nout * out_esz;
  }
  if (plan->empty) return ST_OK;

  // ---- emit one batched line transform; `dims` are the batch dimensions (element units)
  struct TwDim { bool on = false; uint32_t n = 0; };  // four-step: which dim indexes n2, and N
  auto emit = [&](int kind, int layout, bool forward, uint32_t N, int64_t es_in, int64_t es_out,
                  std::vector<Dim> dims, int tw_key /*index into dims of the line-index dim, -1 = none*/, uint32_t tw4_n,
                  const void *mul_tab, uint32_t mul_stride, int blue_stage,
                  size_t esz_in, size_t esz_out, int src, int dst, int64_t src_base, int64_t dst_base,
                  bool takes_fct, uint64_t umul_mod = 0, bool conv_mid = false, bool conv_whole = false,
                  bool col_whole = false) -> int {
    std::vector<int> key(dims.size());
    for (size_t i = 0; i < dims.size(); ++i) key[i] = (int)i;
    std::vector<size_t> order(dims.size());
    for (size_t i = 0; i < order.size(); ++i) order[i] = i;
    // Lines that are contiguous ROWS on the input side and strided on the output side (second launch of a
    // split): the dimension that is adjacent in the OUTPUT leads, which is what the rows-in column kernel tiles.
    bool rows_job = std::llabs(es_in) == 1 && std::llabs(es_out) > 1 && !dims.empty();
    for (auto &dm : dims) rows_job = rows_job && (uint64_t)std::llabs(dm.sin) >= (uint64_t)N;
    std::sort(order.begin(), order.end(), [&](size_t a, size_t b) {
      const int64_t a1 = std::llabs(rows_job ? dims[a].sout : dims[a].sin), b1 = std::llabs(rows_job ? dims[b].sout : dims[b].sin);
      const int64_t a2 = std::llabs(rows_job ? dims[a].sin : dims[a].sout), b2 = std::llabs(rows_job ? dims[b].sin : dims[b].sout);
      if (a1 != b1) return a1 < b1;
      if (a2 != b2) return a2 < b2;
      return a < b;
    });
    std::vector<Dim> sd;
    std::vector<bool> is_tw;
    for (size_t o : order) { sd.push_back(dims[o]); is_tw.push_back((int)o == tw_key); }
    // merge dims that are contiguous with each other on both sides (never the four-step index dim)
    for (size_t i = 0; i + 1 < sd.size();) {
      if (!is_tw[i] && !is_tw[i + 1] && sd[i + 1].sin == sd[i].sin * (int64_t)sd[i].n &&
          sd[i + 1].sout == sd[i].sout * (int64_t)sd[i].n) {
        sd[i].n *= sd[i + 1].n;
        sd.erase(sd.begin() + (ptrdiff_t)i + 1);
        is_tw.erase(is_tw.begin() + (ptrdiff_t)i + 1);
      } else {
        ++i;
      }
    }
    // the four-step index dim must be one of the three dims the kernel sees
    if (tw_key >= 0) {
      size_t pos = 0;
      while (pos < sd.size() && !is_tw[pos]) ++pos;
      if (pos >= (size_t)kMaxBatchDims) {
        std::swap(sd[pos], sd[kMaxBatchDims - 1]);
        is_tw[pos] = false;
        is_tw[kMaxBatchDims - 1] = true;
      }
    }
    // fused multiply: walk the dim along which the multiplier repeats (the image index of a filter bank)
    // BEFORE the other outer dim, so that a block of multipliers is reused from L2 by every image in turn
    if (umul_mod && sd.size() == 3 && sd[2].sout % (int64_t)umul_mod == 0 && sd[1].sout % (int64_t)umul_mod != 0) {
      std::swap(sd[1], sd[2]);
      const bool t = is_tw[1]; is_tw[1] = is_tw[2]; is_tw[2] = t;
    }
    LineSpec s;
    s.kind = kind; s.dtype = d.dtype; s.layout = layout; s.forward = forward; s.N = N;
    s.es_in = es_in;
    s.es_out = es_out;
    const size_t nk = std::min<size_t>(sd.size(), kMaxBatchDims);
    for (size_t i = 0; i < nk; ++i) {
      s.bdim[i] = sd[i].n; s.bs_in[i] = sd[i].sin; s.bs_out[i] = sd[i].sout;
      if (is_tw[i]) { s.tw4_n = tw4_n; s.tw4_dim = (uint32_t)i; }
    }
    s.mul_tab = mul_tab;
    s.mul_stride = mul_stride;
    s.umul_mod = umul_mod;
    s.conv_mid = conv_mid;
    s.conv_whole = conv_whole;
    s.col_whole = col_whole;
    s.blue_stage = blue_stage;
    s.r2r_type = d.r2r_type;
    s.ortho = d.ortho;
    std::vector<Dim> outer(sd.begin() + (ptrdiff_t)nk, sd.end());
    uint64_t nouter = 1;
    for (auto &o : outer) nouter *= o.n;
    Step proto;
    int rc = build_line_job(s, &proto.job, &proto.cfg, err);
    if (rc) return rc;
    proto.takes_fct = takes_fct;
    proto.takes_umul = umul_mod != 0;
    if (umul_mod && nouter > 1) { *err = "fused multiply supports at most three batch dimensions"; return ERR_UNSUPPORTED; }
    proto.src = src; proto.dst = dst;
    for (uint64_t it = 0; it < nouter; ++it) {
      Step st = proto;
      uint64_t r = it;
      int64_t oi = 0, oo = 0;
      for (auto &o : outer) { uint64_t idx = r % o.n; r /= o.n; oi += (int64_t)idx * o.sin; oo += (int64_t)idx * o.sout; }
      st.src_off_bytes = src_base + oi * (int64_t)esz_in;
      st.dst_off_bytes = dst_base + oo * (int64_t)esz_out;
      plan->steps.push_back(st);
    }
    return ST_OK;
  };

  // byte span of buffer `b` relative to its base pointer (the four-step scratch mirrors the layout of
  // the array it is fed from, so its accesses coalesce exactly like the source's)
  auto span_lo_hi = [&](int b, ptrdiff_t *lo, ptrdiff_t *hi) {
    if (b == BUF_IN) { *lo = plan->in_lo; *hi = plan->in_hi; }
    else if (b == BUF_OUT) { *lo = plan->out_lo; *hi = plan->out_hi; }
    else if (b == BUF_TMP3) { *lo = 0; *hi = (ptrdiff_t)plan->tmp3_bytes; }
    else if (b == BUF_TMP4) { *lo = 0; *hi = (ptrdiff_t)plan->tmp4_bytes; }
    else if (b == BUF_TMP2) { *lo = 0; *hi = (ptrdiff_t)plan->tmp2_bytes; }
    else { *lo = 0; *hi = (ptrdiff_t)plan->tmp_bytes; }
  };

  const size_t csize_g = d.dtype == DT_F64 ? 16 : 8;
  const uint32_t S_g = d.dtype == DT_F64 ? 8 : 16;
  const size_t budget_g = max_smem - kSmemHeaderBytes;
  auto fits_one = [&](uint64_t n) { return (size_t)(n + S_g) * csize_g <= budget_g; };

  // complex line transform of length N; splits N = N1*N2 over two launches when one CTA cannot hold the
  // line, or cannot hold S adjacent lines of a strided axis.  mul_tab (optional) multiplies output k.
  std::function<int(bool, uint32_t, int64_t, int64_t, const std::vector<Dim> &, size_t, size_t, int, int, int64_t, int64_t,
                    bool, const void *, uint64_t)>
      emit_c2c = [&](bool forward, uint32_t N, int64_t es_in, int64_t es_out, const std::vector<Dim> &dims, size_t esz_in,
                     size_t esz_out, int src, int dst, int64_t src_base, int64_t dst_base, bool takes_fct,
                     const void *mul_tab, uint64_t umul_mod) -> int {
    bool split = false;
    if (N >= 64 && !choose_radices(N).empty()) {
      const bool fitsS = (size_t)S_g * (N + S_g + 1) * csize_g <= budget_g;
      uint64_t fastest_in = ~0ull, fastest_out = ~0ull;
      for (auto &dm : dims) {
        fastest_in = std::min<uint64_t>(fastest_in, (uint64_t)std::llabs(dm.sin));
        fastest_out = std::min<uint64_t>(fastest_out, (uint64_t)std::llabs(dm.sout));
      }
      const bool strided = (!dims.empty()) && (fastest_in < (uint64_t)std::llabs(es_in) || fastest_out < (uint64_t)std::llabs(es_out));
      split = !fits_one(N) || (strided && !fitsS) || env_int("IMPULSE_FFT_FORCE_FOURSTEP", 0);
      // fp32 lines of 16384 points (and strided ones of 8192) fit one CTA but have no register kernel: two
      // column-kernel launches (128 x 128, 64 x 128) measure 2-3x faster than the generic single launch
      // strided power-of-two lines of 1024 points and more that would run as ONE generic launch with a single
      // CTA per SM (8 / 16 adjacent lines fill the shared memory): 32 x 32 ... column-kernel launches instead
      if (!split && strided && (N & (N - 1)) == 0 && N >= 1024 && !env_int("IMPULSE_FFT_NO_FAST", 0) &&
          !env_int("IMPULSE_FFT_NO_COLFAST", 0))
        split = true;
      // ... or, for 1024 / 2048 points over adjacent lines, ONE launch with the whole axis in shared memory
      // (colconvw_kernel in its plain-transform mode; the device backend only, like the fused convolution)
      if (split && strided && (N == 1024 || N == 2048) && !mul_tab && !umul_mod && allow_conv_fusion && !d.no_col_whole &&
          env_int("IMPULSE_FFT_COL_WHOLE", 1) && !env_int("IMPULSE_FFT_NO_FAST", 0) && !env_int("IMPULSE_FFT_FORCE_FOURSTEP", 0)) {
        const size_t n0 = plan->steps.size();
        int rcw = emit(KIND_C2C, RL_HERMITIAN, forward, N, es_in, es_out, dims, -1, 0, nullptr, 1, 0, esz_in, esz_out, src, dst,
                       src_base, dst_base, takes_fct, 0, false, false, true);
        if (!rcw) return ST_OK;
        plan->steps.resize(n0);
        err->clear();
      }
      if (!split && d.dtype == DT_F32 && (N == 16384 || (N == 8192 && strided)) && !env_int("IMPULSE_FFT_NO_FAST", 0) &&
          !env_int("IMPULSE_FFT_NO_COLFAST", 0))
        split = true;
    }
    if (!split)
      return emit(KIND_C2C, RL_HERMITIAN, forward, N, es_in, es_out, dims, -1, 0, mul_tab, 1, 0, esz_in, esz_out, src, dst,
                  src_base, dst_base, takes_fct, umul_mod);
    // N1 = largest divisor of N not above sqrt(N); both halves then run in shared memory
    uint32_t N1 = 1;
    for (uint32_t f = 1; (uint64_t)f * f <= N; ++f) if (N % f == 0) N1 = f;
    // Powers of two from 2^19 up: the square-root factors (512 x 1024, 1024 x 1024, ...) have no column
    // kernel, so peel off 128 points and split the remaining 2^k-point transform again — three (or four)
    // passes that all run on the column kernels beat two passes on the generic engine (measured ~2x).
    const bool pow2 = (N & (N - 1)) == 0;
    const bool peel = pow2 && N >= (1u << 19) && !mul_tab && !umul_mod && !env_int("IMPULSE_FFT_NO_PEEL", 0);
    if (peel) N1 = 128;
    const uint32_t N2 = N / N1;
    if (N1 < 2 || (!peel && !fits_one(N2)) || N > (1u << 28)) { *err = "transform length " + std::to_string(N) + " is not supported by the two-kernel split"; return ERR_UNSUPPORTED; }
    ptrdiff_t slo, shi;
    span_lo_hi(src, &slo, &shi);
    plan->tmp2_bytes = std::max<size_t>(plan->tmp2_bytes, (size_t)(shi - slo));
    // step A: for every (line, n2): FFT over n1 of src[(n1*N2 + n2)*es_in], times W_N^(k1*n2),
    //         written to the scratch at the SAME offsets (k1 in place of n1)
    std::vector<Dim> da;
    da.push_back({N2, es_in, es_in});
    for (auto &dm : dims) da.push_back({dm.n, dm.sin, dm.sin});
    int rc = emit(KIND_C2C, RL_HERMITIAN, forward, N1, es_in * (int64_t)N2, es_in * (int64_t)N2, da, 0, N, nullptr, 0, 0,
                  esz_in, esz_in, src, BUF_TMP2, src_base, -(int64_t)slo + src_base, false);
    if (rc) return rc;
    // step B: for every (line, k1): FFT over n2 of scratch[(k1*N2 + n2)*es_in] -> dst[(k1 + N1*k2)*es_out]
    std::vector<Dim> db;
    db.push_back({N1, es_in * (int64_t)N2, es_out});
    for (auto &dm : dims) db.push_back({dm.n, dm.sin, dm.sout});
    if (peel)   // the N2-point rows of the scratch are split again; its first launch runs in place on the scratch
      return emit_c2c(forward, N2, es_in, es_out * (int64_t)N1, db, esz_in, esz_out, BUF_TMP2, dst, -(int64_t)slo + src_base,
                      dst_base, takes_fct, nullptr, 0);
    return emit(KIND_C2C, RL_HERMITIAN, forward, N2, es_in, es_out * (int64_t)N1, db, mul_tab ? 0 : -1, 0, mul_tab, N1, 0,
                esz_in, esz_out, BUF_TMP2, dst, -(int64_t)slo + src_base, dst_base, takes_fct, umul_mod);
  };

  // elementwise pass around a long transform (AuxJob); `dims` = batch dims with user/work strides, the
  // transform axis is appended as the last dimension
  auto emit_aux = [&](int mode, int layout, uint32_t flags, uint32_t N, uint32_t M, uint32_t extent,
                      const std::vector<Dim> &dims /*sin = user side, sout = work side*/, int64_t es_user, int64_t es_work,
                      const void *tab, int src, int dst, bool takes_fct) -> int {
    if (dims.size() + 1 > (size_t)kMaxAuxDims) { *err = "too many dimensions"; return ERR_INVALID; }
    Step st;
    st.aux = true;
    st.src = src; st.dst = dst; st.takes_fct = takes_fct;
    AuxJob &A = st.aj;
    A.mode = mode; A.dtype = d.dtype; A.layout = layout; A.flags = flags; A.N = N; A.M = M;
    A.in = nullptr; A.out = nullptr; A.tab = tab; A.fct = 1.0;
    A.x_load = 0; A.x_tw = nullptr; A.x_f0 = A.x_f = A.x_fl = A.x_s = A.x_s0 = A.x_sn = 1.0; A.x_shift = 0; A.x_wadd = 0xffffffffu; A.x_im = 0;
    A.ndim = (int)dims.size() + 1;
    A.total = extent;
    for (size_t i = 0; i < dims.size(); ++i) {
      A.shape[i] = (uint32_t)dims[i].n; A.s_user[i] = dims[i].sin; A.s_work[i] = dims[i].sout;
      A.total *= dims[i].n;
    }
    A.shape[dims.size()] = extent; A.s_user[dims.size()] = es_user; A.s_work[dims.size()] = es_work;
    plan->steps.push_back(st);
    return ST_OK;
  };

  // complex line transform of any supported length: direct / split (emit_c2c), or Bluestein with the
  // n2-point work array in global memory when it does not fit a CTA
  std::function<int(bool, uint32_t, int64_t, int64_t, const std::vector<Dim> &, size_t, size_t, int, int, bool, uint64_t)>
      c2c_line = [&](bool forward, uint32_t N, int64_t es_in, int64_t es_out, const std::vector<Dim> &dims, size_t esz_in,
                     size_t esz_out, int src, int dst, bool takes_fct, uint64_t umul_mod) -> int {
    // (no tables are built here for lengths that factor: a split length never runs as ONE engine)
    const bool blue = N > 1 && choose_radices(N).empty();
    if (!(blue && (!fits_one(bluestein_size(N)) || env_int("IMPULSE_FFT_FORCE_BIGBLUE", 0))))
      return emit_c2c(forward, N, es_in, es_out, dims, esz_in, esz_out, src, dst, 0, 0, takes_fct, nullptr, umul_mod);
    const Engine1D *E = nullptr;
    int rc = status_engine(N, d.dtype, &E, err);
    if (rc) return rc;
    if (umul_mod) { *err = "fused multiply is available for single- and split-launch complex transforms"; return ERR_UNSUPPORTED; }
    // ---- multi-launch Bluestein:
    //   1. chirp, zero-padded to n2, into the work array [lines][n2]
    //   2. forward FFT(n2) (two launches), output multiplied by FFT(b)/n2
    //   3. backward FFT(n2) (two launches)
    //   4. chirp + store
    const uint32_t n2 = E->n_fft;
    const void *bkf_nat = nullptr;
    rc = bluestein_natural_table(N, d.dtype, &bkf_nat, err);
    if (rc) return rc;
    uint64_t nlines = 1;
    std::vector<Dim> d_in, d_w, d_out;  // batch dims: source->work, work->work, work->destination
    for (auto &dm : dims) {
      d_in.push_back({dm.n, dm.sin, (int64_t)(nlines * n2)});
      d_w.push_back({dm.n, (int64_t)(nlines * n2), (int64_t)(nlines * n2)});
      d_out.push_back({dm.n, dm.sout, (int64_t)(nlines * n2)});   // (user, work) for emit_aux; swapped for emit below
      nlines *= dm.n;
    }
    plan->tmp3_bytes = std::max<size_t>(plan->tmp3_bytes, (size_t)nlines * n2 * csize_g);
    const bool line_fits = fits_one((uint64_t)N + 1) && !env_int("IMPULSE_FFT_FORCE_AUXBLUE", 0);
    if (line_fits) {   // stage in/out through the line kernel (strided sides stay coalesced)
      rc = emit(KIND_C2C, RL_HERMITIAN, forward, N, es_in, 1, d_in, -1, 0, nullptr, 0, 1, esz_in, csize_g, src, BUF_TMP3, 0, 0, false);
    } else {           // the line itself is too long for a CTA: elementwise chirp passes
      rc = emit_aux(AUX_BLUE_PRE, 0, forward ? 0u : (uint32_t)F_CONJ_IN, N, n2, n2, d_in, es_in, 1, E->d_bk, src, BUF_TMP3, false);
    }
    if (rc) return rc;
    rc = emit_c2c(true, n2, 1, 1, d_w, csize_g, csize_g, BUF_TMP3, BUF_TMP3, 0, 0, false, bkf_nat, 0);
    if (rc) return rc;
    rc = emit_c2c(false, n2, 1, 1, d_w, csize_g, csize_g, BUF_TMP3, BUF_TMP3, 0, 0, false, nullptr, 0);
    if (rc) return rc;
    if (line_fits) {
      std::vector<Dim> d_o;
      for (auto &x : d_out) d_o.push_back({x.n, x.sout, x.sin});
      return emit(KIND_C2C, RL_HERMITIAN, forward, N, 1, es_out, d_o, -1, 0, nullptr, 0, 2, csize_g, esz_out, BUF_TMP3, dst, 0, 0, takes_fct);
    }
    return emit_aux(AUX_BLUE_POST, 0, forward ? 0u : (uint32_t)F_CONJ_RESULT, N, n2, N, d_out, es_out, 1, E->d_bk, BUF_TMP3, dst, takes_fct);
  };

  // one batched line transform along `axis`
  auto add_axis = [&](int kind, int layout, bool forward, size_t axis, uint32_t N,
                      const std::vector<size_t> &bshape, const std::vector<ptrdiff_t> &sin, size_t esz_in,
                      const std::vector<ptrdiff_t> &sout, size_t esz_out, int src, int dst, bool takes_fct,
                      uint64_t umul_mod = 0) -> int {
    std::vector<Dim> dims;
    for (size_t i = 0; i < nd; ++i) {
      if (i == axis || bshape[i] == 1) continue;
      dims.push_back({bshape[i], sin[i] / (ptrdiff_t)esz_in, sout[i] / (ptrdiff_t)esz_out});
    }
    const int64_t es_in = sin[axis] / (ptrdiff_t)esz_in, es_out = sout[axis] / (ptrdiff_t)esz_out;
    if (kind == KIND_DCT || kind == KIND_DST) {
      const size_t n0 = plan->steps.size();
      int rc1 = env_int("IMPULSE_FFT_FORCE_BIGR2R", 0) ? (int)ERR_UNSUPPORTED
                    : emit(kind, layout, forward, N, es_in, es_out, dims, -1, 0, nullptr, 0, 0, esz_in, esz_out, src, dst, 0, 0, takes_fct);
      if (rc1 != ERR_UNSUPPORTED) return rc1;
      // ---- the embedding (2N, 2N-2 or 2N+2 complex points, or its Bluestein work array) does not fit one CTA: build it
      // in a work array with one elementwise pass, transform it with c2c_line (any length), extract with a second pass.
      // pocketfft's T_dct1 / T_dcst23 / T_dcst4 take any N (pocketfft_hdronly.h:2424-2648).
      plan->steps.resize(n0);
      err->clear();
      LineSpec ls;                       // reuse the line planner's parameter selection for this (kind, type, ortho)
      ls.kind = kind; ls.dtype = d.dtype; ls.N = 8; ls.r2r_type = d.r2r_type; ls.ortho = d.ortho;
      LineJob pj; LaunchCfg pc;
      int rc2 = build_line_job(ls, &pj, &pc, err);
      if (rc2) return rc2;
      const uint32_t M = d.r2r_type == 1 ? (kind == KIND_DCT ? 2 * (N - 1) : 2 * (N + 1)) : 2 * N;
      if (kind == KIND_DCT && d.r2r_type == 1 && N < 2) { *err = "DCT-I needs at least two points"; return ERR_INVALID; }
      const void *xtw = nullptr;
      if (d.r2r_type != 1) { rc2 = r2r_twiddle(N, d.dtype, &xtw, err); if (rc2) return rc2; }
      uint64_t nlines = 1;
      std::vector<Dim> d_iw, d_ww, d_ow;   // (input, work), (work, work), (output, work)
      for (auto &dm : dims) {
        d_iw.push_back({dm.n, dm.sin, (int64_t)(nlines * M)});
        d_ow.push_back({dm.n, dm.sout, (int64_t)(nlines * M)});
        d_ww.push_back({dm.n, (int64_t)(nlines * M), (int64_t)(nlines * M)});
        nlines *= dm.n;
      }
      plan->tmp4_bytes = std::max<size_t>(plan->tmp4_bytes, (size_t)nlines * M * csize_g);
      auto fill = [&](AuxJob &A) {
        A.x_load = pj.load_mode; A.x_tw = xtw;
        A.x_f0 = pj.x_f0; A.x_f = pj.x_f; A.x_fl = pj.x_fl; A.x_s = pj.x_s; A.x_s0 = pj.x_s0; A.x_sn = pj.x_sn;
        A.x_shift = pj.x_shift; A.x_wadd = pj.x_wadd; A.x_im = pj.x_im;
      };
      rc2 = emit_aux(AUX_X_EMBED, 0, 0, N, M, M, d_iw, es_in, 1, nullptr, src, BUF_TMP4, false);
      if (rc2) return rc2;
      fill(plan->steps.back().aj);
      rc2 = c2c_line(true, M, 1, 1, d_ww, csize_g, csize_g, BUF_TMP4, BUF_TMP4, false, 0);
      if (rc2) return rc2;
      rc2 = emit_aux(AUX_X_EXTRACT, 0, 0, N, M, N, d_ow, es_out, 1, nullptr, BUF_TMP4, dst, takes_fct);
      if (rc2) return rc2;
      fill(plan->steps.back().aj);
      return ST_OK;
    }
    if (kind == KIND_C2C) return c2c_line(forward, N, es_in, es_out, dims, esz_in, esz_out, src, dst, takes_fct, umul_mod);
    if (umul_mod) { *err = "fused multiply is a c2c option"; return ERR_INVALID; }
    const bool even = N % 2 == 0;
    const uint32_t L = even ? N / 2 : N;  // complex length run on the device
    int rc = ST_OK;
    const bool big = !fits_one((uint64_t)L + 1) || env_int("IMPULSE_FFT_FORCE_BIGREAL", 0);
    if (!big) {
      const Engine1D *E = nullptr;
      rc = status_engine(L, d.dtype, &E, err);
      if (rc) return rc;
      if (E->blue && (!fits_one(E->n_fft) || env_int("IMPULSE_FFT_FORCE_BIGBLUE", 0))) {
        // multi-launch Bluestein around a real line that fits a CTA: stage in (load + pre-twiddle + chirp),
        // two FFT(n2) in global memory, stage out (chirp + the transform's own store)
        const uint32_t n2 = E->n_fft;
        const void *bkf_nat = nullptr;
        rc = bluestein_natural_table(L, d.dtype, &bkf_nat, err);
        if (rc) return rc;
        uint64_t nlines = 1;
        std::vector<Dim> d_in, d_w, d_out;
        for (auto &dm : dims) {
          d_in.push_back({dm.n, dm.sin, (int64_t)(nlines * n2)});
          d_w.push_back({dm.n, (int64_t)(nlines * n2), (int64_t)(nlines * n2)});
          d_out.push_back({dm.n, (int64_t)(nlines * n2), dm.sout});
          nlines *= dm.n;
        }
        plan->tmp3_bytes = std::max<size_t>(plan->tmp3_bytes, (size_t)nlines * n2 * csize_g);
        rc = emit(kind, layout, forward, N, es_in, 1, d_in, -1, 0, nullptr, 0, 1, esz_in, csize_g, src, BUF_TMP3, 0, 0, false);
        if (rc) return rc;
        rc = emit_c2c(true, n2, 1, 1, d_w, csize_g, csize_g, BUF_TMP3, BUF_TMP3, 0, 0, false, bkf_nat, 0);
        if (rc) return rc;
        rc = emit_c2c(false, n2, 1, 1, d_w, csize_g, csize_g, BUF_TMP3, BUF_TMP3, 0, 0, false, nullptr, 0);
        if (rc) return rc;
        return emit(kind, layout, forward, N, 1, es_out, d_out, -1, 0, nullptr, 0, 2, csize_g, esz_out, BUF_TMP3, dst, 0, 0, takes_fct);
      }
      return emit(kind, layout, forward, N, es_in, es_out, dims, -1, 0, nullptr, 0, 0, esz_in, esz_out, src, dst, 0, 0, takes_fct);
    }
    // ---- long real lines: the complex transform runs through c2c_line on a work array [lines][L] (or, for
    // even N, directly on the packed view of the real side), with elementwise conversion passes around it
    const bool r2c = kind == KIND_R2C;
    // r2r_fftpack with real2hermitian != forward: elements 2, 4, ... of the REAL side change sign (hdronly.h:3134-3140);
    // the conversion passes carry that, the spectrum side is plain halfcomplex
    const bool neg = layout == RL_HALFCOMPLEX_NEG;
    if (neg) layout = RL_HALFCOMPLEX;
    const uint32_t neg_in = (neg && r2c) ? (uint32_t)F_NEG_EVEN_IN : 0u, neg_out = (neg && !r2c) ? (uint32_t)F_NEG_EVEN_OUT : 0u;
    const void *twr = nullptr;
    if (even) { rc = real_twiddle(N, d.dtype, &twr, err); if (rc) return rc; }
    uint64_t nlines = 1;
    std::vector<Dim> d_uw, d_ww, d_rw, d_view_in, d_view_out;   // (spectrum side, work), (work, work), (real side, work), packed complex views of the real side
    bool view_ok = true;
    for (auto &dm : dims) {
      d_uw.push_back({dm.n, r2c ? dm.sout : dm.sin, (int64_t)(nlines * L)});
      d_rw.push_back({dm.n, r2c ? dm.sin : dm.sout, (int64_t)(nlines * L)});
      d_ww.push_back({dm.n, (int64_t)(nlines * L), (int64_t)(nlines * L)});
      if ((r2c ? dm.sin : dm.sout) % 2) view_ok = false;
      d_view_in.push_back({dm.n, dm.sin / 2, (int64_t)(nlines * L)});
      d_view_out.push_back({dm.n, (int64_t)(nlines * L), dm.sout / 2});
      nlines *= dm.n;
    }
    plan->tmp4_bytes = std::max<size_t>(plan->tmp4_bytes, (size_t)nlines * L * csize_g);
    if (even) {
      // contiguous lines with even row strides: the real side IS a packed complex array (no gather pass).  Anything
      // else — a strided axis, odd row strides, negated elements — is gathered into / scattered from the work array by
      // one more elementwise pass, as general_r2c / general_c2r accept any byte stride (hdronly.h:3125-3250).
      const bool view = (r2c ? es_in : es_out) == 1 && view_ok && !neg && !env_int("IMPULSE_FFT_NO_REAL_VIEW", 0);
      if (r2c) {
        if (view) {
          plan->cplx_view_in = true;
          rc = c2c_line(true, L, 1, 1, d_view_in, csize_g, csize_g, src, BUF_TMP4, false, 0);
        } else {
          rc = emit_aux(AUX_R2C_PACK_EVEN, layout, neg_in, N, L, L, d_rw, es_in, 1, nullptr, src, BUF_TMP4, false);
          if (rc) return rc;
          rc = c2c_line(true, L, 1, 1, d_ww, csize_g, csize_g, BUF_TMP4, BUF_TMP4, false, 0);
        }
        if (rc) return rc;
        return emit_aux(AUX_R2C_POST_EVEN, layout, forward ? 0u : (uint32_t)F_CONJ_RESULT, N, L, L + 1, d_uw, es_out, 1, twr,
                        BUF_TMP4, dst, takes_fct);
      }
      rc = emit_aux(AUX_C2R_PRE_EVEN, layout, forward ? (uint32_t)F_CONJ_IN : 0u, N, L, L / 2 + 1, d_uw, es_in, 1, twr, src,
                    BUF_TMP4, false);
      if (rc) return rc;
      if (view) {
        plan->cplx_view_out = true;
        return c2c_line(false, L, 1, 1, d_view_out, csize_g, csize_g, BUF_TMP4, dst, takes_fct, 0);
      }
      rc = c2c_line(false, L, 1, 1, d_ww, csize_g, csize_g, BUF_TMP4, BUF_TMP4, false, 0);
      if (rc) return rc;
      return emit_aux(AUX_C2R_UNPACK_EVEN, layout, neg_out, N, L, L, d_rw, es_out, 1, nullptr, BUF_TMP4, dst, takes_fct);
    }
    if (r2c) {
      rc = emit_aux(AUX_R2C_PRE_ODD, layout, neg_in, N, L, N, d_rw, es_in, 1, nullptr, src, BUF_TMP4, false);
      if (rc) return rc;
      rc = c2c_line(true, L, 1, 1, d_ww, csize_g, csize_g, BUF_TMP4, BUF_TMP4, false, 0);
      if (rc) return rc;
      return emit_aux(AUX_R2C_POST_ODD, layout, forward ? 0u : (uint32_t)F_CONJ_RESULT, N, L, (N + 1) / 2, d_uw, es_out, 1, nullptr,
                      BUF_TMP4, dst, takes_fct);
    }
    rc = emit_aux(AUX_C2R_PRE_ODD, layout, forward ? (uint32_t)F_CONJ_IN : 0u, N, L, (N + 1) / 2, d_uw, es_in, 1, nullptr, src,
                  BUF_TMP4, false);
    if (rc) return rc;
    rc = c2c_line(false, L, 1, 1, d_ww, csize_g, csize_g, BUF_TMP4, BUF_TMP4, false, 0);
    if (rc) return rc;
    return emit_aux(AUX_C2R_POST_ODD, layout, neg_out, N, L, N, d_rw, es_out, 1, nullptr, BUF_TMP4, dst, takes_fct);
  };

  int rc = ST_OK;
  if (d.kind == KIND_C2C) {
    // general_nd: first axis in -> out with fct, remaining axes in place on out (hdronly.h:3018-3048)
    for (size_t i = 0; i < d.axes.size() && !rc; ++i) {
      const bool first = i == 0;
      const bool lastax = i + 1 == d.axes.size();
      rc = add_axis(KIND_C2C, RL_HERMITIAN, d.forward, d.axes[i], (uint32_t)d.shape[d.axes[i]], d.shape,
                    first ? d.stride_in : d.stride_out, csz, d.stride_out, csz,
                    first ? BUF_IN : BUF_OUT, BUF_OUT, first, lastax ? d.umul_mod : 0);
    }
  } else if (d.kind == KIND_CONV_AXIS) {
    // out = fct * IFFT_axis(FFT_axis(in) .* m[offset % umul_mod]) over one axis of a complex array
    if (d.axes.size() != 1 || !d.umul_mod) { *err = "axis convolution takes one axis and a multiplier"; return ERR_INVALID; }
    const size_t ax = d.axes[0];
    const uint32_t N = (uint32_t)d.shape[ax];
    std::vector<Dim> dims;
    for (size_t i = 0; i < nd; ++i)
      if (i != ax && d.shape[i] != 1) dims.push_back({d.shape[i], d.stride_in[i] / (ptrdiff_t)csz, d.stride_out[i] / (ptrdiff_t)csz});
    const int64_t es = d.stride_out[ax] / (ptrdiff_t)csz;
    bool fused = false;
    // the whole axis in shared memory: one launch, no scratch, in place or out of place (colconvw_kernel)
    if (allow_conv_fusion && es > 1 && d.stride_in[ax] == d.stride_out[ax] && dims.size() <= (size_t)kMaxBatchDims &&
        env_int("IMPULSE_FFT_CONV_WHOLE", 1) && !env_int("IMPULSE_FFT_NO_CONV_FUSION", 0)) {
      const size_t n0 = plan->steps.size();
      rc = emit(KIND_C2C, RL_HERMITIAN, true, N, es, es, dims, -1, 0, nullptr, 0, 0, csz, csz, BUF_IN, BUF_OUT, 0, 0, true,
                d.umul_mod, false, true);
      fused = !rc && plan->steps.size() == n0 + 1;
      if (!fused) { plan->steps.resize(n0); rc = ST_OK; err->clear(); }
    }
    uint32_t N1 = 1;
    for (uint32_t f = 1; (uint64_t)f * f <= N; ++f) if (N % f == 0) N1 = f;
    const uint32_t N2 = N / N1;
    // (positive strides, possibly padded rows: the scratch arrays mirror the layout, padding included)
    if (!fused && allow_conv_fusion && d.stride_in == d.stride_out && plan->out_lo == 0 && es > 1 && N1 >= 32 &&
        !env_int("IMPULSE_FFT_NO_CONV_FUSION", 0)) {
      // three passes through two scratch arrays that mirror the array's own layout (see colconv2_kernel)
      const size_t n0 = plan->steps.size();
      const size_t span = (size_t)(plan->out_hi - plan->out_lo);
      plan->tmp2_bytes = std::max(plan->tmp2_bytes, span);
      plan->tmp3_bytes = std::max(plan->tmp3_bytes, span);
      std::vector<Dim> da, dm, db;
      da.push_back({N2, es, es});                                   // A: FFT over n1 for every n2, times W_N^(k1*n2)
      dm.push_back({N1, es * (int64_t)N2, es});                     // middle: line k1 <- row block k1, -> rows k1' * N1 + k1
      db.push_back({N2, es * (int64_t)N1, es});                     // B': FFT over n2' for every k1'
      for (auto &x : dims) { da.push_back({x.n, x.sout, x.sout}); dm.push_back({x.n, x.sout, x.sout}); db.push_back({x.n, x.sout, x.sout}); }
      rc = emit(KIND_C2C, RL_HERMITIAN, true, N1, es * (int64_t)N2, es * (int64_t)N2, da, 0, N, nullptr, 0, 0, csz, csz, BUF_IN,
                BUF_TMP2, 0, 0, false);
      if (!rc) rc = emit(KIND_C2C, RL_HERMITIAN, true, N2, es, es * (int64_t)N1, dm, 0, N, nullptr, 0, 0, csz, csz, BUF_TMP2,
                         BUF_TMP3, 0, 0, false, d.umul_mod, true);
      if (!rc) rc = emit(KIND_C2C, RL_HERMITIAN, false, N1, es, es * (int64_t)N2, db, -1, 0, nullptr, 0, 0, csz, csz, BUF_TMP3,
                         BUF_OUT, 0, 0, true);
      fused = !rc && plan->steps.size() == n0 + 3;
      if (fused)   // every pass must have landed on a register kernel, otherwise the plain sequence is the better plan
        for (size_t i = n0; i < plan->steps.size(); ++i) fused = fused && plan->steps[i].job.fast_id != FAST_NONE;
      if (!fused) { plan->steps.resize(n0); rc = ST_OK; err->clear(); }
    }
    if (!fused) {
      rc = add_axis(KIND_C2C, RL_HERMITIAN, true, ax, N, d.shape, d.stride_in, csz, d.stride_out, csz, BUF_IN, BUF_OUT, false,
                    d.umul_mod);
      if (!rc) rc = add_axis(KIND_C2C, RL_HERMITIAN, false, ax, N, d.shape, d.stride_out, csz, d.stride_out, csz, BUF_OUT, BUF_OUT, true);
    }
  } else if (d.kind == KIND_FFTPACK) {
    // general_nd with ExecR2R (hdronly.h:3123-3143, 3392-3403): every axis in the given order.  The vendored
    // engine passes `forward` (not `real2hermitian`) to the 1-D plan, so the direction of the real transform
    // follows `forward`, and the mixed combinations only add the sign flips of elements 2,4,6,... —
    // reproduced here as the reference computes it.
    const int kind1 = d.forward ? KIND_R2C : KIND_C2R;
    const int lay = (d.real2hermitian != d.forward) ? RL_HALFCOMPLEX_NEG : RL_HALFCOMPLEX;
    for (size_t i = 0; i < d.axes.size() && !rc; ++i) {
      const bool first = i == 0;
      rc = add_axis(kind1, lay, d.forward, d.axes[i], (uint32_t)d.shape[d.axes[i]], d.shape,
                    first ? d.stride_in : d.stride_out, rsz, d.stride_out, rsz, first ? BUF_IN : BUF_OUT, BUF_OUT, first);
    }
  } else if (d.kind == KIND_HARTLEY_SEP || (d.kind == KIND_HARTLEY_GEN && d.axes.size() == 1)) {
    // general_nd with ExecHartley (hdronly.h:3097-3103): forward real transform, then Re + Im / Re - Im
    for (size_t i = 0; i < d.axes.size() && !rc; ++i) {
      const bool first = i == 0;
      rc = add_axis(KIND_R2C, RL_HARTLEY, true, d.axes[i], (uint32_t)d.shape[d.axes[i]], d.shape,
                    first ? d.stride_in : d.stride_out, rsz, d.stride_out, rsz, first ? BUF_IN : BUF_OUT, BUF_OUT, first);
    }
  } else if (d.kind == KIND_HARTLEY_GEN) {
    // r2c over all axes into a contiguous temporary, then the fold (hdronly.h:3423-3444)
    if (nd > (size_t)kMaxCombineDims) { *err = "too many dimensions"; return ERR_INVALID; }
    std::vector<size_t> hs = d.shape;
    hs[last] = d.shape[last] / 2 + 1;
    std::vector<ptrdiff_t> st(nd);
    st[nd - 1] = (ptrdiff_t)csz;
    for (size_t i = nd - 1; i-- > 0;) st[i] = st[i + 1] * (ptrdiff_t)hs[i + 1];
    size_t nval = 1;
    for (size_t s : hs) nval *= s;
    plan->tmp_bytes = nval * csz;
    rc = add_axis(KIND_R2C, RL_HERMITIAN, true, last, (uint32_t)d.shape[last], d.shape, d.stride_in, rsz, st, csz, BUF_IN,
                  BUF_TMP, true);
    for (size_t i = 0; i + 1 < d.axes.size() && !rc; ++i)
      rc = add_axis(KIND_C2C, RL_HERMITIAN, true, d.axes[i], (uint32_t)hs[d.axes[i]], hs, st, csz, st, csz, BUF_TMP, BUF_TMP,
                    false);
    if (!rc) {
      Step fold;
      fold.combine = true;
      fold.src = BUF_TMP; fold.dst = BUF_OUT;
      CombineJob &C = fold.cj;
      C.in = nullptr; C.out = nullptr;
      C.dtype = d.dtype; C.ndim = (int)nd; C.half_axis = (uint32_t)last; C.total = nval;
      for (size_t i = 0; i < nd; ++i) {
        C.hshape[i] = (uint32_t)hs[i]; C.full[i] = (uint32_t)d.shape[i]; C.rev[i] = 0;
        C.so[i] = d.stride_out[i] / (ptrdiff_t)rsz;
      }
      for (size_t ax : d.axes) C.rev[ax] = 1;
      plan->steps.push_back(fold);
    }
  } else if (nd_r2r) {
    // general_nd with ExecDcst (hdronly.h:3105-3121): every axis is the same 1-D transform, fct once
    for (size_t i = 0; i < d.axes.size() && !rc; ++i) {
      const bool first = i == 0;
      rc = add_axis(d.kind, RL_HERMITIAN, true, d.axes[i], (uint32_t)d.shape[d.axes[i]], d.shape,
                    first ? d.stride_in : d.stride_out, rsz, d.stride_out, rsz, first ? BUF_IN : BUF_OUT, BUF_OUT, first);
    }
  } else if (d.kind == KIND_R2C) {
    // r2c on axes.back(), then c2c in place over the rest on the reduced shape (hdronly.h:3334-3349)
    rc = add_axis(KIND_R2C, d.layout, d.forward, last, (uint32_t)d.shape[last], d.shape, d.stride_in, rsz,
                  d.stride_out, out_esz, BUF_IN, BUF_OUT, true);
    for (size_t i = 0; i + 1 < d.axes.size() && !rc; ++i)
      rc = add_axis(KIND_C2C, RL_HERMITIAN, d.forward, d.axes[i], (uint32_t)cshape[d.axes[i]], cshape,
                    d.stride_out, csz, d.stride_out, csz, BUF_OUT, BUF_OUT, false);
  } else if (d.kind == KIND_C2R) {
    if (d.axes.size() == 1) {
      rc = add_axis(KIND_C2R, d.layout, d.forward, last, (uint32_t)d.shape[last], d.shape, d.stride_in, in_esz,
                    d.stride_out, rsz, BUF_IN, BUF_OUT, true);
    } else {
      // c2c over axes[:-1] into a contiguous temporary, then c2r on axes.back() (hdronly.h:3366-3390)
      std::vector<ptrdiff_t> st(nd);
      st[nd - 1] = (ptrdiff_t)csz;
      for (size_t i = nd - 1; i-- > 0;) st[i] = st[i + 1] * (ptrdiff_t)cshape[i + 1];
      size_t nval = 1;
      for (size_t s : cshape) nval *= s;
      plan->tmp_bytes = nval * csz;
      for (size_t i = 0; i + 1 < d.axes.size() && !rc; ++i) {
        const bool first = i == 0;
        rc = add_axis(KIND_C2C, RL_HERMITIAN, d.forward, d.axes[i], (uint32_t)cshape[d.axes[i]], cshape,
                      first ? d.stride_in : st, csz, st, csz, first ? BUF_IN : BUF_TMP, BUF_TMP, false);
      }
      if (!rc)
        rc = add_axis(KIND_C2R, RL_HERMITIAN, d.forward, last, (uint32_t)d.shape[last], d.shape, st, csz,
                      d.stride_out, rsz, BUF_TMP, BUF_OUT, true);
    }
  } else {
    *err = "bad transform kind";
    return ERR_INVALID;
  }
  if (!rc) mark_fusable_pairs(plan);
  return rc;
}

// Two consecutive steps  A: source -> four-step scratch (column kernel, store twiddle along batch dim 1)
//                        B: scratch -> destination (column kernel)
// over the same groups of adjacent lines are the two launches of ONE split column transform: the device backend runs
// them as a single persistent kernel whose intermediate never leaves L2 (colfuse_device.cuh).
void PlanCache::mark_fusable_pairs(NdPlan *plan) {
  plan->tmp2_bytes_fused = plan->tmp2_bytes;
  bool other_tmp2_user = false;
  std::vector<bool> in_pair(plan->steps.size(), false);
  for (size_t i = 0; i + 1 < plan->steps.size(); ++i) {
    Step &a = plan->steps[i];
    const Step &b = plan->steps[i + 1];
    if (in_pair[i] || a.aux || a.combine || b.aux || b.combine) continue;
    if (a.dst != BUF_TMP2 || b.src != BUF_TMP2 || a.src == BUF_TMP2 || b.dst == BUF_TMP2) continue;
    const LineJob &A = a.job, &B = b.job;
    const bool colA = A.fast_id >= COL2_64_F64 && A.fast_id <= COL2_128_F32 && A.fast_id != COL2_256_F64 && A.fast_id != COL2_256_F32;
    const bool colB = B.fast_id >= COL2_64_F64 && B.fast_id <= COL2_128_F32 && B.fast_id != COL2_256_F64 && B.fast_id != COL2_256_F32;
    if (!colA || !colB || A.dtype != B.dtype) continue;
    if (A.n_fft > B.n_fft) continue;                                   // instantiated pairs: 64x64, 64x128, 128x128
    // measured (profiles/r02_ab_plain.txt, r02_ab_colfuse.txt): 8192-point columns 1.355 -> 1.275 ms fused, but 4096-point
    // columns (64 x 64) 0.335 -> 0.396 ms: short tiles leave too little work per dependency round
    if ((uint64_t)A.n_fft * B.n_fft < (uint64_t)env_int("IMPULSE_FFT_FUSE_MIN_N", 8192)) continue;
    if (!A.tw4_n || A.tw4_dim != 1 || B.tw4_n || A.col_in_rows || B.col_in_rows || A.seg_len || B.seg_len) continue;
    if (A.umul_mod || B.umul_mod || A.mul_tab || B.mul_tab || a.takes_umul || b.takes_umul) continue;
    if (A.bdim[0] != B.bdim[0] || A.bdim[2] != B.bdim[2] || A.bdim[1] != B.n_fft || B.bdim[1] != A.n_fft) continue;
    if (A.bdim[0] < 16 || (A.flags & F_CONJ_SEQ) != (B.flags & F_CONJ_SEQ)) continue;
    const uint64_t g0n = (A.bdim[0] + 15) / 16, tiles = g0n * A.bdim[2];
    if (tiles > 0x00ffffffull) continue;
    a.fuse_with_next = true;
    a.fuse_tiles = (uint32_t)tiles;
    a.fuse_g0n = (uint32_t)g0n;
    in_pair[i] = in_pair[i + 1] = true;
  }
  for (size_t i = 0; i < plan->steps.size(); ++i)
    if (!in_pair[i] && (plan->steps[i].src == BUF_TMP2 || plan->steps[i].dst == BUF_TMP2)) other_tmp2_user = true;
  if (!other_tmp2_user) plan->tmp2_bytes_fused = 0;
}

}  // namespace impulse
