// Host planner: factorisation, radix schedule, twiddle / permutation / Bluestein tables,
// LineJob construction and the N-D axis driver.  Pure C++ (no CUDA runtime): tables are
// handed to the device through the TableAlloc hook so the same planner drives the CUDA
// engine (abi.cpp) and the host emulation used by the CPU-only tests (tests/emu).
//
// Reference counterparts: make_cfft_plan / make_rfft_plan / make_fftblue_plan
// (c_pocketfft/pocketfft.c:2066-2153, 1889-1935), cfftp_factorize / comp_twiddle
// (pocketfft.c:953-1031), general_nd / r2c / c2r composition and sanity_check
// (cpp_pocketfft/pocketfft_hdronly.h:446-476, 3011-3050, 3320-3390) and the 16-entry
// plan cache get_plan (pocketfft_hdronly.h:2655-2706).
#pragma once
#include <cstddef>
#include <cstdint>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "fft_types.h"

namespace impulse {

struct TableAlloc {
  virtual ~TableAlloc() {}
  virtual void *upload(const void *host, size_t bytes) = 0;  // returns device pointer or null
  virtual void release(void *dev) = 0;
};

enum Kind : int {
  KIND_C2C = 0, KIND_R2C = 1, KIND_C2R = 2, KIND_DCT = 3, KIND_DST = 4,
  KIND_FFTPACK = 5,       // r2r_fftpack: halfcomplex real transform on every axis (pocketfft_hdronly.h:3392-3403)
  KIND_HARTLEY_SEP = 6,   // r2r_separable_hartley                               (pocketfft_hdronly.h:3405-3415)
  KIND_HARTLEY_GEN = 7,   // r2r_genuine_hartley                                 (pocketfft_hdronly.h:3417-3445)
  KIND_CONV_AXIS = 8,     // complex: IFFT_axis(FFT_axis(x) .* m), one axis (impulse_fft_convolve_axis)
};
enum DType : int { DT_F32 = 0, DT_F64 = 1 };
enum RealLayout : int {
  RL_HERMITIAN = 0, RL_HALFCOMPLEX = 1, RL_FULLSYM = 2,
  // internal (not part of the ABI): halfcomplex with elements 2,4,6,... of the REAL side negated, and the
  // Hartley combination of an r2c result
  RL_HALFCOMPLEX_NEG = 3, RL_HARTLEY = 4,
};

enum Status : int {
  ST_OK = 0,
  ERR_INVALID = -1,      // bad argument (ndim, axis, null pointer, zero length)
  ERR_STRIDE = -2,       // stride mismatch / not a multiple of the element size / misaligned pointer
  ERR_UNSUPPORTED = -3,  // valid request the engine cannot run yet (line too long for shared memory)
  ERR_NOMEM = -4,
  ERR_CUDA = -5,
  ERR_NO_DEVICE = -6,
};

// Complex FFT of logical length L executed in shared memory.
struct Engine1D {
  uint32_t L = 0, n_fft = 0;
  bool blue = false;
  std::vector<uint32_t> radices;  // DIF order
  void *d_tw = nullptr, *d_perm = nullptr, *d_bk = nullptr, *d_bkf = nullptr;
  void *d_bkf_nat = nullptr;  // FFT(b)/n2 in natural order (multi-launch Bluestein), uploaded on demand
  std::vector<double> bkf_nat_host;  // interleaved re,im in double; kept for the on-demand upload
  TableAlloc *owner = nullptr;       // releases the tables when the last reference (cache entry or plan) goes
  ~Engine1D() {
    if (owner) for (void *p : {d_tw, d_perm, d_bk, d_bkf, d_bkf_nat}) if (p) owner->release(p);
  }
};

struct LaunchCfg {
  int threads = 0;
  size_t smem_bytes = 0;
  uint64_t n_tiles = 0;
};

struct LineSpec {
  int kind = KIND_C2C, dtype = DT_F64, layout = RL_HERMITIAN;
  bool forward = true;
  uint32_t N = 0;  // transform length (real length for r2c / c2r)
  uint64_t bdim[kMaxBatchDims] = {1, 1, 1};
  int64_t bs_in[kMaxBatchDims] = {0, 0, 0}, bs_out[kMaxBatchDims] = {0, 0, 0};
  int64_t es_in = 1, es_out = 1;  // element units of the respective side
  uint32_t tw4_n = 0, tw4_dim = 3;  // four-step store twiddle (first of the two launches); dim 3 = none
  uint32_t zero_pad_from = 0;
  const void *mul_tab = nullptr;  // ST_C: multiply by a table indexed line_index + mul_stride*e
  uint32_t mul_stride = 0;
  int r2r_type = 2;    // DCT/DST type 1..4 (pocketfft_hdronly.h:3284-3318)
  bool ortho = false;
  uint64_t umul_mod = 0;  // caller-supplied multiplier fused into the store (pointer set at execute time)
  bool conv_mid = false;  // fused middle pass of an axis convolution (colconv2_kernel); needs tw4_n and umul_mod
  bool col_whole = false;   // plain c2c of a strided power-of-two axis in one launch, the axis in shared memory (colconvw_kernel)
  bool conv_whole = false;  // whole-axis convolution in one launch (colconvw_kernel); needs umul_mod and adjacent lines
  int blue_stage = 0;  // multi-launch Bluestein: 1 = load+chirp+zero-pad to scratch, 2 = scratch+chirp+store
};

enum BufId : int { BUF_IN = 0, BUF_OUT = 1, BUF_TMP = 2, BUF_TMP2 = 3, BUF_TMP3 = 4, BUF_TMP4 = 5 };

struct Step {
  LineJob job;
  LaunchCfg cfg;
  int src = BUF_IN, dst = BUF_OUT;
  int64_t src_off_bytes = 0, dst_off_bytes = 0;  // host-looped outer dims
  bool aux = false;         // not a line job: the elementwise pass `aj` (long real / long Bluestein transforms)
  AuxJob aj{};
  bool combine = false;     // not a line job: the genuine-Hartley fold `cj` (src/dst as usual)
  CombineJob cj{};
  // this step and the NEXT one are the two launches of a four-step split that the device backend may run as ONE
  // persistent kernel with the intermediate in an L2-resident ring (colfuse2_kernel); the emulation and
  // IMPULSE_FFT_NO_COLFUSE run them as the two launches they are
  bool fuse_with_next = false;
  uint32_t fuse_tiles = 0, fuse_g0n = 0;   // tiles = groups of adjacent lines x outer batch index
  bool takes_umul = false;  // this step multiplies its output by the caller's array (impulse_fft_c2c_mul)
  bool takes_fct = false;  // the scaling factor is applied once, in these steps (hdronly.h:3048)
};

struct NdDesc {
  int kind = KIND_C2C, dtype = DT_F64, layout = RL_HERMITIAN;
  bool forward = true;
  std::vector<size_t> shape;            // c2c: array shape; r2c/c2r: shape of the REAL array
  std::vector<ptrdiff_t> stride_in, stride_out;  // bytes
  std::vector<size_t> axes;
  int r2r_type = 2;  // DCT/DST only
  bool ortho = false;
  bool real2hermitian = true;  // KIND_FFTPACK only
  uint64_t umul_mod = 0;  // c2c only: multiply the result by umul[offset % umul_mod]
  bool no_col_whole = false;  // keep strided axes on the column kernels that read segmented input (impulse_fft_cols_from_parts)
};

struct NdPlan {
  NdDesc desc;
  std::vector<Step> steps;
  size_t tmp_bytes = 0;   // c2r N-D intermediate (hdronly.h:3384)
  size_t tmp2_bytes = 0;  // four-step scratch
  size_t tmp2_bytes_fused = 0;  // ... what is left of it when every fusable pair runs fused (device backend)
  size_t tmp3_bytes = 0;  // multi-launch Bluestein work array [lines][n2]
  size_t tmp4_bytes = 0;  // long real transforms: complex work array [lines][L]
  // long even real transforms address the real side as packed complex pairs: the base pointer must be
  // aligned to a complex element
  bool cplx_view_in = false, cplx_view_out = false;
  // byte spans touched relative to the base pointers (for host staging)
  ptrdiff_t in_lo = 0, in_hi = 0, out_lo = 0, out_hi = 0;
  bool empty = false;  // zero-size array: nothing to do
  bool out_dense = true;  // the output span has no gaps between elements
  // the cached device tables the steps point into: held for the life of the plan, so that the table cache can
  // evict entries without invalidating plans
  std::vector<std::shared_ptr<void>> keep;
};

// Least-recently-used map with an access counter, the policy of get_plan (pocketfft_hdronly.h:2655-2706: 16 entries
// per type, `last_access` stamped on every hit, the oldest entry replaced).  Values are shared_ptrs: eviction only
// drops the CACHE's reference — a plan that was built from an entry keeps it alive (NdPlan::keep) until the plan is
// destroyed, so no launch ever sees a freed table.
template <typename K, typename V> struct LruMap {
  struct Ent { V v; uint64_t used = 0; };
  std::map<K, Ent> m;
  uint64_t clock = 0;
  V *find(const K &k) {
    auto it = m.find(k);
    if (it == m.end()) return nullptr;
    it->second.used = ++clock;
    return &it->second.v;
  }
  V &insert(const K &k, V v, size_t cap) {
    while (cap && m.size() >= cap) {
      auto old = m.begin();
      for (auto it = m.begin(); it != m.end(); ++it) if (it->second.used < old->second.used) old = it;
      m.erase(old);
    }
    Ent &e = m[k];
    e.v = std::move(v);
    e.used = ++clock;
    return e.v;
  }
};
using TableRef = std::shared_ptr<void>;

class PlanCache {
 public:
  explicit PlanCache(TableAlloc *alloc);
  ~PlanCache();
  // entries kept per table family and precision-independent key (IMPULSE_FFT_TABLE_CACHE overrides; 0 = unbounded)
  size_t capacity = 64;
  size_t entries() const;   // tables currently referenced by the cache (tests)
  // max shared memory per CTA the engine may use (bytes), set by the backend
  size_t max_smem = 227 * 1024;
  // the fused convolution pass exists as a register kernel only; the host emulator turns it off
  bool allow_conv_fusion = true;
  int status_engine(uint32_t L, int dtype, const Engine1D **out, std::string *err);
  int real_twiddle(uint32_t N, int dtype, const void **out, std::string *err);
  int bluestein_natural_table(uint32_t L, int dtype, const void **out, std::string *err);
  int r2r_twiddle(uint32_t N, int dtype, const void **out, std::string *err);
  int fastblue_tables(uint32_t L, uint32_t M, int dtype, const void **bf, const void **corr, uint32_t *d, std::string *err);
  int fast3_tables(uint32_t N, uint32_t R1, uint32_t R2, uint32_t R3, int dtype, const void **tw1, const void **tw2, std::string *err);
  int four_step_tables(uint32_t N, int dtype, const void **hi, const void **lo, uint32_t *shift, std::string *err);
  int build_line_job(const LineSpec &s, LineJob *job, LaunchCfg *cfg, std::string *err);
  int build_nd(const NdDesc &d, NdPlan *plan, std::string *err);
  static void mark_fusable_pairs(NdPlan *plan);

 private:
  TableRef own(void *dev);          // device table -> reference-counted handle that releases it
  void pin(const TableRef &r);      // remember the table in the plan being built (if any)
  TableAlloc *alloc_;
  mutable std::mutex mu_;
  LruMap<std::pair<uint32_t, int>, std::shared_ptr<Engine1D>> engines_;
  LruMap<std::pair<uint32_t, int>, TableRef> real_tw_;
  LruMap<std::pair<uint32_t, int>, TableRef> r2r_tw_;
  LruMap<std::pair<uint32_t, int>, std::pair<TableRef, TableRef>> tw4_;
  LruMap<std::pair<uint64_t, int>, std::pair<TableRef, TableRef>> f3_;
  LruMap<std::pair<uint64_t, int>, std::pair<TableRef, TableRef>> fb_;
};

// planner utilities exposed for tests
std::vector<uint32_t> choose_radices(uint32_t L);      // empty when a prime factor > kMaxGenericRadix
// smallest 7-smooth n2 >= 2L-1; a power of two instead when that does not fit one CTA (`one_cta_limit` points)
uint32_t bluestein_size(uint32_t L, uint64_t one_cta_limit = ~0ull);
std::vector<uint32_t> dif_positions(uint32_t n, const std::vector<uint32_t> &radices);  // pos_of_k

}  // namespace impulse
