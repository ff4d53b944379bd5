// TEST INFRASTRUCTURE — host emulation of the CUDA engine's phase interpreter.
//
// Compiles impulse_b200/csrc/fft_device.cuh and planner.cpp for the CPU and replaces a CTA
// by a loop over thread ids (one full sweep per phase = one __syncthreads()).  It exists
// because the build container has no GPU: index logic, permutation tables, Bluestein
// plumbing and the N-D driver are checked here before GPU time is spent.  It is NOT a
// product code path: libimpulse_fft_b200.so does not contain it and has no CPU transform.
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../impulse_b200/csrc/fft_device.cuh"
#include "../../impulse_b200/csrc/planner.h"

using namespace impulse;

namespace {
struct HostAlloc : TableAlloc {
  void *upload(const void *h, size_t n) override { void *p = std::malloc(n ? n : 1); if (p) std::memcpy(p, h, n); return p; }
  void release(void *p) override { std::free(p); }
};
HostAlloc g_alloc;
PlanCache *g_cache = nullptr;
std::string g_err;
int g_col_jobs = 0;    // jobs that ran on the column-kernel emulation since the last emu_set_fast_cols()
int g_fast_cols = 0;   // > 0: run column-kernel jobs on the thread-level emulation (emu_col.cpp); value = pipeline groups + 1

template <typename T> void run_job(const LineJob &J, const LaunchCfg &cfg) {
  std::vector<unsigned char> raw(cfg.smem_bytes + 64, 0xCD);  // poison: catches reads of unwritten slots
  unsigned char *base = raw.data();
  base += (16 - ((uintptr_t)base & 15)) & 15;
  int64_t *offs = (int64_t *)base;
  cx<T> *smem = (cx<T> *)(base + kSmemHeaderBytes);
  const uint32_t nthr = (uint32_t)cfg.threads;
  for (uint64_t tile = 0; tile < cfg.n_tiles; ++tile) {
    TileCtx tc = tile_ctx(J, tile);
    for (uint32_t t = 0; t < nthr; ++t) phase_prolog(J, tc, t, offs);
    for (uint32_t t = 0; t < nthr; ++t) phase_load<T>(J, tc, t, nthr, offs, smem);
    for (int p = 0; p < J.nphases; ++p)
      for (uint32_t t = 0; t < nthr; ++t) phase_mid<T>(J, J.ph[p], t, nthr, smem);
    for (uint32_t t = 0; t < nthr; ++t) phase_store<T>(J, tc, t, nthr, offs, smem);
  }
}
}  // namespace

// emu_col.cpp: 0 = ran on an emulated column kernel, 1 = not a column-kernel job, < 0 = error
int emu_run_col_job(const LineJob &J, unsigned pipe_groups);

static void ensure_cache() {
  if (!g_cache) g_cache = new PlanCache(&g_alloc);
  // the fused convolution pass exists as a register kernel only: plan it only when those kernels are emulated
  g_cache->allow_conv_fusion = g_fast_cols > 0;
}

extern "C" {

const char *emu_last_error() { return g_err.c_str(); }

// 0: every job on the generic engine's phase emulation (default); g >= 1: column-kernel jobs on the thread-level
// emulation of colfast2 / colpipe2 / colconv2, with g - 1 pipeline groups for colpipe2 (the product's default is 2)
void emu_set_fast_cols(int g) { g_fast_cols = g; g_col_jobs = 0; }
int emu_col_job_count() { return g_col_jobs; }

// mirrors impulse_fft_nd() of the product ABI, on host memory
static int emu_nd_impl(int kind, int dtype, int layout, size_t ndim, const size_t *shape, const ptrdiff_t *stride_in,
                       const ptrdiff_t *stride_out, size_t naxes, const size_t *axes, int forward, const void *in,
                       void *out, double fct, int r2r_type, int ortho, const void *umul = nullptr, size_t umul_mod = 0,
                       int real2hermitian = 1);

int emu_nd(int kind, int dtype, int layout, size_t ndim, const size_t *shape, const ptrdiff_t *stride_in,
           const ptrdiff_t *stride_out, size_t naxes, const size_t *axes, int forward, const void *in,
           void *out, double fct) {
  return emu_nd_impl(kind, dtype, layout, ndim, shape, stride_in, stride_out, naxes, axes, forward, in, out, fct, 2, 0);
}

// mirrors impulse_fft_c2c_mul: complex transform with a pointwise multiply fused into the last store
int emu_c2c_mul(int dtype, size_t ndim, const size_t *shape, const ptrdiff_t *stride_in, const ptrdiff_t *stride_out,
                size_t naxes, const size_t *axes, int forward, const void *in, void *out, double fct, const void *mul,
                size_t mul_elems) {
  return emu_nd_impl(KIND_C2C, dtype, RL_HERMITIAN, ndim, shape, stride_in, stride_out, naxes, axes, forward, in, out, fct,
                     2, 0, mul, mul_elems);
}

// mirrors impulse_fft_convolve_axis (the plain FFT -> multiply -> inverse FFT plan; the fused kernel is GPU-only)
int emu_convolve_axis(int dtype, size_t ndim, const size_t *shape, const ptrdiff_t *stride_in, const ptrdiff_t *stride_out,
                      size_t axis, const void *in, void *out, double fct, const void *mul, size_t mul_elems) {
  const size_t axes[1] = {axis};
  return emu_nd_impl(KIND_CONV_AXIS, dtype, RL_HERMITIAN, ndim, shape, stride_in, stride_out, 1, axes, 1, in, out, fct, 2, 0, mul,
                     mul_elems);
}

// mirrors impulse_fft_r2r_fftpack (which = 0), _separable_hartley (1), _genuine_hartley (2)
int emu_r2r_real(int which, int dtype, size_t ndim, const size_t *shape, const ptrdiff_t *stride_in, const ptrdiff_t *stride_out,
                 size_t naxes, const size_t *axes, int real2hermitian, int forward, const void *in, void *out, double fct) {
  const int kind = which == 0 ? KIND_FFTPACK : which == 1 ? KIND_HARTLEY_SEP : KIND_HARTLEY_GEN;
  return emu_nd_impl(kind, dtype, RL_HERMITIAN, ndim, shape, stride_in, stride_out, naxes, axes, which == 0 ? forward : 1, in,
                     out, fct, 2, 0, nullptr, 0, real2hermitian);
}

// DCT (cosine != 0) / DST of type 1..4, mirrors impulse_fft_dct / impulse_fft_dst
int emu_r2r(int cosine, int type, int ortho, int dtype, size_t ndim, const size_t *shape, const ptrdiff_t *stride_in,
            const ptrdiff_t *stride_out, size_t naxes, const size_t *axes, const void *in, void *out, double fct) {
  return emu_nd_impl(cosine ? KIND_DCT : KIND_DST, dtype, RL_HERMITIAN, ndim, shape, stride_in, stride_out, naxes, axes, 1,
                     in, out, fct, type, ortho);
}

static int emu_nd_impl(int kind, int dtype, int layout, size_t ndim, const size_t *shape, const ptrdiff_t *stride_in,
                       const ptrdiff_t *stride_out, size_t naxes, const size_t *axes, int forward, const void *in,
                       void *out, double fct, int r2r_type, int ortho, const void *umul, size_t umul_mod,
                       int real2hermitian) {
  ensure_cache();
  NdDesc d;
  d.r2r_type = r2r_type; d.ortho = ortho != 0;
  d.umul_mod = umul_mod;
  d.real2hermitian = real2hermitian != 0;
  d.kind = kind; d.dtype = dtype; d.layout = layout; d.forward = forward != 0;
  d.shape.assign(shape, shape + ndim);
  d.stride_in.assign(stride_in, stride_in + ndim);
  d.stride_out.assign(stride_out, stride_out + ndim);
  d.axes.assign(axes, axes + naxes);
  NdPlan plan;
  int rc = g_cache->build_nd(d, &plan, &g_err);
  if (rc) return rc;
  if (plan.empty) return 0;
  std::vector<unsigned char> tmp(plan.tmp_bytes + 16), tmp2(plan.tmp2_bytes + 16, 0xCD), tmp3(plan.tmp3_bytes + 16, 0xCD),
      tmp4(plan.tmp4_bytes + 16, 0xCD);
  for (Step &st : plan.steps) {
    const unsigned char *bufs_in[6] = {(const unsigned char *)in, (const unsigned char *)out, tmp.data(), tmp2.data(), tmp3.data(), tmp4.data()};
    unsigned char *bufs_out[6] = {nullptr, (unsigned char *)out, tmp.data(), tmp2.data(), tmp3.data(), tmp4.data()};
    if (st.aux) {
      AuxJob aj = st.aj;
      aj.in = bufs_in[st.src] + st.src_off_bytes;
      aj.out = bufs_out[st.dst] + st.dst_off_bytes;
      aj.fct = st.takes_fct ? fct : 1.0;
      for (uint64_t g = 0; g < aj.total; ++g) {
        if (dtype == DT_F64) aux_one<double>(aj, g); else aux_one<float>(aj, g);
      }
      continue;
    }
    if (st.combine) {
      CombineJob cj = st.cj;
      cj.in = bufs_in[st.src] + st.src_off_bytes;
      cj.out = bufs_out[st.dst] + st.dst_off_bytes;
      for (uint64_t g = 0; g < cj.total; ++g) {
        if (dtype == DT_F64) hartley_combine_one<double>(cj, g); else hartley_combine_one<float>(cj, g);
      }
      continue;
    }
    st.job.in = bufs_in[st.src] + st.src_off_bytes;
    st.job.out = bufs_out[st.dst] + st.dst_off_bytes;
    st.job.fct = st.takes_fct ? fct : 1.0;
    if (st.takes_umul) st.job.umul = umul; else st.job.umul_mod = 0;
    if (g_fast_cols > 0) {
      const int cr = emu_run_col_job(st.job, (unsigned)(g_fast_cols - 1));
      if (cr < 0) { g_err = "column-kernel emulation failed"; return -100 + cr; }
      if (cr == 0) { ++g_col_jobs; continue; }
    }
    if (dtype == DT_F64) run_job<double>(st.job, st.cfg); else run_job<float>(st.job, st.cfg);
  }
  return 0;
}

// number of kernel launches the N-D plan would issue (tests: four-step split, host-looped dims)
int emu_nd_steps(int kind, int dtype, int layout, size_t ndim, const size_t *shape, const ptrdiff_t *stride_in,
                 const ptrdiff_t *stride_out, size_t naxes, const size_t *axes, int forward) {
  ensure_cache();
  NdDesc d;
  d.kind = kind; d.dtype = dtype; d.layout = layout; d.forward = forward != 0;
  d.shape.assign(shape, shape + ndim);
  d.stride_in.assign(stride_in, stride_in + ndim);
  d.stride_out.assign(stride_out, stride_out + ndim);
  d.axes.assign(axes, axes + naxes);
  NdPlan plan;
  int rc = g_cache->build_nd(d, &plan, &g_err);
  return rc ? rc : (int)plan.steps.size();
}

// which specialised kernel (FastId, 0 = generic engine) the planner picks for each step of an N-D transform
int emu_nd_fast_ids(int kind, int dtype, int layout, size_t ndim, const size_t *shape, const ptrdiff_t *stride_in,
                    const ptrdiff_t *stride_out, size_t naxes, const size_t *axes, int forward, uint32_t *ids, int max_ids) {
  ensure_cache();
  NdDesc d;
  d.kind = kind; d.dtype = dtype; d.layout = layout; d.forward = forward != 0;
  d.shape.assign(shape, shape + ndim);
  d.stride_in.assign(stride_in, stride_in + ndim);
  d.stride_out.assign(stride_out, stride_out + ndim);
  d.axes.assign(axes, axes + naxes);
  NdPlan plan;
  int rc = g_cache->build_nd(d, &plan, &g_err);
  if (rc) return rc;
  int n = 0;
  for (const Step &st : plan.steps)
    if (n < max_ids) ids[n++] = (st.aux || st.combine) ? 0xffffffffu : st.job.fast_id;
  return n;
}

// which steps the planner marked as the first of a fusable pair (device backend: colfuse2_kernel)
int emu_nd_fuse_flags(int kind, int dtype, int layout, size_t ndim, const size_t *shape, const ptrdiff_t *stride_in,
                      const ptrdiff_t *stride_out, size_t naxes, const size_t *axes, int forward, uint32_t *flags, int max_n) {
  ensure_cache();
  NdDesc d;
  d.kind = kind; d.dtype = dtype; d.layout = layout; d.forward = forward != 0;
  d.shape.assign(shape, shape + ndim);
  d.stride_in.assign(stride_in, stride_in + ndim);
  d.stride_out.assign(stride_out, stride_out + ndim);
  d.axes.assign(axes, axes + naxes);
  NdPlan plan;
  int rc = g_cache->build_nd(d, &plan, &g_err);
  if (rc) return rc;
  int n = 0;
  for (const Step &st : plan.steps)
    if (n < max_n) flags[n++] = st.fuse_with_next ? 1u : 0u;
  return n;
}

// plan introspection for tests
int emu_plan_info(uint32_t L, int dtype, uint32_t *n_fft, int *blue, uint32_t *radices, int max_r) {
  ensure_cache();
  const Engine1D *e = nullptr;
  int rc = g_cache->status_engine(L, dtype, &e, &g_err);
  if (rc) return rc;
  *n_fft = e->n_fft; *blue = e->blue;
  int n = 0;
  for (uint32_t r : e->radices) if (n < max_r) radices[n++] = r;
  return n;
}
}
