// C ABI of libimpulse_fft_b200.so (declared in include/impulse_fft_b200.h and include/pocketfft.h).
//
// Host-side responsibilities: per-device context (table cache, kernel attributes), plan objects,
// pointer classification (device vs host), host staging with a chunked copy/compute pipeline,
// the one-shot plan cache, and error reporting.  The transforms themselves run only on the GPU
// (fft_kernels.cu); if no CUDA device is usable every entry point fails with
// IMPULSE_FFT_ERR_NO_DEVICE — there is no CPU fallback.
#include <cuda_runtime.h>
#include <sched.h>

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <tuple>
#include <vector>

#include "../../include/impulse_fft_b200.h"
#include "../../include/pocketfft.h"
#include "fft_kernels.h"
#include "planner.h"

using namespace impulse;

namespace {

thread_local std::string g_err;
std::atomic<uint64_t> g_launches{0};

int fail(int code, const std::string &msg) {
  g_err = msg;
  return code;
}
int cuda_fail(cudaError_t e, const char *what) {
  g_err = std::string(what) + ": " + cudaGetErrorString(e);
  return (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) ? IMPULSE_FFT_ERR_NO_DEVICE
                                                                        : IMPULSE_FFT_ERR_CUDA;
}

struct CudaAlloc : TableAlloc {
  void *upload(const void *h, size_t n) override {
    void *d = nullptr;
    if (cudaMalloc(&d, n ? n : 1) != cudaSuccess) return nullptr;
    if (cudaMemcpy(d, h, n, cudaMemcpyHostToDevice) != cudaSuccess) { cudaFree(d); return nullptr; }
    return d;
  }
  void release(void *d) override { cudaFree(d); }
};

// Staging resources for host-pointer calls, kept across calls: cudaMalloc/cudaFree and stream
// creation cost milliseconds, comparable to the PCIe time of a whole step.
constexpr int kStageRing = 3;
struct StagePool {
  std::mutex mu;
  cudaStream_t s[2] = {nullptr, nullptr};
  unsigned char *in[2] = {nullptr, nullptr}, *out[2] = {nullptr, nullptr};
  size_t cap_in = 0, cap_out = 0;
  // ring pipeline: one stream per engine (H2D copies, kernels, D2H copies), kStageRing buffers per side, events per slot
  cudaStream_t rs[3] = {nullptr, nullptr, nullptr};
  unsigned char *rin[kStageRing] = {}, *rout[kStageRing] = {};
  cudaEvent_t ev_in[kStageRing] = {}, ev_k[kStageRing] = {}, ev_out[kStageRing] = {};
  size_t rcap_in = 0, rcap_out = 0;
  // whole-span staging (multi-launch plans): buffers kept between calls
  unsigned char *win = nullptr, *wout = nullptr;
  size_t wcap_in = 0, wcap_out = 0;
  cudaError_t ensure_ring(size_t need_in, size_t need_out) {
    cudaError_t e = cudaSuccess;
    for (int i = 0; i < 3 && e == cudaSuccess; ++i)
      if (!rs[i]) e = cudaStreamCreateWithFlags(&rs[i], cudaStreamNonBlocking);
    for (int i = 0; i < kStageRing && e == cudaSuccess; ++i) {
      if (!ev_in[i]) e = cudaEventCreateWithFlags(&ev_in[i], cudaEventDisableTiming);
      if (e == cudaSuccess && !ev_k[i]) e = cudaEventCreateWithFlags(&ev_k[i], cudaEventDisableTiming);
      if (e == cudaSuccess && !ev_out[i]) e = cudaEventCreateWithFlags(&ev_out[i], cudaEventDisableTiming);
    }
    if (e == cudaSuccess && need_in > rcap_in) {
      for (int i = 0; i < kStageRing; ++i) { if (rin[i]) cudaFree(rin[i]); rin[i] = nullptr; }
      rcap_in = 0;
      for (int i = 0; i < kStageRing && e == cudaSuccess; ++i) e = cudaMalloc(&rin[i], need_in);
      if (e == cudaSuccess) rcap_in = need_in;
    }
    if (e == cudaSuccess && need_out > rcap_out) {
      for (int i = 0; i < kStageRing; ++i) { if (rout[i]) cudaFree(rout[i]); rout[i] = nullptr; }
      rcap_out = 0;
      for (int i = 0; i < kStageRing && e == cudaSuccess; ++i) e = cudaMalloc(&rout[i], need_out);
      if (e == cudaSuccess) rcap_out = need_out;
    }
    return e;
  }
  cudaError_t ensure_whole(size_t need_in, size_t need_out) {
    cudaError_t e = cudaSuccess;
    if (!rs[0]) e = ensure_ring(0, 0);
    if (e == cudaSuccess && need_in > wcap_in) {
      if (win) cudaFree(win);
      win = nullptr; wcap_in = 0;
      e = cudaMalloc(&win, need_in);
      if (e == cudaSuccess) wcap_in = need_in;
    }
    if (e == cudaSuccess && need_out > wcap_out) {
      if (wout) cudaFree(wout);
      wout = nullptr; wcap_out = 0;
      e = cudaMalloc(&wout, need_out);
      if (e == cudaSuccess) wcap_out = need_out;
    }
    return e;
  }
  cudaError_t ensure(size_t need_in, size_t need_out) {
    cudaError_t e = cudaSuccess;
    for (int i = 0; i < 2 && e == cudaSuccess; ++i)
      if (!s[i]) e = cudaStreamCreateWithFlags(&s[i], cudaStreamNonBlocking);
    if (e == cudaSuccess && need_in > cap_in) {
      for (int i = 0; i < 2; ++i) { if (in[i]) cudaFree(in[i]); in[i] = nullptr; }
      cap_in = 0;
      for (int i = 0; i < 2 && e == cudaSuccess; ++i) e = cudaMalloc(&in[i], need_in);
      if (e == cudaSuccess) cap_in = need_in;
    }
    if (e == cudaSuccess && need_out > cap_out) {
      for (int i = 0; i < 2; ++i) { if (out[i]) cudaFree(out[i]); out[i] = nullptr; }
      cap_out = 0;
      for (int i = 0; i < 2 && e == cudaSuccess; ++i) e = cudaMalloc(&out[i], need_out);
      if (e == cudaSuccess) cap_out = need_out;
    }
    return e;
  }
};

struct DeviceCtx {
  int dev = -1;
  CudaAlloc alloc;
  std::unique_ptr<PlanCache> cache;
  int sm_count = 0;
  StagePool stage;
};

std::mutex g_ctx_mu;
std::map<int, std::unique_ptr<DeviceCtx>> g_ctx;

int get_ctx(DeviceCtx **out) {
  int dev = -1;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) { cudaGetLastError(); return cuda_fail(e, "no usable CUDA device (this library has no CPU path)"); }
  std::lock_guard<std::mutex> lk(g_ctx_mu);
  auto it = g_ctx.find(dev);
  if (it != g_ctx.end()) { *out = it->second.get(); return 0; }
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, dev);
  if (e != cudaSuccess) return cuda_fail(e, "cudaGetDeviceProperties");
  if (prop.major < 10)
    return fail(IMPULSE_FFT_ERR_NO_DEVICE, std::string("device '") + prop.name + "' is not sm_100-class; this library is built for B200 (sm_100a) only");
  std::unique_ptr<DeviceCtx> c(new DeviceCtx);
  c->dev = dev;
  c->sm_count = prop.multiProcessorCount;
  c->cache.reset(new PlanCache(&c->alloc));
  c->cache->max_smem = prop.sharedMemPerBlockOptin;
  {  // scratch comes from the stream-ordered pool: keep freed blocks instead of returning them to the OS
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
      uint64_t keep = ~0ull;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    cudaGetLastError();
  }
  int rc = configure_kernels(prop.sharedMemPerBlockOptin);
  if (rc) return cuda_fail((cudaError_t)rc, "cudaFuncSetAttribute");
  rc = init_sched_slots();
  if (rc) return cuda_fail((cudaError_t)rc, "row-scheduler allocation");
  *out = c.get();
  g_ctx[dev] = std::move(c);
  return 0;
}

bool is_device_ptr(const void *p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

}  // namespace

struct impulse_fft_plan_s {
  NdPlan nd;
  DeviceCtx *ctx = nullptr;
  size_t in_esz = 0, out_esz = 0;  // alignment required of the base pointers
};

namespace {

size_t side_align(const NdDesc &d, bool input) {
  const size_t r = d.dtype == DT_F64 ? 8 : 4;
  bool real_side;
  const bool r2r = d.kind >= KIND_DCT && d.kind <= KIND_HARTLEY_GEN;  // DCT, DST, FFTPACK, Hartley: real on both sides
  if (input) real_side = r2r || d.kind == KIND_R2C || (d.kind == KIND_C2R && d.layout == RL_HALFCOMPLEX);
  else real_side = r2r || d.kind == KIND_C2R || (d.kind == KIND_R2C && d.layout == RL_HALFCOMPLEX);
  return real_side ? r : 2 * r;
}

// run all steps with device pointers on `stream`
int run_device(impulse_fft_plan p, const void *in, void *out, double fct, cudaStream_t stream, const void *umul = nullptr) {
  const NdPlan &nd = p->nd;
  if (nd.empty) return 0;
  if (((uintptr_t)in % p->in_esz) || ((uintptr_t)out % p->out_esz))
    return fail(IMPULSE_FFT_ERR_STRIDE, "data pointer is not aligned to its element size");
  if (in == out && (nd.desc.kind == KIND_C2C || (nd.desc.kind >= KIND_DCT && nd.desc.kind <= KIND_HARTLEY_GEN)) &&
      nd.desc.stride_in != nd.desc.stride_out)
    return fail(IMPULSE_FFT_ERR_STRIDE, "stride mismatch");  // hdronly.h:455-456
  if ((nd.cplx_view_in && ((uintptr_t)in % (2 * p->in_esz))) || (nd.cplx_view_out && ((uintptr_t)out % (2 * p->out_esz))))
    return fail(IMPULSE_FFT_ERR_STRIDE, "long real lines are processed as complex pairs: the real array must be aligned to 2 elements");
  void *tmp = nullptr, *tmp2 = nullptr, *tmp3 = nullptr, *tmp4 = nullptr;
  if (nd.tmp4_bytes) {
    cudaError_t e = cudaMallocAsync(&tmp4, nd.tmp4_bytes, stream);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMallocAsync(tmp4)");
  }
  if (nd.tmp_bytes) {
    cudaError_t e = cudaMallocAsync(&tmp, nd.tmp_bytes, stream);
    if (e != cudaSuccess) { if (tmp4) cudaFreeAsync(tmp4, stream); return cuda_fail(e, "cudaMallocAsync(tmp)"); }
  }
  // pairs of launches that run as one fused kernel keep their intermediate in a small L2-resident ring instead of the
  // full-size four-step scratch
  static const bool no_fuse = [] { const char *e = std::getenv("IMPULSE_FFT_NO_COLFUSE"); return e && std::atoi(e) != 0; }();
  size_t fuse_bytes = 0;
  bool all_fused = true;   // every pair the planner marked runs fused (then the full-size scratch shrinks to tmp2_bytes_fused)
  auto can_fuse = [&](size_t i) {
    return !no_fuse && i + 1 < nd.steps.size() && nd.steps[i].fuse_with_next && !nd.steps[i].job.seg_len &&
           colfuse_pair_supported(nd.steps[i].job.fast_id, nd.steps[i + 1].job.fast_id);
  };
  for (size_t i = 0; i + 1 < nd.steps.size(); ++i) {
    if (!nd.steps[i].fuse_with_next) continue;
    if (can_fuse(i))
      fuse_bytes = std::max(fuse_bytes, colfuse_scratch_bytes(nd.steps[i].job, nd.steps[i + 1].job, nd.steps[i].fuse_tiles, nullptr, nullptr));
    else
      all_fused = false;
  }
  const size_t tmp2_need = (fuse_bytes && all_fused) ? nd.tmp2_bytes_fused : nd.tmp2_bytes;
  void *fuse_scratch = nullptr;
  if (fuse_bytes) {
    cudaError_t e = cudaMallocAsync(&fuse_scratch, fuse_bytes, stream);
    if (e != cudaSuccess) { if (tmp) cudaFreeAsync(tmp, stream); if (tmp4) cudaFreeAsync(tmp4, stream); return cuda_fail(e, "cudaMallocAsync(ring)"); }
  }
  if (tmp2_need) {
    cudaError_t e = cudaMallocAsync(&tmp2, tmp2_need, stream);
    if (e != cudaSuccess) { if (tmp) cudaFreeAsync(tmp, stream); if (tmp4) cudaFreeAsync(tmp4, stream); return cuda_fail(e, "cudaMallocAsync(tmp2)"); }
  }
  if (nd.tmp3_bytes) {
    cudaError_t e = cudaMallocAsync(&tmp3, nd.tmp3_bytes, stream);
    if (e != cudaSuccess) {
      if (tmp) cudaFreeAsync(tmp, stream);
      if (tmp2) cudaFreeAsync(tmp2, stream);
      if (tmp4) cudaFreeAsync(tmp4, stream);
      return cuda_fail(e, "cudaMallocAsync(tmp3)");
    }
  }
  int rc = 0;
  for (size_t si = 0; si < nd.steps.size(); ++si) {
    const Step &st = nd.steps[si];
    LineJob J = st.job;
    const unsigned char *src = st.src == BUF_IN ? (const unsigned char *)in : st.src == BUF_OUT ? (const unsigned char *)out
                               : st.src == BUF_TMP ? (const unsigned char *)tmp
                               : st.src == BUF_TMP2 ? (const unsigned char *)tmp2
                               : st.src == BUF_TMP3 ? (const unsigned char *)tmp3 : (const unsigned char *)tmp4;
    unsigned char *dst = st.dst == BUF_OUT ? (unsigned char *)out : st.dst == BUF_TMP ? (unsigned char *)tmp
                         : st.dst == BUF_TMP2 ? (unsigned char *)tmp2
                         : st.dst == BUF_TMP3 ? (unsigned char *)tmp3 : (unsigned char *)tmp4;
    if (st.aux) {
      AuxJob aj = st.aj;
      aj.in = src + st.src_off_bytes;
      aj.out = dst + st.dst_off_bytes;
      aj.fct = st.takes_fct ? fct : 1.0;
      int e = launch_aux(aj, p->ctx->sm_count, stream);
      g_launches.fetch_add(1, std::memory_order_relaxed);
      if (e) { rc = cuda_fail((cudaError_t)e, "kernel launch"); break; }
      continue;
    }
    if (st.combine) {
      CombineJob cj = st.cj;
      cj.in = src + st.src_off_bytes;
      cj.out = dst + st.dst_off_bytes;
      int e = launch_hartley_combine(cj, p->ctx->sm_count, stream);
      g_launches.fetch_add(1, std::memory_order_relaxed);
      if (e) { rc = cuda_fail((cudaError_t)e, "kernel launch"); break; }
      continue;
    }
    J.in = src + st.src_off_bytes;
    J.out = dst + st.dst_off_bytes;
    J.fct = st.takes_fct ? fct : 1.0;
    if (fuse_scratch && can_fuse(si)) {
      // this launch and the next one as ONE persistent kernel; the intermediate stays in the ring
      const Step &sb = nd.steps[si + 1];
      LineJob JB = sb.job;
      unsigned char *dstb = sb.dst == BUF_OUT ? (unsigned char *)out : sb.dst == BUF_TMP ? (unsigned char *)tmp
                            : sb.dst == BUF_TMP3 ? (unsigned char *)tmp3 : (unsigned char *)tmp4;
      JB.out = dstb + sb.dst_off_bytes;
      JB.fct = sb.takes_fct ? fct : 1.0;
      {
        size_t ctrl_off = 0, ctrl_bytes = 0;
        colfuse_scratch_bytes(J, JB, st.fuse_tiles, &ctrl_off, &ctrl_bytes);
        cudaError_t me = cudaMemsetAsync((unsigned char *)fuse_scratch + ctrl_off, 0, ctrl_bytes, stream);
        if (me != cudaSuccess) { rc = cuda_fail(me, "cudaMemsetAsync(ring control)"); break; }
        int e = launch_colfuse_pair(J, JB, st.fuse_tiles, st.fuse_g0n, fuse_scratch, p->ctx->sm_count, stream);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        if (e) { rc = e == -1 ? fail(IMPULSE_FFT_ERR_UNSUPPORTED, "internal: no fused kernel for a pair marked fusable") : cuda_fail((cudaError_t)e, "kernel launch"); break; }
        ++si;   // the next step ran inside this launch
        continue;
      }
    }
    if (st.takes_umul) {
      if (!umul) { rc = fail(IMPULSE_FFT_ERR_INVALID, "this plan needs a multiplier array"); break; }
      J.umul = umul;
    } else {
      J.umul_mod = 0;
    }
    const size_t r = J.dtype == 1 ? 8 : 4;
    if ((J.flags & F_VEC_IN) && ((uintptr_t)J.in % (2 * r))) J.flags &= ~F_VEC_IN;
    if ((J.flags & F_VEC_OUT) && ((uintptr_t)J.out % (2 * r))) J.flags &= ~F_VEC_OUT;
    // the register kernels address real rows as complex pairs: fall back to the generic engine otherwise
    if (((J.fast_id >= FAST3_2048_F64 && J.fast_id <= FAST3_1000_F64) || (J.fast_id >= FAST3R_256_F64 && J.fast_id <= FAST3P_512_F32) || (J.fast_id >= FAST3_1536_F64 && J.fast_id <= FAST3_6561_F64) || (J.fast_id >= FAST3_1536_F32 && J.fast_id <= FAST3_6561_F32) || (J.fast_id >= FAST2R_8_F64 && J.fast_id <= FAST2R_4_F32)) && (((uintptr_t)J.in % (2 * r)) || ((uintptr_t)J.out % (2 * r)))) J.fast_id = FAST_NONE;
    int e = launch_line_job(J, st.cfg.threads, st.cfg.smem_bytes, st.cfg.n_tiles, stream);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    if (e) { rc = cuda_fail((cudaError_t)e, "kernel launch"); break; }
  }
  if (tmp) cudaFreeAsync(tmp, stream);
  if (tmp2) cudaFreeAsync(tmp2, stream);
  if (tmp3) cudaFreeAsync(tmp3, stream);
  if (tmp4) cudaFreeAsync(tmp4, stream);
  if (fuse_scratch) cudaFreeAsync(fuse_scratch, stream);
  return rc;
}

// ---- host staging -----------------------------------------------------------
// Pipelined path for the common case (single-step plan whose slowest batch dimension cuts both
// arrays into disjoint slabs): chunks flow H2D -> kernel -> D2H on alternating streams so that
// the two PCIe directions and the kernel overlap.  Anything else: whole-span staging.
int run_host(impulse_fft_plan p, const void *in, void *out, double fct) {
  const NdPlan &nd = p->nd;
  if (nd.empty) return 0;
  const bool inplace = (in == out);
  // IMPULSE_FFT_ZEROCOPY=1 (prepared, not yet measured): a single-launch plan on PINNED host buffers (cudaHostAlloc /
  // cudaHostRegister memory is mapped into the device's address space under UVA) runs the kernel directly on them, so
  // the row loads and stores cross PCIe themselves — no staging copies, no pipeline fill and drain.  Pageable
  // memory, multi-launch plans and everything else take the staged paths below.
  static const bool zero_copy = [] { const char *e = std::getenv("IMPULSE_FFT_ZEROCOPY"); return e && std::atoi(e) != 0; }();
  if (zero_copy && nd.steps.size() == 1 && !nd.steps[0].aux && !nd.steps[0].combine && nd.tmp_bytes == 0 && nd.tmp2_bytes == 0 &&
      nd.tmp3_bytes == 0 && nd.tmp4_bytes == 0) {
    cudaPointerAttributes ai, ao;
    if (cudaPointerGetAttributes(&ai, in) == cudaSuccess && cudaPointerGetAttributes(&ao, out) == cudaSuccess &&
        ai.type == cudaMemoryTypeHost && ao.type == cudaMemoryTypeHost && ai.devicePointer && ao.devicePointer) {
      int rc = run_device(p, ai.devicePointer, ao.devicePointer, fct, nullptr);
      cudaError_t se = cudaStreamSynchronize(nullptr);
      if (!rc && se != cudaSuccess) rc = cuda_fail(se, "zero-copy sync");
      return rc;
    }
    cudaGetLastError();
  }
  // ---- try the chunked pipeline
  if (nd.steps.size() == 1 && !nd.steps[0].aux && !nd.steps[0].combine && nd.tmp_bytes == 0 && nd.tmp2_bytes == 0 &&
      nd.tmp3_bytes == 0 && nd.tmp4_bytes == 0) {
    const Step &st = nd.steps[0];
    const LineJob &J0 = st.job;
    int od = -1;  // outermost batch dim in use
    for (int d = kMaxBatchDims - 1; d >= 0; --d) if (J0.bdim[d] > 1) { od = d; break; }
    const size_t in_e = side_align(nd.desc, true), out_e = side_align(nd.desc, false);
    if (od >= 0 && J0.bs_in[od] > 0 && J0.bs_out[od] > 0) {
      const int64_t sin_b = J0.bs_in[od] * (int64_t)in_e, sout_b = J0.bs_out[od] * (int64_t)out_e;
      // span of one slab (index fixed along od) on each side
      const int64_t in_slab = (nd.in_hi - nd.in_lo) - (int64_t)(J0.bdim[od] - 1) * sin_b;
      const int64_t out_slab = (nd.out_hi - nd.out_lo) - (int64_t)(J0.bdim[od] - 1) * sout_b;
      const bool disjoint = in_slab <= sin_b && out_slab <= sout_b && nd.in_lo == 0 && nd.out_lo == 0;
      const uint64_t total_bytes = (uint64_t)(nd.in_hi - nd.in_lo) + (uint64_t)(nd.out_hi - nd.out_lo);
      if (disjoint && J0.bdim[od] >= 4 && total_bytes >= (8u << 20)) {
        uint64_t inner_lines = 1;
        for (int d = 0; d < od; ++d) inner_lines *= J0.bdim[d];
        // ~64 MiB of traffic per chunk (IMPULSE_FFT_STAGE_MB overrides; measured 8:30.6 16:26.0 32:23.9 64:23.2 128:23.1 ms), at least 4 chunks
        static const uint64_t stage_bytes = [] {
          const char *e = std::getenv("IMPULSE_FFT_STAGE_MB");
          const long mb = e ? std::atol(e) : 64;
          return (uint64_t)(mb > 0 ? mb : 64) << 20;
        }();
        uint64_t per = std::max<uint64_t>(1, stage_bytes / std::max<int64_t>(1, sin_b + sout_b));
        per = std::min<uint64_t>(per, (J0.bdim[od] + 3) / 4);
        // chunk schedule: equal chunks.  IMPULSE_FFT_STAGE_RAMP=1 adds a geometric ramp at both ends (the pipeline's fill
        // and drain shrink from a whole chunk to an eighth of one); measured on config 2 it is 0.2 ms SLOWER (23.6 vs
        // 23.4 ms, profiles/r02_ab_e2e.txt): the 1 GiB each way over PCIe is the floor, not the fill / drain.
        static const bool ramp = [] { const char *e = std::getenv("IMPULSE_FFT_STAGE_RAMP"); return e && std::atoi(e) != 0; }();
        std::vector<uint64_t> sizes;
        {
          const uint64_t total = J0.bdim[od];
          std::vector<uint64_t> head;
          if (ramp && per >= 8 && total >= 4 * per)
            for (uint64_t c = std::max<uint64_t>(1, per / 8); c < per; c *= 2) head.push_back(c);
          uint64_t used = 0;
          for (uint64_t c : head) used += 2 * c;           // the ramp appears at both ends
          for (uint64_t c : head) sizes.push_back(c);
          uint64_t mid = total - used;
          while (mid > 0) { const uint64_t c = std::min(per, mid); sizes.push_back(c); mid -= c; }
          for (size_t i = head.size(); i-- > 0;) sizes.push_back(head[i]);
        }
        StagePool &sg = p->ctx->stage;
        std::lock_guard<std::mutex> stage_lock(sg.mu);
        const size_t cin = (size_t)((per - 1) * sin_b + in_slab), cout = (size_t)((per - 1) * sout_b + out_slab);
        // IMPULSE_FFT_STAGE_RING (default 1): one stream per engine — H2D copies, kernels, D2H copies — over a ring of
        // kStageRing buffer pairs with an event per slot, so that neither copy engine ever waits for the OTHER direction
        // of its own chunk (the two-stream form serialises H2D(c+2) behind D2H(c)); 0 = the two-stream form below
        static const bool ring = [] { const char *e = std::getenv("IMPULSE_FFT_STAGE_RING"); return !e || std::atoi(e) != 0; }();
        if (ring) {
          cudaError_t e = sg.ensure_ring(cin, inplace ? 0 : cout);
          int rc = 0;
          if (e != cudaSuccess) rc = cuda_fail(e, "staging allocation");
          uint64_t lo = 0;
          for (size_t c = 0; c < sizes.size() && !rc; ++c) {
            const int b = (int)(c % kStageRing);
            const bool reuse = c >= (size_t)kStageRing;
            const uint64_t cnt = sizes[c];
            const size_t bin = (size_t)((cnt - 1) * sin_b + in_slab), bout = (size_t)((cnt - 1) * sout_b + out_slab);
            const unsigned char *hin = (const unsigned char *)in + lo * sin_b;
            unsigned char *hout = (unsigned char *)out + lo * sout_b;
            lo += cnt;
            unsigned char *din = sg.rin[b], *dout = inplace ? sg.rin[b] : sg.rout[b];
            // H2D: the input slot is free once the kernel of chunk c - ring has read it (in place: once its D2H is done)
            if (reuse) e = cudaStreamWaitEvent(sg.rs[0], inplace ? sg.ev_out[b] : sg.ev_k[b], 0);
            if (e == cudaSuccess) e = cudaMemcpyAsync(din, hin, bin, cudaMemcpyHostToDevice, sg.rs[0]);
            if (e == cudaSuccess && !nd.out_dense && !inplace) {   // strided output with gaps: preload so the gaps survive
              if (reuse) e = cudaStreamWaitEvent(sg.rs[0], sg.ev_out[b], 0);
              if (e == cudaSuccess) e = cudaMemcpyAsync(dout, hout, bout, cudaMemcpyHostToDevice, sg.rs[0]);
            }
            if (e == cudaSuccess) e = cudaEventRecord(sg.ev_in[b], sg.rs[0]);
            if (e != cudaSuccess) { rc = cuda_fail(e, "H2D copy"); break; }
            // kernel: input landed, output slot drained
            e = cudaStreamWaitEvent(sg.rs[1], sg.ev_in[b], 0);
            if (e == cudaSuccess && reuse) e = cudaStreamWaitEvent(sg.rs[1], sg.ev_out[b], 0);
            if (e != cudaSuccess) { rc = cuda_fail(e, "stream wait"); break; }
            LineJob J = J0;
            J.bdim[od] = cnt;
            J.n_lines = inner_lines * cnt;
            J.in = din;
            J.out = dout;
            J.fct = st.takes_fct ? fct : 1.0;
            const uint64_t C = 1ull << J.log_c;
            int le = launch_line_job(J, st.cfg.threads, st.cfg.smem_bytes, (J.n_lines + C - 1) / C, sg.rs[1]);
            g_launches.fetch_add(1, std::memory_order_relaxed);
            if (le) { rc = cuda_fail((cudaError_t)le, "kernel launch"); break; }
            e = cudaEventRecord(sg.ev_k[b], sg.rs[1]);
            // D2H
            if (e == cudaSuccess) e = cudaStreamWaitEvent(sg.rs[2], sg.ev_k[b], 0);
            if (e == cudaSuccess) e = cudaMemcpyAsync(hout, dout, inplace ? bin : bout, cudaMemcpyDeviceToHost, sg.rs[2]);
            if (e == cudaSuccess) e = cudaEventRecord(sg.ev_out[b], sg.rs[2]);
            if (e != cudaSuccess) { rc = cuda_fail(e, "D2H copy"); break; }
          }
          for (int i = 0; i < 3; ++i)
            if (sg.rs[i]) { cudaError_t se = cudaStreamSynchronize(sg.rs[i]); if (se != cudaSuccess && !rc) rc = cuda_fail(se, "staging sync"); }
          return rc;
        }
        cudaError_t e = sg.ensure(cin, inplace ? 0 : cout);
        unsigned char **dbuf_in = sg.in, **dbuf_out = sg.out;
        int rc = 0;
        if (e != cudaSuccess) rc = cuda_fail(e, "staging allocation");
        uint64_t lo = 0;
        for (size_t c = 0; c < sizes.size() && !rc; ++c) {
          const int b = (int)(c & 1);
          const uint64_t cnt = sizes[c];
          const size_t bin = (size_t)((cnt - 1) * sin_b + in_slab), bout = (size_t)((cnt - 1) * sout_b + out_slab);
          const unsigned char *hin = (const unsigned char *)in + lo * sin_b;
          unsigned char *hout = (unsigned char *)out + lo * sout_b;
          lo += cnt;
          e = cudaMemcpyAsync(dbuf_in[b], hin, bin, cudaMemcpyHostToDevice, sg.s[b]);
          // strided output with gaps: preload so the gaps survive the write-back
          if (e == cudaSuccess && !nd.out_dense && !inplace)
            e = cudaMemcpyAsync(dbuf_out[b], hout, bout, cudaMemcpyHostToDevice, sg.s[b]);
          if (e != cudaSuccess) { rc = cuda_fail(e, "H2D copy"); break; }
          LineJob J = J0;
          J.bdim[od] = cnt;
          J.n_lines = inner_lines * cnt;
          J.in = dbuf_in[b];
          J.out = inplace ? (void *)dbuf_in[b] : (void *)dbuf_out[b];
          J.fct = st.takes_fct ? fct : 1.0;
          const uint64_t C = 1ull << J.log_c;
          int le = launch_line_job(J, st.cfg.threads, st.cfg.smem_bytes, (J.n_lines + C - 1) / C, sg.s[b]);
          g_launches.fetch_add(1, std::memory_order_relaxed);
          if (le) { rc = cuda_fail((cudaError_t)le, "kernel launch"); break; }
          e = cudaMemcpyAsync(hout, inplace ? dbuf_in[b] : dbuf_out[b], inplace ? bin : bout, cudaMemcpyDeviceToHost, sg.s[b]);
          if (e != cudaSuccess) { rc = cuda_fail(e, "D2H copy"); break; }
        }
        for (int i = 0; i < 2; ++i)
          if (sg.s[i]) { cudaError_t se = cudaStreamSynchronize(sg.s[i]); if (se != cudaSuccess && !rc) rc = cuda_fail(se, "staging sync"); }
        return rc;
      }
    }
  }
  // ---- whole-span staging (multi-launch plans: every launch needs the whole array).  The device copies are kept
  // between calls (two cudaMalloc / cudaFree of the array size per call cost milliseconds and synchronise the device)
  // and copies and launches run on one stream of the pool.
  const size_t bin = (size_t)(nd.in_hi - nd.in_lo), bout = (size_t)(nd.out_hi - nd.out_lo);
  StagePool &sg = p->ctx->stage;
  std::lock_guard<std::mutex> stage_lock(sg.mu);
  cudaError_t e = sg.ensure_whole(bin, inplace ? 0 : bout);
  if (e != cudaSuccess) return cuda_fail(e, "staging allocation");
  unsigned char *din = sg.win, *dout = inplace ? nullptr : sg.wout;
  cudaStream_t ws = sg.rs[1];
  int rc = 0;
  e = cudaMemcpyAsync(din, (const unsigned char *)in + nd.in_lo, bin, cudaMemcpyHostToDevice, ws);
  if (e == cudaSuccess && !inplace && !nd.out_dense)  // preserve what lies between strided output elements
    e = cudaMemcpyAsync(dout, (const unsigned char *)out + nd.out_lo, bout, cudaMemcpyHostToDevice, ws);
  if (e != cudaSuccess) rc = cuda_fail(e, "H2D copy");
  if (!rc) {
    unsigned char *o = inplace ? din : dout;
    const ptrdiff_t olo = inplace ? nd.in_lo : nd.out_lo;
    rc = run_device(p, din - nd.in_lo, o - olo, fct, ws);
    if (!rc) {
      e = cudaMemcpyAsync((unsigned char *)out + olo, o, inplace ? bin : bout, cudaMemcpyDeviceToHost, ws);
      if (e != cudaSuccess) rc = cuda_fail(e, "D2H copy");
    }
  }
  e = cudaStreamSynchronize(ws);
  if (e != cudaSuccess && !rc) rc = cuda_fail(e, "staging sync");
  // the buffers stay for the next call up to IMPULSE_FFT_STAGE_KEEP_MB (default 4096 MiB for the pair); larger ones are
  // returned at once, so that one huge host transform does not hold device memory for the life of the process
  static const size_t keep_bytes = [] {
    const char *v = std::getenv("IMPULSE_FFT_STAGE_KEEP_MB");
    const long mb = v ? std::atol(v) : 4096;
    return (size_t)(mb > 0 ? mb : 0) << 20;
  }();
  if (sg.wcap_in + sg.wcap_out > keep_bytes) {
    if (sg.win) cudaFree(sg.win);
    if (sg.wout) cudaFree(sg.wout);
    sg.win = sg.wout = nullptr;
    sg.wcap_in = sg.wcap_out = 0;
  }
  return rc;
}

int make_desc(NdDesc *d, int kind, int dtype, int layout, int forward, size_t ndim, const size_t *shape,
              const ptrdiff_t *sin, const ptrdiff_t *sout, size_t naxes, const size_t *axes) {
  if (!shape || !sin || !sout || !axes) return fail(IMPULSE_FFT_ERR_INVALID, "null descriptor array");
  if (ndim < 1) return fail(IMPULSE_FFT_ERR_INVALID, "ndim must be >= 1");
  if (ndim > IMPULSE_FFT_MAX_DIMS || naxes > IMPULSE_FFT_MAX_DIMS) return fail(IMPULSE_FFT_ERR_INVALID, "too many dimensions");
  d->kind = kind; d->dtype = dtype; d->layout = layout; d->forward = forward != 0;
  d->shape.assign(shape, shape + ndim);
  d->stride_in.assign(sin, sin + ndim);
  d->stride_out.assign(sout, sout + ndim);
  d->axes.assign(axes, axes + naxes);
  return 0;
}

int create_plan(impulse_fft_plan *out, const NdDesc &d) {
  DeviceCtx *ctx = nullptr;
  int rc = get_ctx(&ctx);
  if (rc) return rc;
  std::unique_ptr<impulse_fft_plan_s> p(new impulse_fft_plan_s);
  p->ctx = ctx;
  std::string err;
  rc = ctx->cache->build_nd(d, &p->nd, &err);
  if (rc) return fail(rc, err);
  p->in_esz = side_align(d, true);
  p->out_esz = side_align(d, false);
  *out = p.release();
  return 0;
}

// one-shot plan cache (the role of get_plan, hdronly.h:2655-2706)
struct OneShotKey {
  int dev, kind, dtype, layout, forward;  // for DCT/DST `layout` carries the type and `forward` the ortho flag
  std::vector<size_t> shape, axes;
  std::vector<ptrdiff_t> sin, sout;
  uint64_t umul_mod = 0;
  bool operator<(const OneShotKey &o) const {
    return std::tie(dev, kind, dtype, layout, forward, shape, axes, sin, sout, umul_mod) <
           std::tie(o.dev, o.kind, o.dtype, o.layout, o.forward, o.shape, o.axes, o.sin, o.sout, o.umul_mod);
  }
};
std::mutex g_os_mu;
LruMap<OneShotKey, std::shared_ptr<impulse_fft_plan_s>> g_os;   // evicts the least recently used plan, one at a time
constexpr size_t kOneShotCap = 64;

int one_shot(int kind, int dtype, int layout, size_t ndim, const size_t *shape, const ptrdiff_t *sin,
             const ptrdiff_t *sout, size_t naxes, const size_t *axes, int forward, const void *in, void *out,
             double fct, void *stream, const void *umul = nullptr, uint64_t umul_mod = 0) {
  NdDesc d;
  // real-to-real kinds carry their own options in the `layout` / `forward` slots of this helper
  const bool r2r = kind == KIND_DCT || kind == KIND_DST, fpk = kind == KIND_FFTPACK,
             hart = kind == KIND_HARTLEY_SEP || kind == KIND_HARTLEY_GEN;
  int rc = make_desc(&d, kind, dtype, (r2r || fpk || hart) ? RL_HERMITIAN : layout, (r2r || hart) ? 1 : forward, ndim, shape,
                     sin, sout, naxes, axes);
  if (rc) return rc;
  if (r2r) { d.r2r_type = layout; d.ortho = forward != 0; }
  if (fpk) d.real2hermitian = layout != 0;
  d.umul_mod = umul_mod;
  int dev = -1;
  if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); dev = -1; }
  OneShotKey key{dev, kind, dtype, layout, forward != 0, d.shape, d.axes, d.stride_in, d.stride_out, umul_mod};
  std::shared_ptr<impulse_fft_plan_s> plan;
  {
    std::lock_guard<std::mutex> lk(g_os_mu);
    if (auto *hit = g_os.find(key)) plan = *hit;
  }
  if (!plan) {
    impulse_fft_plan raw = nullptr;
    rc = create_plan(&raw, d);
    if (rc) return rc;
    plan.reset(raw);
    std::lock_guard<std::mutex> lk(g_os_mu);
    g_os.insert(key, plan, kOneShotCap);   // a thread still executing an evicted plan holds its own reference
  }
  if (umul_mod) {  // fused-multiply plans: device pointers (positive output strides: checked by the entry points)
    if (!in || !out) return fail(IMPULSE_FFT_ERR_INVALID, "null data pointer");
    if (!is_device_ptr(in) || !is_device_ptr(out) || !is_device_ptr(umul))
      return fail(IMPULSE_FFT_ERR_INVALID, "transforms with a fused multiply take device pointers");
    return run_device(plan.get(), in, out, fct, static_cast<cudaStream_t>(stream), umul);
  }
  return impulse_fft_execute(plan.get(), in, out, fct, stream);
}

}  // namespace

// =============================================================================
extern "C" {

const char *impulse_fft_last_error(void) { return g_err.c_str(); }
const char *impulse_fft_version(void) { return "impulse_fft_b200 0.1 (sm_100a)"; }
const char *impulse_fft_last_kernel(void) { return impulse::g_last_kernel; }
uint64_t impulse_fft_launch_count(void) { return g_launches.load(); }

int impulse_fft_plan_create(impulse_fft_plan *out, const impulse_fft_desc *desc) {
  if (!out || !desc) return fail(IMPULSE_FFT_ERR_INVALID, "null argument");
  *out = nullptr;
  if (desc->ndim > IMPULSE_FFT_MAX_DIMS || desc->naxes > IMPULSE_FFT_MAX_DIMS)
    return fail(IMPULSE_FFT_ERR_INVALID, "too many dimensions");
  NdDesc d;
  int rc = make_desc(&d, desc->kind, desc->dtype, desc->real_layout, desc->forward, desc->ndim, desc->shape,
                     desc->stride_in, desc->stride_out, desc->naxes, desc->axes);
  if (rc) return rc;
  return create_plan(out, d);
}

int impulse_fft_plan_destroy(impulse_fft_plan plan) {
  delete plan;
  return 0;
}

int impulse_fft_execute(impulse_fft_plan plan, const void *in, void *out, double fct, void *stream) {
  if (!plan) return fail(IMPULSE_FFT_ERR_INVALID, "null plan");
  if (plan->nd.empty) return 0;
  if (!in || !out) return fail(IMPULSE_FFT_ERR_INVALID, "null data pointer");
  int dev = -1;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return cuda_fail(e, "cudaGetDevice");
  if (dev != plan->ctx->dev) return fail(IMPULSE_FFT_ERR_INVALID, "plan was created on another device");
  {  // in-place is defined for c2c / r2r with equal strides and for the packed real layout only
    const NdDesc &d = plan->nd.desc;     // (README_pocketfft.md:133-135)
    if (in == out && (d.kind == KIND_R2C || d.kind == KIND_C2R) && d.layout != RL_HALFCOMPLEX)
      return fail(IMPULSE_FFT_ERR_INVALID, "in-place r2c/c2r is only defined for the FFTPACK halfcomplex layout");
  }
  const bool din = is_device_ptr(in), dout = is_device_ptr(out);
  if (din != dout) return fail(IMPULSE_FFT_ERR_INVALID, "input and output must both be device or both be host memory");
  if (din) return run_device(plan, in, out, fct, static_cast<cudaStream_t>(stream));
  return run_host(plan, in, out, fct);
}

int impulse_fft_plan_get_info(impulse_fft_plan plan, impulse_fft_plan_info *info) {
  if (!plan || !info) return fail(IMPULSE_FFT_ERR_INVALID, "null argument");
  std::memset(info, 0, sizeof(*info));
  info->n_steps = (uint32_t)plan->nd.steps.size();
  info->tmp_bytes = plan->nd.tmp_bytes + plan->nd.tmp2_bytes + plan->nd.tmp3_bytes + plan->nd.tmp4_bytes;
  if (!plan->nd.steps.empty()) {
    const Step &s = plan->nd.steps[0];
    info->n_fft = s.job.n_fft;
    info->lines_per_cta = 1u << s.job.log_c;
    info->threads = (uint32_t)s.cfg.threads;
    info->smem_bytes = (uint32_t)s.cfg.smem_bytes;
    uint32_t n = 0;
    for (int i = 0; i < s.job.nphases; ++i) {
      if (s.job.ph[i].op == OP_BLUE_PRE) info->bluestein = 1;
      if (s.job.ph[i].op == OP_PASS_DIF && n < 32) info->radices[n++] = s.job.ph[i].radix;
    }
    info->n_radices = n;
  }
  return 0;
}

int impulse_fft_c2c(int dtype, size_t ndim, const size_t *shape, const ptrdiff_t *stride_in,
                    const ptrdiff_t *stride_out, size_t naxes, const size_t *axes, int forward,
                    const void *data_in, void *data_out, double fct, size_t, void *stream) {
  return one_shot(KIND_C2C, dtype, RL_HERMITIAN, ndim, shape, stride_in, stride_out, naxes, axes, forward,
                  data_in, data_out, fct, stream);
}
int impulse_fft_r2c(int dtype, size_t ndim, const size_t *shape, const ptrdiff_t *stride_in,
                    const ptrdiff_t *stride_out, size_t naxes, const size_t *axes, int forward,
                    const void *data_in, void *data_out, double fct, size_t, void *stream) {
  return one_shot(KIND_R2C, dtype, RL_HERMITIAN, ndim, shape, stride_in, stride_out, naxes, axes, forward,
                  data_in, data_out, fct, stream);
}
int impulse_fft_c2r(int dtype, size_t ndim, const size_t *shape, const ptrdiff_t *stride_in,
                    const ptrdiff_t *stride_out, size_t naxes, const size_t *axes, int forward,
                    const void *data_in, void *data_out, double fct, size_t, void *stream) {
  return one_shot(KIND_C2R, dtype, RL_HERMITIAN, ndim, shape, stride_in, stride_out, naxes, axes, forward,
                  data_in, data_out, fct, stream);
}

int impulse_fft_c2c_mul(int dtype, size_t ndim, const size_t *shape, const ptrdiff_t *stride_in,
                        const ptrdiff_t *stride_out, size_t naxes, const size_t *axes, int forward, const void *data_in,
                        void *data_out, double fct, const void *mul, size_t mul_elems, void *stream) {
  if (!mul || !mul_elems) return fail(IMPULSE_FFT_ERR_INVALID, "null multiplier");
  if (!stride_out) return fail(IMPULSE_FFT_ERR_INVALID, "null descriptor array");
  for (size_t i = 0; i < ndim && i < IMPULSE_FFT_MAX_DIMS; ++i)
    if (stride_out[i] <= 0) return fail(IMPULSE_FFT_ERR_STRIDE, "the output of a fused multiply needs positive strides");
  return one_shot(KIND_C2C, dtype, RL_HERMITIAN, ndim, shape, stride_in, stride_out, naxes, axes, forward, data_in, data_out, fct,
                  stream, mul, mul_elems);
}

int impulse_fft_convolve_axis(int dtype, size_t ndim, const size_t *shape, const ptrdiff_t *stride_in,
                              const ptrdiff_t *stride_out, size_t axis, const void *data_in, void *data_out, double fct,
                              const void *mul, size_t mul_elems, void *stream) {
  if (!mul || !mul_elems) return fail(IMPULSE_FFT_ERR_INVALID, "null multiplier");
  if (!stride_out) return fail(IMPULSE_FFT_ERR_INVALID, "null descriptor array");
  for (size_t i = 0; i < ndim && i < IMPULSE_FFT_MAX_DIMS; ++i)
    if (stride_out[i] <= 0) return fail(IMPULSE_FFT_ERR_STRIDE, "the output of a fused multiply needs positive strides");
  const size_t axes[1] = {axis};
  return one_shot(KIND_CONV_AXIS, dtype, RL_HERMITIAN, ndim, shape, stride_in, stride_out, 1, axes, 1, data_in, data_out, fct,
                  stream, mul, mul_elems);
}

int impulse_fft_dct(int dtype, size_t ndim, const size_t *shape, const ptrdiff_t *stride_in,
                    const ptrdiff_t *stride_out, size_t naxes, const size_t *axes, int type, const void *data_in,
                    void *data_out, double fct, int ortho, size_t, void *stream) {
  if (type < 1 || type > 4) return fail(IMPULSE_FFT_ERR_INVALID, "invalid DCT type");  // hdronly.h:3288
  return one_shot(KIND_DCT, dtype, type, ndim, shape, stride_in, stride_out, naxes, axes, ortho != 0, data_in, data_out, fct, stream);
}
int impulse_fft_dst(int dtype, size_t ndim, const size_t *shape, const ptrdiff_t *stride_in,
                    const ptrdiff_t *stride_out, size_t naxes, const size_t *axes, int type, const void *data_in,
                    void *data_out, double fct, int ortho, size_t, void *stream) {
  if (type < 1 || type > 4) return fail(IMPULSE_FFT_ERR_INVALID, "invalid DST type");  // hdronly.h:3305
  return one_shot(KIND_DST, dtype, type, ndim, shape, stride_in, stride_out, naxes, axes, ortho != 0, data_in, data_out, fct, stream);
}

// pocketfft::r2r_fftpack / r2r_separable_hartley / r2r_genuine_hartley (pocketfft_hdronly.h:3392-3445)
int impulse_fft_r2r_fftpack(int dtype, size_t ndim, const size_t *shape, const ptrdiff_t *stride_in,
                            const ptrdiff_t *stride_out, size_t naxes, const size_t *axes, int real2hermitian, int forward,
                            const void *data_in, void *data_out, double fct, size_t, void *stream) {
  return one_shot(KIND_FFTPACK, dtype, real2hermitian != 0, ndim, shape, stride_in, stride_out, naxes, axes, forward != 0,
                  data_in, data_out, fct, stream);
}
int impulse_fft_r2r_separable_hartley(int dtype, size_t ndim, const size_t *shape, const ptrdiff_t *stride_in,
                                      const ptrdiff_t *stride_out, size_t naxes, const size_t *axes, const void *data_in,
                                      void *data_out, double fct, size_t, void *stream) {
  return one_shot(KIND_HARTLEY_SEP, dtype, 0, ndim, shape, stride_in, stride_out, naxes, axes, 1, data_in, data_out, fct, stream);
}
int impulse_fft_r2r_genuine_hartley(int dtype, size_t ndim, const size_t *shape, const ptrdiff_t *stride_in,
                                    const ptrdiff_t *stride_out, size_t naxes, const size_t *axes, const void *data_in,
                                    void *data_out, double fct, size_t, void *stream) {
  return one_shot(KIND_HARTLEY_GEN, dtype, 0, ndim, shape, stride_in, stride_out, naxes, axes, 1, data_in, data_out, fct, stream);
}

int impulse_fft_cfft_rows(double *data, size_t nrows, size_t length, int forward, double fct, void *stream) {
  if (length == 0) return fail(IMPULSE_FFT_ERR_INVALID, "zero-length transform");
  if (length == 1) fct = 1.0;  // pocketfft.c:875: a length-1 plan returns before scaling
  size_t shape[2] = {nrows, length}, axes[1] = {1};
  ptrdiff_t st[2] = {(ptrdiff_t)(length * 16), 16};
  return one_shot(KIND_C2C, DT_F64, RL_HERMITIAN, 2, shape, st, st, 1, axes, forward, data, data, fct, stream);
}
int impulse_fft_rfft_rows(double *data, size_t nrows, size_t length, int forward, double fct, void *stream) {
  if (length == 0) return fail(IMPULSE_FFT_ERR_INVALID, "zero-length transform");
  if (length == 1) fct = 1.0;  // pocketfft.c:1704: a length-1 plan returns before scaling
  size_t shape[2] = {nrows, length}, axes[1] = {1};
  ptrdiff_t st[2] = {(ptrdiff_t)(length * 8), 8};
  return one_shot(forward ? KIND_R2C : KIND_C2R, DT_F64, RL_HALFCOMPLEX, 2, shape, st, st, 1, axes, forward != 0, data, data,
                  fct, stream);
}

int impulse_fft_cmul(int dtype, const void *a, const void *filter, void *out, size_t n_inner, size_t n_batch,
                     double scale, void *stream) {
  DeviceCtx *ctx = nullptr;
  int rc = get_ctx(&ctx);
  if (rc) return rc;
  if (!a || !filter || !out) return fail(IMPULSE_FFT_ERR_INVALID, "null data pointer");
  if (!is_device_ptr(a) || !is_device_ptr(filter) || !is_device_ptr(out))
    return fail(IMPULSE_FFT_ERR_INVALID, "impulse_fft_cmul takes device pointers");
  int e = launch_cmul(dtype, a, filter, out, n_inner, n_batch, scale, ctx->sm_count, stream);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return e ? cuda_fail((cudaError_t)e, "kernel launch") : 0;
}

int impulse_fft_transpose(int dtype, const void *in, void *out, size_t rows, size_t cols, size_t ld_in, size_t ld_out,
                          size_t batch, void *stream) {
  DeviceCtx *ctx = nullptr;
  int rc = get_ctx(&ctx);
  if (rc) return rc;
  if (!in || !out) return fail(IMPULSE_FFT_ERR_INVALID, "null data pointer");
  if (in == out) return fail(IMPULSE_FFT_ERR_INVALID, "transpose is out of place");
  if (ld_in < cols || ld_out < rows) return fail(IMPULSE_FFT_ERR_STRIDE, "leading dimension smaller than the row length");
  if (!is_device_ptr(in) || !is_device_ptr(out)) return fail(IMPULSE_FFT_ERR_INVALID, "impulse_fft_transpose takes device pointers");
  int e = launch_transpose(dtype, in, out, rows, cols, ld_in, ld_out, batch, ctx->sm_count, stream);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return e ? cuda_fail((cudaError_t)e, "kernel launch") : 0;
}

int impulse_fft_copy2d(int dtype, const void *in, void *out, size_t rows, size_t cols, size_t ld_in, size_t ld_out,
                       size_t batch, size_t bs_in, size_t bs_out, void *stream) {
  DeviceCtx *ctx = nullptr;
  int rc = get_ctx(&ctx);
  if (rc) return rc;
  if (!in || !out) return fail(IMPULSE_FFT_ERR_INVALID, "null data pointer");
  if (ld_in < cols || ld_out < cols) return fail(IMPULSE_FFT_ERR_STRIDE, "leading dimension smaller than the row length");
  if (!is_device_ptr(in) || !is_device_ptr(out)) return fail(IMPULSE_FFT_ERR_INVALID, "impulse_fft_copy2d takes device pointers");
  int e = launch_copy2d(dtype, in, out, rows, cols, ld_in, ld_out, batch, bs_in, bs_out, ctx->sm_count, stream);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return e ? cuda_fail((cudaError_t)e, "kernel launch") : 0;
}

int impulse_fft_ipc_alloc(size_t bytes, void **ptr, void *handle64) {
  if (!ptr || !handle64 || !bytes) return fail(IMPULSE_FFT_ERR_INVALID, "null argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  cudaError_t e = cudaMalloc(ptr, bytes);
  if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc");
  cudaIpcMemHandle_t h;
  e = cudaIpcGetMemHandle(&h, *ptr);
  if (e != cudaSuccess) { cudaFree(*ptr); *ptr = nullptr; return cuda_fail(e, "cudaIpcGetMemHandle"); }
  std::memcpy(handle64, &h, 64);
  return 0;
}
int impulse_fft_ipc_free(void *ptr) {
  cudaError_t e = cudaFree(ptr);
  return e == cudaSuccess ? 0 : cuda_fail(e, "cudaFree");
}
int impulse_fft_ipc_open(const void *handle64, void **ptr) {
  if (!ptr || !handle64) return fail(IMPULSE_FFT_ERR_INVALID, "null argument");
  cudaIpcMemHandle_t h;
  std::memcpy(&h, handle64, 64);
  cudaError_t e = cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess);
  return e == cudaSuccess ? 0 : cuda_fail(e, "cudaIpcOpenMemHandle");
}
int impulse_fft_ipc_close(void *ptr) {
  cudaError_t e = cudaIpcCloseMemHandle(ptr);
  return e == cudaSuccess ? 0 : cuda_fail(e, "cudaIpcCloseMemHandle");
}

int impulse_fft_enable_peer_access(int peer_device) {
  int dev = -1;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return cuda_fail(e, "cudaGetDevice");
  if (peer_device == dev) return 0;
  int can = 0;
  e = cudaDeviceCanAccessPeer(&can, dev, peer_device);
  if (e != cudaSuccess) return cuda_fail(e, "cudaDeviceCanAccessPeer");
  if (!can) return fail(IMPULSE_FFT_ERR_UNSUPPORTED, "no peer access between these devices");
  e = cudaDeviceEnablePeerAccess(peer_device, 0);
  if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); return 0; }
  return e == cudaSuccess ? 0 : cuda_fail(e, "cudaDeviceEnablePeerAccess");
}

// Host-side NUMA placement.  Pinned staging memory is placed on the node of the thread that allocates it; with one
// process per GPU all ranks default to node 0 and their DMA traffic contends there (round 1: e2e efficiency 0.18 at
// 8 GPUs).  This binds the CALLING THREAD (and the threads it creates afterwards) to the CPUs of the NUMA node the
// device's PCIe root hangs off, read from sysfs.  No NUMA information (single node, container without sysfs) is
// not an error: *numa_node = -1 and nothing changes.
int impulse_fft_bind_host_to_device(int device, int *numa_node) {
  if (numa_node) *numa_node = -1;
  char bus[32] = {0};
  cudaError_t e = cudaDeviceGetPCIBusId(bus, (int)sizeof(bus), device);
  if (e != cudaSuccess) return cuda_fail(e, "cudaDeviceGetPCIBusId");
  for (char *c = bus; *c; ++c) if (*c >= 'A' && *c <= 'F') *c = (char)(*c - 'A' + 'a');
  std::string path = std::string("/sys/bus/pci/devices/") + bus + "/numa_node";
  FILE *f = std::fopen(path.c_str(), "r");
  if (!f) return 0;
  int node = -1;
  const int got = std::fscanf(f, "%d", &node);
  std::fclose(f);
  if (got != 1 || node < 0) return 0;
  path = "/sys/devices/system/node/node" + std::to_string(node) + "/cpulist";
  f = std::fopen(path.c_str(), "r");
  if (!f) return 0;
  char list[4096] = {0};
  const bool ok = std::fgets(list, sizeof(list), f) != nullptr;
  std::fclose(f);
  if (!ok) return 0;
  cpu_set_t set;
  CPU_ZERO(&set);
  int ncpu = 0;
  for (char *p = list; *p;) {            // "0-31,64-95"
    char *end = nullptr;
    long a = std::strtol(p, &end, 10);
    if (end == p) break;
    long b = a;
    if (*end == '-') { p = end + 1; b = std::strtol(p, &end, 10); }
    for (long c = a; c <= b && c < CPU_SETSIZE; ++c) { CPU_SET((int)c, &set); ++ncpu; }
    p = (*end == ',') ? end + 1 : end;
    if (*end != ',' ) break;
  }
  if (ncpu == 0) return 0;
  if (sched_setaffinity(0, sizeof(set), &set) != 0) return 0;   // not permitted (cgroup cpuset): leave as is
  if (numa_node) *numa_node = node;
  return 0;
}

int impulse_fft_gather_parts(int dtype, size_t nparts, const void *const *parts, size_t rows_per_part, size_t ld_part, size_t col0,
                             size_t ncols, void *out, size_t ld_out, int ctas, void *stream) {
  if (!parts || !out || nparts < 1 || nparts > 8 || !rows_per_part || !ncols)
    return fail(IMPULSE_FFT_ERR_INVALID, "bad argument (1 <= nparts <= 8)");
  if (ld_part < col0 + ncols || ld_out < ncols) return fail(IMPULSE_FFT_ERR_STRIDE, "leading dimension too small");
  DeviceCtx *ctx = nullptr;
  int rc = get_ctx(&ctx);
  if (rc) return rc;
  const size_t csz = dtype == DT_F64 ? 16 : 8, per16 = 16 / csz;   // elements per 16-byte unit
  if (ld_part % per16 || col0 % per16 || ncols % per16 || ld_out % per16 || (uintptr_t)out % 16)
    return fail(IMPULSE_FFT_ERR_STRIDE, "the gather moves 16-byte units: offsets and leading dimensions must be multiples of 16 bytes");
  for (size_t q = 0; q < nparts; ++q)
    if (!parts[q] || ((uintptr_t)parts[q] % 16)) return fail(IMPULSE_FFT_ERR_STRIDE, "part pointer null or misaligned");
  if (rows_per_part > 0xffffffffull || ncols / per16 > 0xffffffffull) return fail(IMPULSE_FFT_ERR_INVALID, "slab too large");
  int e = launch_gather_parts(parts, (uint32_t)nparts, (uint32_t)rows_per_part, ld_part / per16, col0 / per16,
                              (uint32_t)(ncols / per16), out, ld_out / per16, ctas, stream);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  if (e) return cuda_fail((cudaError_t)e, "kernel launch");
  return 0;
}

int impulse_fft_cols_from_parts(int dtype, size_t nparts, const void *const *parts, size_t rows_per_part, size_t ld_part,
                                size_t col0, size_t ncols, void *out, size_t ld_out, int forward, double fct, void *stream) {
  if (!parts || !out || nparts < 1 || nparts > 8 || !rows_per_part || !ncols)
    return fail(IMPULSE_FFT_ERR_INVALID, "bad argument (1 <= nparts <= 8)");
  if (ld_part < col0 + ncols || ld_out < ncols) return fail(IMPULSE_FFT_ERR_STRIDE, "leading dimension too small");
  const size_t csz = dtype == DT_F64 ? 16 : 8;
  const size_t R = nparts * rows_per_part;
  // plan the column transform of a virtual [R, ncols] array whose rows are ld_part apart, then point the
  // launches that read the input at the parts
  NdDesc d;
  d.kind = KIND_C2C; d.dtype = dtype; d.layout = RL_HERMITIAN; d.forward = forward != 0;
  d.shape = {R, ncols};
  d.stride_in = {(ptrdiff_t)(ld_part * csz), (ptrdiff_t)csz};
  d.stride_out = {(ptrdiff_t)(ld_out * csz), (ptrdiff_t)csz};
  d.axes = {0};
  d.no_col_whole = true;   // the whole-axis kernel does not read segmented input
  impulse_fft_plan raw = nullptr;
  int rc = create_plan(&raw, d);
  if (rc) return rc;
  std::unique_ptr<impulse_fft_plan_s> plan(raw);
  for (Step &st : plan->nd.steps) {
    if (st.src != BUF_IN) continue;
    LineJob &J = st.job;
    if (J.load_mode != LD_C || J.es_in <= 0 || (rows_per_part * ld_part) % (size_t)J.es_in)
      return fail(IMPULSE_FFT_ERR_UNSUPPORTED, "rows per part must be a multiple of the column split");
    J.seg_len = (uint32_t)((rows_per_part * ld_part) / (size_t)J.es_in);
    if (J.fast_id != FAST_NONE && !(J.fast_id >= COL2_64_F64 && J.fast_id <= COL2_512_F32)) J.fast_id = FAST_NONE;
    for (size_t q = 0; q < nparts; ++q) {
      if (!parts[q] || ((uintptr_t)parts[q] % csz)) return fail(IMPULSE_FFT_ERR_STRIDE, "part pointer null or misaligned");
      J.seg_base[q] = (const unsigned char *)parts[q] + col0 * csz;
    }
  }
  // the first part stands in for `in` (only its alignment is looked at; every load goes through seg_base)
  return run_device(plan.get(), (const unsigned char *)parts[0] + col0 * csz, out, fct, static_cast<cudaStream_t>(stream));
}

// ---- single-process multi-GPU driver (impulse_fft_dist_create / execute / destroy) -----------------
// The caller of the C ABI is ONE process (a Nim / C / C++ host), so the devices are driven from here: one stream per
// device, peer access between all pairs, events for the only cross-device dependency (row slabs complete -> column
// pass).  BATCH_SHARD has no communication at all; SLAB_2D exchanges through the column kernels' peer loads
// (impulse_fft_cols_from_parts) — there is no pack step and no all-to-all buffer.  The reference's analogue is
// general_nd handing line ranges to its thread pool (pocketfft_hdronly.h:3012-3050).
}  // extern "C"

struct impulse_fft_dist_s {
  int mode = 0;
  NdDesc desc;
  std::vector<int> devs;
  std::vector<cudaStream_t> streams;
  std::vector<cudaEvent_t> ev_rows;
  std::vector<impulse_fft_plan> plans;         // BATCH_SHARD: the shard's plan; SLAB_2D: the row pass of the slab
  std::vector<size_t> lo, hi;                  // rows of dimension 0 per device
  std::vector<void *> rowbuf;                  // SLAB_2D: row-FFT output [R/G, C], read by every peer
  std::vector<void *> stage_in, stage_out;     // SLAB_2D host path: device copies of the row slab / column slab
  size_t esz = 0;
};

namespace {
struct DeviceGuard {
  int prev = -1;
  DeviceGuard() { cudaGetDevice(&prev); }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

void dist_free(impulse_fft_dist h) {
  for (size_t g = 0; g < h->devs.size(); ++g) {
    if (cudaSetDevice(h->devs[g]) != cudaSuccess) { cudaGetLastError(); continue; }
    cudaDeviceSynchronize();
    if (g < h->plans.size() && h->plans[g]) delete h->plans[g];
    if (g < h->rowbuf.size() && h->rowbuf[g]) cudaFree(h->rowbuf[g]);
    if (g < h->stage_in.size() && h->stage_in[g]) cudaFree(h->stage_in[g]);
    if (g < h->stage_out.size() && h->stage_out[g]) cudaFree(h->stage_out[g]);
    if (g < h->ev_rows.size() && h->ev_rows[g]) cudaEventDestroy(h->ev_rows[g]);
    if (g < h->streams.size() && h->streams[g]) cudaStreamDestroy(h->streams[g]);
  }
  delete h;
}

// SLAB_2D on device pointers: in_parts[g] = row slab of device g (strides of the descriptor), out_parts[g] = dense
// column slab [R, C/G].  Asynchronous on the per-device streams; the caller synchronises.
int slab_run(impulse_fft_dist h, const void *const *in_parts, void *const *out_parts, double fct) {
  const size_t G = h->devs.size(), R = h->desc.shape[0], Cc = h->desc.shape[1], rl = R / G, cb = Cc / G;
  for (size_t g = 0; g < G; ++g) {
    cudaError_t e = cudaSetDevice(h->devs[g]);
    if (e != cudaSuccess) return cuda_fail(e, "cudaSetDevice");
    int rc = run_device(h->plans[g], in_parts[g], h->rowbuf[g], fct, h->streams[g]);
    if (rc) return rc;
    e = cudaEventRecord(h->ev_rows[g], h->streams[g]);
    if (e != cudaSuccess) return cuda_fail(e, "cudaEventRecord");
  }
  for (size_t g = 0; g < G; ++g) {
    cudaError_t e = cudaSetDevice(h->devs[g]);
    if (e != cudaSuccess) return cuda_fail(e, "cudaSetDevice");
    for (size_t q = 0; q < G; ++q)
      if (q != g) { e = cudaStreamWaitEvent(h->streams[g], h->ev_rows[q], 0); if (e != cudaSuccess) return cuda_fail(e, "cudaStreamWaitEvent"); }
    int rc = impulse_fft_cols_from_parts(h->desc.dtype, G, h->rowbuf.data(), rl, Cc, g * cb, cb, out_parts[g], cb,
                                         h->desc.forward ? 1 : 0, 1.0, h->streams[g]);
    if (rc) return rc;
  }
  return 0;
}

int dist_sync(impulse_fft_dist h) {
  int rc = 0;
  for (size_t g = 0; g < h->devs.size(); ++g) {
    cudaError_t e = cudaSetDevice(h->devs[g]);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->streams[g]);
    if (e != cudaSuccess && !rc) rc = cuda_fail(e, "multi-GPU synchronise");
  }
  return rc;
}
}  // namespace

extern "C" {

int impulse_fft_dist_create(impulse_fft_dist *out, int mode, const impulse_fft_desc *desc, int ndev, const int *devices) {
  if (!out || !desc || !devices) return fail(IMPULSE_FFT_ERR_INVALID, "null argument");
  *out = nullptr;
  if (ndev < 1 || ndev > 8) return fail(IMPULSE_FFT_ERR_INVALID, "1 <= ndev <= 8 (one NVSwitch box)");
  if (mode != IMPULSE_FFT_DIST_BATCH_SHARD && mode != IMPULSE_FFT_DIST_SLAB_2D) return fail(IMPULSE_FFT_ERR_INVALID, "unknown partitioning mode");
  if (desc->ndim > IMPULSE_FFT_MAX_DIMS || desc->naxes > IMPULSE_FFT_MAX_DIMS) return fail(IMPULSE_FFT_ERR_INVALID, "too many dimensions");
  for (int i = 0; i < ndev; ++i)
    for (int j = 0; j < i; ++j)
      if (devices[i] == devices[j]) return fail(IMPULSE_FFT_ERR_INVALID, "a device is listed twice");
  NdDesc d;
  int rc = make_desc(&d, desc->kind, desc->dtype, desc->real_layout, desc->forward, desc->ndim, desc->shape, desc->stride_in,
                     desc->stride_out, desc->naxes, desc->axes);
  if (rc) return rc;
  const size_t G = (size_t)ndev;
  if (mode == IMPULSE_FFT_DIST_BATCH_SHARD) {
    if (d.shape.size() < 2) return fail(IMPULSE_FFT_ERR_INVALID, "batch sharding needs a batch dimension (dimension 0)");
    for (size_t a : d.axes) if (a == 0) return fail(IMPULSE_FFT_ERR_INVALID, "dimension 0 is the sharded batch dimension: it cannot be a transform axis");
  } else {
    if (d.kind != KIND_C2C || d.shape.size() != 2 || d.axes.size() != 2 || d.axes[0] == d.axes[1] || d.axes[0] > 1 || d.axes[1] > 1)
      return fail(IMPULSE_FFT_ERR_INVALID, "the slab decomposition serves 2-D complex transforms over both axes");
    if (d.shape[0] % G || d.shape[1] % G) return fail(IMPULSE_FFT_ERR_INVALID, "rows and columns must be divisible by the number of devices");
  }
  DeviceGuard guard;
  std::unique_ptr<impulse_fft_dist_s, void (*)(impulse_fft_dist)> h(new impulse_fft_dist_s, dist_free);
  h->mode = mode; h->desc = d;
  h->devs.assign(devices, devices + ndev);
  h->esz = (d.dtype == DT_F64 ? 8 : 4) * 2;
  h->streams.assign(G, nullptr); h->ev_rows.assign(G, nullptr); h->plans.assign(G, nullptr);
  h->rowbuf.assign(G, nullptr); h->stage_in.assign(G, nullptr); h->stage_out.assign(G, nullptr);
  h->lo.resize(G); h->hi.resize(G);
  const size_t B = d.shape[0];
  for (size_t g = 0; g < G; ++g) { h->lo[g] = B * g / G; h->hi[g] = B * (g + 1) / G; }   // contiguous split
  for (size_t g = 0; g < G; ++g) {
    cudaError_t e = cudaSetDevice(h->devs[g]);
    if (e != cudaSuccess) return cuda_fail(e, "cudaSetDevice");
    e = cudaStreamCreateWithFlags(&h->streams[g], cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_rows[g], cudaEventDisableTiming);
    if (e != cudaSuccess) return cuda_fail(e, "stream / event creation");
    if (mode == IMPULSE_FFT_DIST_SLAB_2D) {
      for (size_t q = 0; q < G; ++q) {
        if (q == g) continue;
        int can = 0;
        e = cudaDeviceCanAccessPeer(&can, h->devs[g], h->devs[q]);
        if (e != cudaSuccess) return cuda_fail(e, "cudaDeviceCanAccessPeer");
        if (!can) return fail(IMPULSE_FFT_ERR_UNSUPPORTED, "no peer access between the listed devices");
        e = cudaDeviceEnablePeerAccess(h->devs[q], 0);
        if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); e = cudaSuccess; }
        if (e != cudaSuccess) return cuda_fail(e, "cudaDeviceEnablePeerAccess");
      }
    }
    NdDesc s = d;
    s.shape[0] = h->hi[g] - h->lo[g];
    if (mode == IMPULSE_FFT_DIST_SLAB_2D) {   // row pass of the slab into the dense published buffer
      s.axes = {1};
      s.stride_out = {(ptrdiff_t)(d.shape[1] * h->esz), (ptrdiff_t)h->esz};
      e = cudaMalloc(&h->rowbuf[g], s.shape[0] * d.shape[1] * h->esz);
      if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc (row slab)");
    } else if (d.kind == KIND_R2C || d.kind == KIND_C2R) {
      // (shape is the REAL array's shape on both sides: nothing else to adjust)
    }
    if (s.shape[0] == 0) continue;           // more devices than rows: this one idles
    rc = create_plan(&h->plans[g], s);
    if (rc) return rc;
  }
  *out = h.release();
  return 0;
}

int impulse_fft_dist_destroy(impulse_fft_dist h) {
  if (!h) return 0;
  DeviceGuard guard;
  dist_free(h);
  return 0;
}

int impulse_fft_dist_shard(impulse_fft_dist h, int index, size_t *lo, size_t *hi) {
  if (!h || index < 0 || (size_t)index >= h->devs.size() || !lo || !hi) return fail(IMPULSE_FFT_ERR_INVALID, "bad argument");
  *lo = h->lo[(size_t)index]; *hi = h->hi[(size_t)index];
  return 0;
}

int impulse_fft_dist_execute_parts(impulse_fft_dist h, const void *const *in_parts, void *const *out_parts, double fct) {
  if (!h || !in_parts || !out_parts) return fail(IMPULSE_FFT_ERR_INVALID, "null argument");
  DeviceGuard guard;
  const size_t G = h->devs.size();
  for (size_t g = 0; g < G; ++g)
    if (h->hi[g] > h->lo[g] && (!in_parts[g] || !out_parts[g])) return fail(IMPULSE_FFT_ERR_INVALID, "null shard pointer");
  int rc = 0;
  if (h->mode == IMPULSE_FFT_DIST_SLAB_2D) {
    rc = slab_run(h, in_parts, out_parts, fct);
  } else {
    for (size_t g = 0; g < G && !rc; ++g) {
      if (!h->plans[g]) continue;
      cudaError_t e = cudaSetDevice(h->devs[g]);
      if (e != cudaSuccess) { rc = cuda_fail(e, "cudaSetDevice"); break; }
      rc = impulse_fft_execute(h->plans[g], in_parts[g], out_parts[g], fct, h->streams[g]);
    }
  }
  const int rs = dist_sync(h);
  return rc ? rc : rs;
}

int impulse_fft_dist_execute(impulse_fft_dist h, const void *in, void *out, double fct) {
  if (!h || !in || !out) return fail(IMPULSE_FFT_ERR_INVALID, "null argument");
  if (is_device_ptr(in) || is_device_ptr(out))
    return fail(IMPULSE_FFT_ERR_INVALID, "impulse_fft_dist_execute takes HOST arrays (device shards: impulse_fft_dist_execute_parts)");
  DeviceGuard guard;
  const size_t G = h->devs.size();
  const NdDesc &d = h->desc;
  if (h->mode == IMPULSE_FFT_DIST_BATCH_SHARD) {
    // every device runs the ordinary host-pointer call on its shard (chunked H2D -> kernel -> D2H pipeline), one host
    // thread per device so that the PCIe links of all devices are busy at once
    std::vector<int> rcs(G, 0);
    std::vector<std::string> msgs(G);
    std::vector<std::thread> th;
    for (size_t g = 0; g < G; ++g) {
      if (!h->plans[g]) continue;
      th.emplace_back([&, g] {
        cudaError_t e = cudaSetDevice(h->devs[g]);
        if (e != cudaSuccess) { rcs[g] = IMPULSE_FFT_ERR_CUDA; msgs[g] = cudaGetErrorString(e); return; }
        int node = -1;
        impulse_fft_bind_host_to_device(h->devs[g], &node);   // staging threads run next to their GPU
        const unsigned char *pi = (const unsigned char *)in + (ptrdiff_t)h->lo[g] * d.stride_in[0];
        unsigned char *po = (unsigned char *)out + (ptrdiff_t)h->lo[g] * d.stride_out[0];
        rcs[g] = impulse_fft_execute(h->plans[g], pi, po, fct, nullptr);
        if (rcs[g]) msgs[g] = g_err;
      });
    }
    for (auto &t : th) t.join();
    for (size_t g = 0; g < G; ++g) if (rcs[g]) return fail(rcs[g], "device " + std::to_string(h->devs[g]) + ": " + msgs[g]);
    return 0;
  }
  // SLAB_2D from / to host arrays: H2D of each row slab, the two passes, D2H of each column slab into its columns of
  // the full output array (strided copy) — the result arrives in natural [R, C] layout
  const size_t R = d.shape[0], Cc = d.shape[1], rl = R / G, cb = Cc / G, esz = h->esz;
  if (d.stride_in[1] != (ptrdiff_t)esz || d.stride_out[1] != (ptrdiff_t)esz || d.stride_in[0] < (ptrdiff_t)(Cc * esz) ||
      d.stride_out[0] < (ptrdiff_t)(Cc * esz))
    return fail(IMPULSE_FFT_ERR_UNSUPPORTED, "the host path of the slab transform takes row-major arrays (unit stride along dimension 1)");
  std::vector<const void *> ins(G);
  std::vector<void *> outs(G);
  int rc = 0;
  for (size_t g = 0; g < G && !rc; ++g) {
    cudaError_t e = cudaSetDevice(h->devs[g]);
    if (e == cudaSuccess && !h->stage_in[g]) e = cudaMalloc(&h->stage_in[g], rl * Cc * esz);
    if (e == cudaSuccess && !h->stage_out[g]) e = cudaMalloc(&h->stage_out[g], R * cb * esz);
    if (e == cudaSuccess)
      e = cudaMemcpy2DAsync(h->stage_in[g], Cc * esz, (const unsigned char *)in + (ptrdiff_t)(g * rl) * d.stride_in[0], (size_t)d.stride_in[0],
                            Cc * esz, rl, cudaMemcpyHostToDevice, h->streams[g]);
    if (e != cudaSuccess) rc = cuda_fail(e, "slab H2D");
    ins[g] = h->stage_in[g]; outs[g] = h->stage_out[g];
  }
  if (!rc) {
    // the row plans were built for the descriptor's input strides: the staged slab is dense, so plan it that way
    for (size_t g = 0; g < G && !rc; ++g) {
      if (d.stride_in[0] == (ptrdiff_t)(Cc * esz)) continue;
      cudaSetDevice(h->devs[g]);
      NdDesc s = d;
      s.shape[0] = rl; s.axes = {1};
      s.stride_in = s.stride_out = {(ptrdiff_t)(Cc * esz), (ptrdiff_t)esz};
      impulse_fft_plan p = nullptr;
      rc = create_plan(&p, s);
      if (!rc) { delete h->plans[g]; h->plans[g] = p; }
    }
    if (!rc) { h->desc.stride_in = {(ptrdiff_t)(Cc * esz), (ptrdiff_t)esz}; }
  }
  if (!rc) rc = slab_run(h, ins.data(), outs.data(), fct);
  for (size_t g = 0; g < G && !rc; ++g) {
    cudaError_t e = cudaSetDevice(h->devs[g]);
    if (e == cudaSuccess)
      e = cudaMemcpy2DAsync((unsigned char *)out + g * cb * esz, (size_t)d.stride_out[0], h->stage_out[g], cb * esz, cb * esz, R,
                            cudaMemcpyDeviceToHost, h->streams[g]);
    if (e != cudaSuccess) rc = cuda_fail(e, "slab D2H");
  }
  const int rs = dist_sync(h);
  return rc ? rc : rs;
}

// ---- the ten pocketfft symbols (include/pocketfft.h) -------------------------
struct cfft_plan_i { size_t length; };
struct rfft_plan_i { size_t length; };

cfft_plan make_cfft_plan(size_t length) {
  if (length == 0) { g_err = "zero-length transform"; return nullptr; }
  DeviceCtx *ctx = nullptr;
  if (get_ctx(&ctx)) return nullptr;
  // plan the transform now, exactly as execute will (tables land in the per-device cache; lengths that run split
  // build only their sub-transform tables), so that a length this build cannot run fails HERE with a NULL plan
  if (length > 0xffffffffull) { g_err = "length exceeds 2^32-1"; return nullptr; }
  {
    NdDesc d;
    d.kind = KIND_C2C; d.dtype = DT_F64; d.layout = RL_HERMITIAN; d.forward = true;
    d.shape = {1, length};
    d.stride_in = d.stride_out = {(ptrdiff_t)(length * 16), 16};
    d.axes = {1};
    impulse_fft_plan raw = nullptr;
    if (create_plan(&raw, d)) return nullptr;
    delete raw;
  }
  cfft_plan p = new cfft_plan_i;
  p->length = length;
  return p;
}
void destroy_cfft_plan(cfft_plan plan) { delete plan; }
size_t cfft_length(cfft_plan plan) { return plan->length; }
int cfft_forward(cfft_plan plan, double c[], double fct) {
  return impulse_fft_cfft_rows(c, 1, plan->length, 1, fct, nullptr) ? -1 : 0;
}
int cfft_backward(cfft_plan plan, double c[], double fct) {
  return impulse_fft_cfft_rows(c, 1, plan->length, 0, fct, nullptr) ? -1 : 0;
}

rfft_plan make_rfft_plan(size_t length) {
  if (length == 0) { g_err = "zero-length transform"; return nullptr; }
  DeviceCtx *ctx = nullptr;
  if (get_ctx(&ctx)) return nullptr;
  if (length > 0xffffffffull) { g_err = "length exceeds 2^32-1"; return nullptr; }
  {
    NdDesc d;
    d.kind = KIND_R2C; d.dtype = DT_F64; d.layout = RL_HALFCOMPLEX; d.forward = true;
    d.shape = {1, length};
    d.stride_in = d.stride_out = {(ptrdiff_t)(length * 8), 8};
    d.axes = {1};
    impulse_fft_plan raw = nullptr;
    if (create_plan(&raw, d)) return nullptr;
    delete raw;
  }
  rfft_plan p = new rfft_plan_i;
  p->length = length;
  return p;
}
void destroy_rfft_plan(rfft_plan plan) { delete plan; }
size_t rfft_length(rfft_plan plan) { return plan->length; }
int rfft_forward(rfft_plan plan, double c[], double fct) {
  return impulse_fft_rfft_rows(c, 1, plan->length, 1, fct, nullptr) ? -1 : 0;
}
int rfft_backward(rfft_plan plan, double c[], double fct) {
  return impulse_fft_rfft_rows(c, 1, plan->length, 0, fct, nullptr) ? -1 : 0;
}

}  // extern "C"
