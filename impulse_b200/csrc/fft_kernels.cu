// sm_100a kernels of the generic engine: the CTA-level driver around the phase functions of
// fft_device.cuh, and the launcher used by abi.cu.
//
// Per tile of C lines: PROLOG (line offsets) | LOAD global->smem | radix passes and
// element-wise phases in smem | STORE smem->global, one __syncthreads() between phases.
// All HBM traffic is one coalesced read and one coalesced write per element; everything in
// between lives in shared memory (up to 227 KB per CTA on B200).
#include <cuda_runtime.h>
#include <cstdlib>

#include "fft_device.cuh"
#include "fft_kernels.h"

namespace impulse {

thread_local const char *g_last_kernel = "";

// BIG = tiles above half the SM's shared memory (one CTA per SM anyway): 512 threads, 128 registers.
template <typename T, bool BIG>
__global__ void __launch_bounds__(BIG ? kMaxThreadsBig : kMaxThreads, BIG ? 1 : kMinCtasPerSm)
line_fft_kernel(const __grid_constant__ LineJob J) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  int64_t *offs = reinterpret_cast<int64_t *>(smem_raw);
  cx<T> *smem = reinterpret_cast<cx<T> *>(smem_raw + kSmemHeaderBytes);
  const uint32_t tid = threadIdx.x, nthr = blockDim.x;
  const uint64_t n_tiles = (J.n_lines + (1ull << J.log_c) - 1) >> J.log_c;
  for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const TileCtx tc = tile_ctx(J, tile);
    phase_prolog(J, tc, tid, offs);
    __syncthreads();
    phase_load<T>(J, tc, tid, nthr, offs, smem);
    __syncthreads();
    for (int p = 0; p < J.nphases; ++p) {
      phase_mid<T>(J, J.ph[p], tid, nthr, smem);
      __syncthreads();
    }
    phase_store<T>(J, tc, tid, nthr, offs, smem);
    __syncthreads();
  }
}

int configure_kernels(size_t max_dyn_smem) {
  const void *ks[4] = {(const void *)line_fft_kernel<double, false>, (const void *)line_fft_kernel<double, true>,
                       (const void *)line_fft_kernel<float, false>, (const void *)line_fft_kernel<float, true>};
  for (const void *k : ks) {
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_dyn_smem);
    if (e != cudaSuccess) return (int)e;
    // without this the driver may pick a carve-out that fits one CTA only
    e = cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return (int)e;
  }
  return 0;
}

int launch_line_job(const LineJob &J, int threads, size_t smem_bytes, uint64_t n_tiles, void *stream) {
  if (n_tiles == 0) return 0;
  if (J.fast_id != FAST_NONE) {
    static int sm_counts[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    int &sm_count = sm_counts[(dev >= 0 && dev < 64) ? dev : 0];
    if (!sm_count) cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
    return launch_fast_job(J, sm_count, stream);
  }
  const unsigned grid = (unsigned)(n_tiles < 0x7fffffffull ? n_tiles : 0x7fffffffull);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  g_last_kernel = J.dtype == 1 ? "line_fft_kernel<double>" : "line_fft_kernel<float>";
  const bool big = threads > kMaxThreads;
  {
    // The radix passes read their twiddles through L1.  With the carve-out pinned at "all shared memory" L1
    // shrinks to its minimum and every twiddle becomes an L2 round trip (ncu: long-scoreboard stalls dominate);
    // ask only for the shared memory the resident CTAs need and leave the rest to L1.
    static const int mode = [] { const char *e = getenv("IMPULSE_FFT_CARVEOUT"); return e ? atoi(e) : 0; }();  // 0 = adaptive, else percent
    const size_t per_cta = smem_bytes + 1024;
    size_t ctas = big ? 1 : (size_t)kMinCtasPerSm;
    while (ctas > 1 && ctas * per_cta > 227 * 1024) --ctas;
    int pct = mode > 0 ? mode : (int)((ctas * per_cta * 100 + 228 * 1024 - 1) / (228 * 1024));
    if (pct > 100) pct = 100;
    static int last_pct[64][4] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    const int kidx = (J.dtype == 1 ? 0 : 2) + (big ? 1 : 0);
    int &lp = last_pct[(dev >= 0 && dev < 64) ? dev : 0][kidx];
    if (lp != pct + 1) {
      const void *k = J.dtype == 1 ? (big ? (const void *)line_fft_kernel<double, true> : (const void *)line_fft_kernel<double, false>)
                                   : (big ? (const void *)line_fft_kernel<float, true> : (const void *)line_fft_kernel<float, false>);
      cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
      if (e != cudaSuccess) return (int)e;
      lp = pct + 1;
    }
  }
  if (J.dtype == 1) {
    if (big) line_fft_kernel<double, true><<<grid, threads, smem_bytes, s>>>(J);
    else line_fft_kernel<double, false><<<grid, threads, smem_bytes, s>>>(J);
  } else {
    if (big) line_fft_kernel<float, true><<<grid, threads, smem_bytes, s>>>(J);
    else line_fft_kernel<float, false><<<grid, threads, smem_bytes, s>>>(J);
  }
  return (int)cudaGetLastError();
}

}  // namespace impulse
