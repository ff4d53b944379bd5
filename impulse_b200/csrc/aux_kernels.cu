// Auxiliary HBM-bound kernels that sit between FFT passes:
//   * cmul       — spectrum[b][i] *= filter[i] * scale   (FFT-based filtering, BASELINE config 5:
//                  r2c -> pointwise multiply -> c2r)
//   * transpose  — batched tiled shared-memory transpose of complex matrices (slab fft2:
//                  re-blocking around the all-to-all)
// Both are pure streaming kernels: 128-bit accesses, grid sized in multiples of the SM count.
#include <cuda_runtime.h>

#include "fft_device.cuh"
#include "fft_kernels.h"

namespace impulse {

template <typename T>
__global__ void __launch_bounds__(256) cmul_kernel(const cx<T> *__restrict__ a, const cx<T> *__restrict__ f,
                                                   cx<T> *__restrict__ out, size_t n_inner, size_t n_batch, T scale) {
  const size_t total = n_inner * n_batch;
  for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (size_t)gridDim.x * blockDim.x) {
    const size_t i = g % n_inner;
    cx<T> v = cmul(a[g], __ldg(f + i));
    v.x *= scale; v.y *= scale;
    out[g] = v;
  }
}

template <typename T>
__global__ void __launch_bounds__(256) transpose_kernel(const cx<T> *__restrict__ in, cx<T> *__restrict__ out,
                                                        size_t rows, size_t cols, size_t ld_in, size_t ld_out,
                                                        size_t tiles_r, size_t tiles_c, size_t n_tiles) {
  __shared__ cx<T> tile[32][33];
  const unsigned tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (size_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    const size_t b = t / (tiles_r * tiles_c), rem = t % (tiles_r * tiles_c);
    const size_t tr = rem / tiles_c, tc = rem % tiles_c;
    const cx<T> *src = in + b * rows * ld_in;
    cx<T> *dst = out + b * cols * ld_out;
#pragma unroll
    for (unsigned k = 0; k < 32; k += 8) {
      const size_t r = tr * 32 + ty + k, c = tc * 32 + tx;
      if (r < rows && c < cols) tile[ty + k][tx] = src[r * ld_in + c];
    }
    __syncthreads();
#pragma unroll
    for (unsigned k = 0; k < 32; k += 8) {
      const size_t c = tc * 32 + ty + k, r = tr * 32 + tx;
      if (r < rows && c < cols) dst[c * ld_out + r] = tile[tx][ty + k];
    }
    __syncthreads();
  }
}

// batched strided matrix copy: out[b][r][c] = in[b][r][c] with independent leading dimensions and
// batch strides (elements).  Packs the per-peer blocks of a slab into contiguous send buffers.
template <typename T>
__global__ void __launch_bounds__(256) copy2d_kernel(const cx<T> *__restrict__ in, cx<T> *__restrict__ out, size_t rows,
                                                     size_t cols, size_t ld_in, size_t ld_out, size_t batch,
                                                     size_t bs_in, size_t bs_out) {
  const size_t per = rows * cols, total = per * batch;
  for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (size_t)gridDim.x * blockDim.x) {
    const size_t b = g / per, rem = g - b * per;
    const size_t r = rem / cols, cc = rem - r * cols;
    out[b * bs_out + r * ld_out + cc] = in[b * bs_in + r * ld_in + cc];
  }
}

int launch_copy2d(int dtype, const void *in, void *out, size_t rows, size_t cols, size_t ld_in, size_t ld_out,
                  size_t batch, size_t bs_in, size_t bs_out, int sm_count, void *stream) {
  const size_t total = rows * cols * batch;
  if (!total) return 0;
  size_t blocks = (total + 255) / 256;
  const size_t cap = (size_t)sm_count * 16;
  if (blocks > cap) blocks = cap;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (dtype == 1) copy2d_kernel<double><<<(unsigned)blocks, 256, 0, s>>>((const cx<double> *)in, (cx<double> *)out, rows, cols, ld_in, ld_out, batch, bs_in, bs_out);
  else copy2d_kernel<float><<<(unsigned)blocks, 256, 0, s>>>((const cx<float> *)in, (cx<float> *)out, rows, cols, ld_in, ld_out, batch, bs_in, bs_out);
  return (int)cudaGetLastError();
}

// Gather of the column block [col0, col0 + ncols) out of `nparts` row slabs (the peers' row-FFT outputs, mapped through
// CUDA IPC) into one local [nparts * rpp, ncols] array: the exchange of the slab 2-D transform as a COPY with deep
// memory-level parallelism (U independent 16-byte loads in flight per thread: NVLink latency is 2-4 k cycles) on a
// deliberately small grid, so that the column transform of the previous chunk keeps the other SMs.
struct GatherParts { const void *part[8]; };
// one warp per row (U rows in flight): lanes walk the row's 16-byte units, so every load instruction is one coalesced
// 512-byte run and the index arithmetic is one 32-bit division per ROW
template <int U>
__global__ void __launch_bounds__(512) gather_parts_kernel(GatherParts P, uint4 *__restrict__ out, uint32_t nparts, uint32_t rpp,
                                                           uint64_t ld_part16, uint64_t col0_16, uint32_t ncols16, uint64_t ld_out16) {
  const uint32_t lane = threadIdx.x & 31, warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
  const uint32_t rows = nparts * rpp;
  for (uint32_t r0 = warp; r0 < rows; r0 += nwarps * U) {
    const uint4 *src[U];
    uint4 *dst[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint32_t r = r0 + nwarps * u, rc = r < rows ? r : r0, q = rc / rpp;
      src[u] = reinterpret_cast<const uint4 *>(P.part[q]) + (uint64_t)(rc - q * rpp) * ld_part16 + col0_16;
      dst[u] = out + (uint64_t)rc * ld_out16;
    }
    for (uint32_t c = lane; c < ncols16; c += 32) {
      uint4 v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) v[u] = src[u][c];
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (r0 + nwarps * u < rows) dst[u][c] = v[u];
    }
  }
}

// all quantities in 16-byte units (the ABI checks divisibility); ctas = grid size (0: 32)
int launch_gather_parts(const void *const *parts, uint32_t nparts, uint32_t rpp, uint64_t ld_part16, uint64_t col0_16,
                        uint32_t ncols16, void *out, uint64_t ld_out16, int ctas, void *stream) {
  if (!nparts || !rpp || !ncols16) return 0;
  GatherParts P;
  for (uint32_t q = 0; q < 8; ++q) P.part[q] = q < nparts ? parts[q] : nullptr;
  gather_parts_kernel<8><<<(unsigned)(ctas > 0 ? ctas : 32), 512, 0, static_cast<cudaStream_t>(stream)>>>(
      P, static_cast<uint4 *>(out), nparts, rpp, ld_part16, col0_16, ncols16, ld_out16);
  return (int)cudaGetLastError();
}

int launch_cmul(int dtype, const void *a, const void *f, void *out, size_t n_inner, size_t n_batch, double scale,
                int sm_count, void *stream) {
  const size_t total = n_inner * n_batch;
  if (!total) return 0;
  size_t blocks = (total + 255) / 256;
  const size_t cap = (size_t)sm_count * 16;
  if (blocks > cap) blocks = cap;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (dtype == 1) cmul_kernel<double><<<(unsigned)blocks, 256, 0, s>>>((const cx<double> *)a, (const cx<double> *)f, (cx<double> *)out, n_inner, n_batch, scale);
  else cmul_kernel<float><<<(unsigned)blocks, 256, 0, s>>>((const cx<float> *)a, (const cx<float> *)f, (cx<float> *)out, n_inner, n_batch, (float)scale);
  return (int)cudaGetLastError();
}

int launch_transpose(int dtype, const void *in, void *out, size_t rows, size_t cols, size_t ld_in, size_t ld_out,
                     size_t batch, int sm_count, void *stream) {
  if (!rows || !cols || !batch) return 0;
  const size_t tr = (rows + 31) / 32, tc = (cols + 31) / 32, nt = tr * tc * batch;
  size_t blocks = nt;
  const size_t cap = (size_t)sm_count * 32;
  if (blocks > cap) blocks = cap;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (dtype == 1) transpose_kernel<double><<<(unsigned)blocks, 256, 0, s>>>((const cx<double> *)in, (cx<double> *)out, rows, cols, ld_in, ld_out, tr, tc, nt);
  else transpose_kernel<float><<<(unsigned)blocks, 256, 0, s>>>((const cx<float> *)in, (cx<float> *)out, rows, cols, ld_in, ld_out, tr, tc, nt);
  return (int)cudaGetLastError();
}

template <typename T>
__global__ void __launch_bounds__(256) hartley_combine_kernel(const __grid_constant__ CombineJob C) {
  for (uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; g < C.total; g += (uint64_t)gridDim.x * blockDim.x)
    hartley_combine_one<T>(C, g);
}

int launch_hartley_combine(const CombineJob &job, int sm_count, void *stream) {
  if (job.total == 0) return 0;
  const uint64_t want = (job.total + 255) / 256, cap = (uint64_t)sm_count * 16;
  const unsigned grid = (unsigned)(want < cap ? want : cap);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (job.dtype == 1) hartley_combine_kernel<double><<<grid, 256, 0, s>>>(job);
  else hartley_combine_kernel<float><<<grid, 256, 0, s>>>(job);
  g_last_kernel = "hartley_combine_kernel";
  return (int)cudaGetLastError();
}

template <typename T>
__global__ void __launch_bounds__(256) aux_kernel(const __grid_constant__ AuxJob A) {
  for (uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; g < A.total; g += (uint64_t)gridDim.x * blockDim.x)
    aux_one<T>(A, g);
}

int launch_aux(const AuxJob &job, int sm_count, void *stream) {
  if (job.total == 0) return 0;
  const uint64_t want = (job.total + 255) / 256, cap = (uint64_t)sm_count * 16;
  const unsigned grid = (unsigned)(want < cap ? want : cap);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (job.dtype == 1) aux_kernel<double><<<grid, 256, 0, s>>>(job);
  else aux_kernel<float><<<grid, 256, 0, s>>>(job);
  g_last_kernel = "aux_kernel";
  return (int)cudaGetLastError();
}

}  // namespace impulse
