// Micro-test (tools, not product): tensor memory as lane-private staging storage for a non-MMA kernel.
// 512 threads; warp w owns lanes 32*(w%4)..+31 and columns 64*(w/4)..+63 of a 256-column allocation; every thread
// stores 64 words, reads them back, checks them.  Then a timing loop.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n" ::"r"(taddr), "r"(v[0]), "r"(v[1]),
               "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]));
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr));
}
__global__ void __launch_bounds__(512, 1) tmem_test(unsigned *errors, unsigned long long *cycles, int iters) {
  __shared__ uint32_t slot;
  const int t = threadIdx.x, warp = t >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)), "n"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n");
  const uint32_t base = slot;
  const uint32_t my = base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(warp >> 2) * 64;
  unsigned bad = 0;
  for (int rep = 0; rep < 2; ++rep) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      uint32_t v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = (uint32_t)(blockIdx.x * 7919 + t * 1000 + i * 8 + k + rep * 77);
      tmem_st8(my + 8 * i, v);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      uint32_t v[8];
      tmem_ld8(my + 8 * i, v);
      asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
      for (int k = 0; k < 8; ++k) bad += v[k] != (uint32_t)(blockIdx.x * 7919 + t * 1000 + i * 8 + k + rep * 77);
    }
    __syncthreads();
  }
  if (bad) atomicAdd(errors, bad);
  // timing: 64 words out and back per iteration
  uint32_t acc[8] = {1, 2, 3, 4, 5, 6, 7, 8};
  const long long c0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) tmem_st8(my + 8 * i, acc);
    asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      uint32_t v[8];
      tmem_ld8(my + 8 * i, v);
      asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] += v[k];
    }
  }
  const long long c1 = clock64();
  if (t == 0 && blockIdx.x == 0) { cycles[0] = (unsigned long long)(c1 - c0); cycles[1] = acc[0]; }
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(base), "n"(256));
}
int main() {
  unsigned *err; unsigned long long *cyc;
  cudaMalloc(&err, 4); cudaMemset(err, 0, 4); cudaMalloc(&cyc, 16);
  const int iters = 1000;
  tmem_test<<<296, 512>>>(err, cyc, iters);   // two waves: the second wave's alloc needs the first wave's dealloc
  cudaError_t e = cudaDeviceSynchronize();
  unsigned h = 0; unsigned long long hc[2] = {0, 0};
  cudaMemcpy(&h, err, 4, cudaMemcpyDeviceToHost); cudaMemcpy(hc, cyc, 16, cudaMemcpyDeviceToHost);
  printf("tmem staging test: %s, mismatches %u, %.1f cycles per 64-word store+load round trip per thread (512 threads)\n",
         cudaGetErrorString(e), h, (double)hc[0] / iters);
  return h != 0 || e != cudaSuccess;
}
