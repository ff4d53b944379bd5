// Tensor memory (256 KB per SM, 128 lanes x 512 columns x 32 bit) as LANE-PRIVATE storage for kernels that issue no MMA:
// warp w owns lanes 32*(w%4)..+31, and a thread reads back exactly the columns of its own lane that it stored
// (tcgen05.st / tcgen05.ld, .32x32b: one 32-bit column per register).  tools/micro/tmem_stage.cu validated the addressing
// (0 mismatches over two waves of alloc / dealloc; 64 words out and back in 460 cycles with 512 threads).
// The host emulations (tests/emu) give every thread a private array.
#pragma once
#include <stdint.h>

namespace impulse {

#if !defined(__CUDA_ARCH__)
static thread_local uint32_t cw_emu_tmem[512];
#endif
__device__ __forceinline__ void cw_tmem_st8(uint32_t taddr, const uint32_t *v) {
#if defined(__CUDA_ARCH__)
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n" ::"r"(taddr), "r"(v[0]), "r"(v[1]),
               "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]));
#else
  for (int k = 0; k < 8; ++k) cw_emu_tmem[(taddr & 0xffffu) + k] = v[k];
#endif
}
__device__ __forceinline__ void cw_tmem_ld8(uint32_t taddr, uint32_t *v) {
#if defined(__CUDA_ARCH__)
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr));
#else
  for (int k = 0; k < 8; ++k) v[k] = cw_emu_tmem[(taddr & 0xffffu) + k];
#endif
}
__device__ __forceinline__ void cw_tmem_wait_st() {
#if defined(__CUDA_ARCH__)
  asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
#endif
}
__device__ __forceinline__ void cw_tmem_wait_ld() {
#if defined(__CUDA_ARCH__)
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#endif
}
// one warp (threads 0..31) allocates COLS columns and publishes the base address through `slot` (shared memory); every
// thread of the CTA calls this; returns the base.  tmem_release: every thread, after its last tensor-memory access.
template <int COLS> __device__ __forceinline__ uint32_t cw_tmem_acquire(uint32_t *slot, int t) {
#if defined(__CUDA_ARCH__)
  if (t < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"((uint32_t)__cvta_generic_to_shared(slot)), "n"(COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n");
  return *slot;
#else
  (void)slot; (void)t;
  return 0u;
#endif
}
template <int COLS> __device__ __forceinline__ void cw_tmem_release(uint32_t base, int t) {
#if defined(__CUDA_ARCH__)
  asm volatile("tcgen05.fence::before_thread_sync;\n");
  __syncthreads();
  if (t < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(base), "n"(COLS));
#else
  (void)base; (void)t;
#endif
}

}  // namespace impulse
