// TMA helpers: 1-D bulk copy global -> shared (cp.async.bulk) completing on an mbarrier, shared by the two-pass
// row kernel (fast_kernels.cu) and the staged three-pass kernels (fast3_device.cuh).
// Under a host compiler (tests/emu: one OS thread per CUDA thread) the same names are emulated with a memcpy
// and an atomic phase word, so that the staged kernels run unchanged in the thread-level emulation.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
namespace impulse {
// Programmatic dependent launch: a kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may begin
// (prologue: barrier init, table loads) while the previous kernel of the stream drains; it must not touch data the
// previous kernel may still write before griddep_wait().  griddep_launch_dependents() lets the NEXT kernel do the same.
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
}  // namespace impulse
#else
#include <atomic>
#include <cstring>
#include <thread>
namespace impulse {
inline void griddep_wait() {}
inline void griddep_launch_dependents() {}
// emulated mbarrier word: low 32 bits = transaction bytes still pending, high 32 bits = completed phases
inline void mbar_init(uint64_t *bar, uint32_t) { __atomic_store_n(bar, 0ull, __ATOMIC_SEQ_CST); }
inline void mbar_init_fence() {}
inline void mbar_expect_tx(uint64_t *bar, uint32_t bytes) { __atomic_fetch_add(bar, (uint64_t)bytes, __ATOMIC_SEQ_CST); }
inline void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  std::memcpy(dst, src, bytes);
  const uint64_t after = __atomic_sub_fetch(bar, (uint64_t)bytes, __ATOMIC_SEQ_CST);
  if ((uint32_t)after == 0u) __atomic_fetch_add(bar, 1ull << 32, __ATOMIC_SEQ_CST);   // phase complete
}
inline void mbar_wait(uint64_t *bar, uint32_t parity) {
  while ((((uint32_t)(__atomic_load_n(bar, __ATOMIC_SEQ_CST) >> 32)) & 1u) == parity) std::this_thread::yield();
}
inline void fence_proxy_async() {}
}  // namespace impulse
#endif
