// Fused column transform: BOTH launches of the four-step split of a strided axis (N = N1*N2) in ONE persistent
// kernel, with the intermediate array kept in L2.
//
// The split of a column transform is   A: for every n2, N1-point FFTs over n1 (rows n1*N2 + n2), times W_N^(k1*n2)
//                                      B: for every k1, N2-point FFTs over n2 (rows k1*N2 + n2) -> rows k1 + N1*k2.
// As two launches the intermediate (the whole array) goes to HBM and comes back: 4 passes of traffic for one axis.
// Here a "tile" is one group of LPC adjacent columns x all N rows (N*LPC elements: 2 MB for 8192 x 16 complex128).
// Work items are dealt out IN ORDER from one queue:   A(tile 0) | A(1) B(0) | A(2) B(1) | ... | B(T-1)
// so a B item only ever waits for A items that were claimed at least one tile earlier (already finished in
// practice; a completion counter per tile makes it correct), and A writes its tile into a small RING of slots
// (RING tiles of scratch, a few MB) that B reads back a few microseconds later: the intermediate lives in the
// 126 MB L2 and is overwritten there before it is ever evicted.  HBM sees one read and one write of the array.
//
// Thread mapping, shared-memory layout and arithmetic of each item are those of colfast2_kernel (col_device.cuh).
#pragma once
#include "col_device.cuh"

namespace impulse {

struct FuseJob {
  LineJob A, B;             // the two launches exactly as the planner built them (A.in = source, B.out = destination;
                            // their scratch-side pointers/strides are ignored: the ring is addressed below)
  void *ring;               // [ring_slots][N][LPC] complex
  unsigned int *ctrl;       // [0] next item; [2 + c] A items of tile c done; [2 + tiles + c] B items of tile c done
  uint32_t ring_slots, tiles, g0n, itemsA, itemsB;   // tiles = g0n * bdim[2]; itemsA = N2, itemsB = N1
};

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int *p) {
  unsigned int v;
#if defined(__CUDA_ARCH__)
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
#else
  v = __atomic_load_n(p, __ATOMIC_ACQUIRE);
#endif
  return v;
}
template <typename T> __device__ __forceinline__ cx<T> ld_l2(const cx<T> *p) {   // L2 only: the ring is rewritten by other SMs
#if defined(__CUDA_ARCH__)
  return __ldcg(p);
#else
  return *p;
#endif
}
template <typename T> __device__ __forceinline__ void st_l2(cx<T> *p, cx<T> v) {
#if defined(__CUDA_ARCH__)
  __stcg(p, v);
#else
  *p = v;
#endif
}

// one group of LPC lines: N = R1*R2 points, two passes through the exchange buffer S
//   in[n*es_in], out[k*es_out] are this thread's line;  IN_L2 / OUT_L2: that side is the ring
template <typename T, int R1, int R2, int LPC, bool BWD, bool IN_L2, bool OUT_L2>
__device__ __forceinline__ void colfuse_item(const cx<T> *in, int64_t es_in, cx<T> *out, int64_t es_out, bool valid, const cx<T> *tw,
                                             const LineJob &J, bool tw4, uint32_t twi, T f, cx<T> *S, int line, int i) {
  constexpr int NB2 = R1 / R2;
  cx<T> x[R1];
#pragma unroll
  for (int j = 0; j < R1; ++j) {
    const cx<T> *p = in + (int64_t)(i + R2 * j) * es_in;
    x[j] = valid ? (IN_L2 ? ld_l2<T>(p) : *p) : mk<T>((T)0, (T)0);
    if (BWD) x[j].y = -x[j].y;
  }
  RegFFT<T, R1>::run(x);
#pragma unroll
  for (int k = 1; k < R1; ++k) x[k] = cmul(x[k], __ldg(tw + i * k));
#pragma unroll
  for (int k = 0; k < R1; ++k) S[(k * R2 + i) * LPC + line] = x[k];
  __syncthreads();
  cx<T> wstep = mk<T>((T)1, (T)0);
  if (tw4) wstep = four_step_w<T>(J, (uint32_t)R1 * twi);
#pragma unroll
  for (int m = 0; m < NB2; ++m) {
    const int k1 = i + R2 * m;
    cx<T> y[R2];
#pragma unroll
    for (int j = 0; j < R2; ++j) y[j] = S[(k1 * R2 + j) * LPC + line];
    RegFFT<T, R2>::run(y);
    cx<T> w = mk<T>((T)1, (T)0);
    if (tw4) w = four_step_w<T>(J, (uint32_t)k1 * twi);
#pragma unroll
    for (int k2 = 0; k2 < R2; ++k2) {
      const int k = k1 + R1 * k2;
      cx<T> v = y[k2];
      if (tw4) { v = cmul(v, w); w = cmul(w, wstep); }
      v.x *= f;
      v.y *= BWD ? -f : f;
      if (valid) {
        if (OUT_L2) st_l2<T>(out + (int64_t)k * es_out, v);
        else out[(int64_t)k * es_out] = v;
      }
    }
  }
}

template <typename T, int RA1, int RA2, int RB1, int RB2, int LPC, bool BWD>
__global__ void __launch_bounds__(LPC * RA2)
colfuse2_kernel(const __grid_constant__ FuseJob F) {
  static_assert(RA2 == RB2, "both phases run on the same threads");
  constexpr int NA = RA1 * RA2, NB = RB1 * RB2, NMAX = NA > NB ? NA : NB;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cx<T> *S = reinterpret_cast<cx<T> *>(smem_raw);
  unsigned int *s_item = reinterpret_cast<unsigned int *>(S + (size_t)NMAX * LPC);
  const int u = threadIdx.x, line = u % LPC, i = u / LPC;
  const uint32_t per_round = F.itemsA + F.itemsB;
  const uint32_t total = F.tiles * per_round;
  unsigned int *doneA = F.ctrl + 2, *doneB = F.ctrl + 2 + F.tiles;
  const uint64_t slot_elems = (uint64_t)NA * NB * LPC;
  for (;;) {
    if (u == 0) *s_item = atomicAdd(&F.ctrl[0], 1u);
    __syncthreads();
    const uint32_t item = *s_item;
    if (item >= total) break;
    // queue order: round 0 = A(0); round r = A(r), B(r-1); round T = B(T-1)
    bool phaseA;
    uint32_t c, q;
    if (item < F.itemsA) { phaseA = true; c = 0; q = item; }
    else {
      const uint32_t s = item - F.itemsA, r = 1 + s / per_round;
      q = s - (r - 1) * per_round;
      if (r < F.tiles && q < F.itemsA) { phaseA = true; c = r; }
      else { phaseA = false; c = r - 1; q = (r < F.tiles) ? q - F.itemsA : q; }
    }
    const uint32_t g0 = c % F.g0n, i2 = c / F.g0n, slot = c % F.ring_slots;
    if (u == 0) {   // dependencies (satisfied long ago in steady state: items are claimed in order)
      if (phaseA) { if (c >= F.ring_slots) while (ld_acquire_u32(&doneB[c - F.ring_slots]) < F.itemsB) {} }
      else while (ld_acquire_u32(&doneA[c]) < F.itemsA) {}
    }
    __syncthreads();
    cx<T> *ring = reinterpret_cast<cx<T> *>(F.ring) + (uint64_t)slot * slot_elems + line;
    const uint32_t l0 = g0 * LPC + line;
    if (phaseA) {
      // group (g0, n2 = q, i2): N1-point FFT over n1 of src rows n1*N2 + n2; output k1 -> ring row k1*N2 + n2
      const LineJob &J = F.A;
      const bool valid = l0 < (uint32_t)J.bdim[0];
      const cx<T> *in = reinterpret_cast<const cx<T> *>(J.in) + (int64_t)l0 + (int64_t)q * J.bs_in[1] + (int64_t)i2 * J.bs_in[2];
      const uint32_t twi = J.tw4_dim == 0 ? l0 : J.tw4_dim == 1 ? q : J.tw4_dim == 2 ? i2 : 0u;
      colfuse_item<T, RA1, RA2, LPC, BWD, false, true>(in, J.es_in, ring + (uint64_t)q * LPC, (int64_t)NB * LPC, valid,
                                                       reinterpret_cast<const cx<T> *>(J.tw), J, J.tw4_n != 0, twi, (T)J.fct, S, line, i);
    } else {
      // group (g0, k1 = q, i2): N2-point FFT over n2 of ring rows k1*N2 + n2; output k2 -> dst element k1 + N1*k2
      const LineJob &J = F.B;
      const bool valid = l0 < (uint32_t)J.bdim[0];
      cx<T> *out = reinterpret_cast<cx<T> *>(J.out) + (int64_t)l0 + (int64_t)q * J.bs_out[1] + (int64_t)i2 * J.bs_out[2];
      colfuse_item<T, RB1, RB2, LPC, BWD, true, false>(ring + (uint64_t)q * NB * LPC, (int64_t)LPC, out, J.es_out, valid,
                                                       reinterpret_cast<const cx<T> *>(J.tw), J, false, 0u, (T)J.fct, S, line, i);
    }
    __threadfence();            // this thread's ring / output stores are visible device-wide ...
    __syncthreads();            // ... for all threads, and the exchange buffer and s_item are free again
    if (u == 0) atomicAdd(phaseA ? &doneA[c] : &doneB[c], 1u);
  }
}

}  // namespace impulse
