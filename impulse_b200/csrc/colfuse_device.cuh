// Fused column transform: BOTH launches of the four-step split of a strided axis (N = N1*N2) in ONE persistent
// kernel, with the intermediate array kept in L2.
//
// The split of a column transform is   A: for every n2, N1-point FFTs over n1 (rows n1*N2 + n2), times W_N^(k1*n2)
//                                      B: for every k1, N2-point FFTs over n2 (rows k1*N2 + n2) -> rows k1 + N1*k2.
// As two launches the intermediate (the whole array) goes to HBM and comes back: 4 passes of traffic for one axis.
// Here a "tile" is one group of LPC adjacent columns x all N rows (N*LPC elements: 2 MB for 8192 x 16 complex128).
// Work items are dealt out IN ORDER from one queue, round r = the A items of tile r followed by the B items of tile
// r - LAG:      A(0) | A(1) | ... | A(LAG) B(0) | A(LAG+1) B(1) | ... | B(T-1)
// LAG is chosen so that LAG rounds hold more items than the grid has in flight (about two per CTA): by the time a B item
// is claimed, the A items of its tile were claimed a whole in-flight window earlier and have completed (a completion
// counter per tile makes that a guarantee instead of an expectation; a first version with LAG = 1 spent 75 % of its stall
// samples spinning on it).  A writes its tile into a RING of 2*LAG slots that B reads back microseconds later: the
// intermediate lives in the 126 MB L2 and is overwritten there before it is ever evicted — HBM sees one read and one
// write of the array (ncu: 1.07 GB + 1.07 GB for the 8192 x 8192 complex128 column transform).
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
  uint32_t lag;             // B(c) is queued in round c + lag  (lag <= tiles, ring_slots >= min(tiles, 2*lag))
};

// Dependency counters are polled with RELAXED gpu-scope loads and the ring is read with gpu-scope (L2) loads too:
// an acquire load would invalidate the SM's whole L1 (CCTL.IVALL) on every poll — measured: 75 % of all stall samples,
// and the twiddle tables evicted with it.  Nothing read after the poll may come from L1, and nothing does: the ring
// loads below are strong loads served by L2, the point of coherence, and they are issued only after the poll's
// value has returned (the spin loop's exit depends on it), i.e. after the producer's fence + count.
__device__ __forceinline__ unsigned int ld_relaxed_u32(const unsigned int *p) {
  unsigned int v;
#if defined(__CUDA_ARCH__)
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
#else
  v = __atomic_load_n(p, __ATOMIC_ACQUIRE);
#endif
  return v;
}
__device__ __forceinline__ cx<double> ld_l2(const cx<double> *p) {
#if defined(__CUDA_ARCH__)
  cx<double> v;
  asm volatile("ld.relaxed.gpu.global.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p) : "memory");
  return v;
#else
  return *p;
#endif
}
__device__ __forceinline__ cx<float> ld_l2(const cx<float> *p) {
#if defined(__CUDA_ARCH__)
  cx<float> v;
  asm volatile("ld.relaxed.gpu.global.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p) : "memory");
  return v;
#else
  return *p;
#endif
}
// 16-byte asynchronous copy global -> shared through L2 only (cp.async.cg): the NEXT item's elements are requested into
// thread-private staging slots while the current item is transformed; a thread waits only for its own copies.
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc) {
#if defined(__CUDA_ARCH__)
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
#else
  *reinterpret_cast<cx<double> *>(smem_dst) = *reinterpret_cast<const cx<double> *>(gsrc);
#endif
}
__device__ __forceinline__ void cp_async_commit() {
#if defined(__CUDA_ARCH__)
  asm volatile("cp.async.commit_group;" ::: "memory");
#endif
}
__device__ __forceinline__ void cp_async_wait_but_one() {   // every group but the most recently committed one has landed
#if defined(__CUDA_ARCH__)
  asm volatile("cp.async.wait_group 1;" ::: "memory");
#endif
}
__device__ __forceinline__ void cp_async_wait_all() {
#if defined(__CUDA_ARCH__)
  asm volatile("cp.async.wait_group 0;" ::: "memory");
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
//   staged != nullptr: this thread's R1 input elements already sit in its staging slots staged[j * (R2*LPC)]
//   mid(): executed by every thread after its exchange writes, before the barrier (the kernel stages the next item there)
//   The exchange slot of element k of thread (i, line) is (k*R2 + i)*LPC + line — exactly the thread's staging slot k —
//   so when the input was staged, S IS the staging buffer: the exchange happens in place, without another barrier.
template <typename T, int R1, int R2, int LPC, bool BWD, bool IN_L2, bool OUT_L2, typename MID>
__device__ __forceinline__ void colfuse_item(const cx<T> *in, int64_t es_in, cx<T> *out, int64_t es_out, bool valid, const cx<T> *tw,
                                             const LineJob &J, bool tw4, uint32_t twi, T f, cx<T> *S, int line, int i,
                                             const cx<T> *staged, MID &&mid) {
  constexpr int NB2 = R1 / R2;
  cx<T> x[R1];
#pragma unroll
  for (int j = 0; j < R1; ++j) {
    if (staged) {
      x[j] = valid ? staged[j * (R2 * LPC)] : mk<T>((T)0, (T)0);
    } else {
      const cx<T> *p = in + (int64_t)(i + R2 * j) * es_in;
      x[j] = valid ? (IN_L2 ? ld_l2(p) : *p) : mk<T>((T)0, (T)0);
    }
    if (BWD) x[j].y = -x[j].y;
  }
  RegFFT<T, R1>::run(x);
#pragma unroll
  for (int k = 1; k < R1; ++k) x[k] = cmul(x[k], __ldg(tw + i * k));
#pragma unroll
  for (int k = 0; k < R1; ++k) S[(k * R2 + i) * LPC + line] = x[k];
  mid();
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

// one claimed work item, decoded once by the control thread and read by everyone from shared memory
struct FuseItem { uint32_t valid, phaseA, c, q, g0, i2, slot, pad; };

__device__ __forceinline__ void colfuse_decode(const FuseJob &F, uint32_t item, FuseItem *it) {
  const uint32_t T = F.tiles, D = F.lag, a = F.itemsA, b = F.itemsB;
  it->valid = item < T * (a + b) ? 1u : 0u;
  if (!it->valid) return;
  // rounds 0..D-1: A only; D..T-1: A(r) then B(r-D); T..T+D-1: B(r-D) only
  const uint32_t head = D * a, mid = (T - D) * (a + b);
  uint32_t phaseA, c, q;
  if (item < head) { phaseA = 1; c = item / a; q = item - c * a; }
  else if (item < head + mid) {
    const uint32_t s = item - head, r = s / (a + b);
    q = s - r * (a + b);
    if (q < a) { phaseA = 1; c = D + r; } else { phaseA = 0; c = r; q -= a; }
  } else {
    const uint32_t s = item - head - mid, r = s / b;
    phaseA = 0; c = (T - D) + r; q = s - r * b;
  }
  it->phaseA = phaseA; it->c = c; it->q = q;
  it->i2 = c / F.g0n; it->g0 = c - it->i2 * F.g0n; it->slot = c % F.ring_slots;
}

// Synchronisation per item: ONE barrier inside the two-pass transform and ONE at its end.  The control thread claims
// items two ahead with an atomic issued at the top of the current one (its latency hides behind the work) and decodes
// them after the arithmetic; completion is published by the control thread alone — barrier, then fence + atomic, the
// grid-synchronisation pattern — instead of a fence per thread.
// Dependencies: every thread polls the NEXT item's counter itself with a relaxed L2 load issued at the top of the
// current item and consumed only after the first pass, so the poll's round trip (it was 17 % of the stall samples when
// consumed at once) overlaps arithmetic, and no barrier follows it.
// complex128: if the poll says the next item may run, the thread requests ITS input elements of that item with
// cp.async.cg into private staging slots (two staging buffers; the exchange then happens in place in those slots),
// so the HBM / L2 latency of the loads — 6.7 stall cycles per issued instruction with direct loads — hides behind
// the second pass, the stores and the barrier.  An item whose dependency was not met at that point is not staged:
// when its turn comes it waits and loads directly (never block on a dependency before the current item is
// published: it may be that very item).
template <typename T, int RA1, int RA2, int RB1, int RB2, int LPC, bool BWD>
__global__ void __launch_bounds__(LPC * RA2)
colfuse2_kernel(const __grid_constant__ FuseJob F) {
  static_assert(RA2 == RB2, "both phases run on the same threads");
  constexpr int NA = RA1 * RA2, NB = RB1 * RB2, NMAX = NA > NB ? NA : NB;
  constexpr bool STAGE = sizeof(cx<T>) == 16;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cx<T> *bufs = reinterpret_cast<cx<T> *>(smem_raw);                         // [STAGE ? 2 : 1][NMAX * LPC]
  FuseItem *s_it = reinterpret_cast<FuseItem *>(bufs + (STAGE ? 2 : 1) * (size_t)NMAX * LPC);   // [3]
  const int u = threadIdx.x, line = u % LPC, i = u / LPC;
  unsigned int *doneA = F.ctrl + 2, *doneB = F.ctrl + 2 + F.tiles;
  const uint64_t slot_elems = (uint64_t)NA * NB * LPC;
  // B needs every A item of its tile; A needs the tile that used its ring slot before to be consumed
  auto dep_counter = [&](const FuseItem &w) -> const unsigned int * {
    if (!w.valid) return nullptr;
    if (w.phaseA) return w.c >= F.ring_slots ? &doneB[w.c - F.ring_slots] : nullptr;
    return &doneA[w.c];
  };
  auto dep_need = [&](const FuseItem &w) -> unsigned int { return w.phaseA ? F.itemsB : F.itemsA; };
  // request this thread's input elements of item w into its slots of staging buffer b (slot j at (i + R2*j)*LPC + line)
  auto stage_item = [&](const FuseItem &w, int b) {
    const uint32_t l0 = w.g0 * LPC + line;
    cx<T> *dst = bufs + (size_t)b * NMAX * LPC + (size_t)i * LPC + line;
    if (w.phaseA) {
      const LineJob &J = F.A;
      if (l0 < (uint32_t)J.bdim[0]) {
        const cx<T> *in = reinterpret_cast<const cx<T> *>(J.in) + (int64_t)l0 + (int64_t)w.q * J.bs_in[1] + (int64_t)w.i2 * J.bs_in[2];
#pragma unroll
        for (int j = 0; j < RA1; ++j) cp_async16(dst + (size_t)j * (RA2 * LPC), in + (int64_t)(i + RA2 * j) * J.es_in);
      }
    } else if (l0 < (uint32_t)F.B.bdim[0]) {
      const cx<T> *in = reinterpret_cast<const cx<T> *>(F.ring) + (uint64_t)w.slot * slot_elems + line + (uint64_t)w.q * NB * LPC;
#pragma unroll
      for (int j = 0; j < RB1; ++j) cp_async16(dst + (size_t)j * (RB2 * LPC), in + (int64_t)(i + RB2 * j) * LPC);
    }
  };
  if (u == 0) {
    colfuse_decode(F, atomicAdd(&F.ctrl[0], 1u), &s_it[0]);
    colfuse_decode(F, atomicAdd(&F.ctrl[0], 1u), &s_it[1]);
  }
  __syncthreads();
  bool ready = false;           // the current item's dependency was seen satisfied (and, STAGE, its input is staged)
  for (uint32_t k = 0;; ++k) {
    const FuseItem it = s_it[k % 3];
    if (!it.valid) break;
    const FuseItem nx = s_it[(k + 1) % 3];
    uint32_t next_item = 0;
    if (u == 0) next_item = atomicAdd(&F.ctrl[0], 1u);      // item k+2, decoded after this item's arithmetic
    // the next item's dependency: requested now, looked at after the first pass
    const unsigned int *nx_ctr = dep_counter(nx);
    unsigned int nx_seen = 0;
    if (nx_ctr) nx_seen = ld_relaxed_u32(nx_ctr);
    if (!ready) {                                           // (first item of the CTA, or a dependency that was late)
      const unsigned int *c = dep_counter(it);
      if (c) while (ld_relaxed_u32(c) < dep_need(it)) {}
    }
    const bool staged_now = STAGE && ready;
    if (staged_now) cp_async_wait_all();                    // this thread's copies of item k have landed in its slots
    cx<T> *S = bufs + (size_t)(STAGE ? (k & 1) : 0) * NMAX * LPC;
    const cx<T> *staged = staged_now ? S + (size_t)i * LPC + line : nullptr;
    cx<T> *ring = reinterpret_cast<cx<T> *>(F.ring) + (uint64_t)it.slot * slot_elems + line;
    const uint32_t l0 = it.g0 * LPC + line, q = it.q;
    bool ready_next = false;
    auto mid = [&]() {
      ready_next = nx.valid && (!nx_ctr || nx_seen >= dep_need(nx));
      if (STAGE && ready_next) { stage_item(nx, (k + 1) & 1); cp_async_commit(); }
    };
    if (it.phaseA) {
      // group (g0, n2 = q, i2): N1-point FFT over n1 of src rows n1*N2 + n2; output k1 -> ring row k1*N2 + n2
      const LineJob &J = F.A;
      const bool valid = l0 < (uint32_t)J.bdim[0];
      const cx<T> *in = reinterpret_cast<const cx<T> *>(J.in) + (int64_t)l0 + (int64_t)q * J.bs_in[1] + (int64_t)it.i2 * J.bs_in[2];
      const uint32_t twi = J.tw4_dim == 0 ? l0 : J.tw4_dim == 1 ? q : J.tw4_dim == 2 ? it.i2 : 0u;
      colfuse_item<T, RA1, RA2, LPC, BWD, false, true>(in, J.es_in, ring + (uint64_t)q * LPC, (int64_t)NB * LPC, valid,
                                                       reinterpret_cast<const cx<T> *>(J.tw), J, J.tw4_n != 0, twi, (T)J.fct, S, line, i, staged, mid);
    } else {
      // group (g0, k1 = q, i2): N2-point FFT over n2 of ring rows k1*N2 + n2; output k2 -> dst element k1 + N1*k2
      const LineJob &J = F.B;
      const bool valid = l0 < (uint32_t)J.bdim[0];
      cx<T> *out = reinterpret_cast<cx<T> *>(J.out) + (int64_t)l0 + (int64_t)q * J.bs_out[1] + (int64_t)it.i2 * J.bs_out[2];
      colfuse_item<T, RB1, RB2, LPC, BWD, true, false>(ring + (uint64_t)q * NB * LPC, (int64_t)LPC, out, J.es_out, valid,
                                                       reinterpret_cast<const cx<T> *>(J.tw), J, false, 0u, (T)J.fct, S, line, i, staged, mid);
    }
    if (u == 0) colfuse_decode(F, next_item, &s_it[(k + 2) % 3]);
    __syncthreads();            // every thread's stores are issued; this item's buffer and s_it[k % 3] are free again
    if (u == 0) {               // publish: bar.sync + fence by one thread orders the whole CTA's stores before the count
      __threadfence();
      atomicAdd(it.phaseA ? &doneA[it.c] : &doneB[it.c], 1u);
    }
    ready = ready_next;
  }
  if (STAGE) cp_async_wait_all();   // nothing may still be landing in shared memory when the CTA retires
}

}  // namespace impulse
