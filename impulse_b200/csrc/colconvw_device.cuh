// Whole-axis convolution kernel (colconvw_kernel): IFFT_axis(FFT_axis(x) .* m) along a STRIDED axis in ONE pass over
// the data.  A CTA keeps W adjacent lines of the full axis (N = R1*R2*R3 points each) in shared memory — 128 KB for
// 4 float32 lines of 4096 points — and runs both transforms and the multiply there, so every element is read from HBM
// once and written once.  The three-launch scheme (pass A, colconv2_kernel, pass B) moves the array three times.
//
// In-place three-pass decomposition (n = i1 + M1*j1, i1 = i2 + R3*j2, k = k1 + R1*k2 + R1*R2*k3, M1 = R2*R3):
//   F1  x[i1 + M1*j1] --radix R1 over j1, * W_N^(i1 k1)-->            A[k1][i1]      at  k1*M1 + i1
//   F2  A[k1][i2 + R3*j2] --radix R2 over j2, * W_M1^(i2 k2)-->       B[k1][k2][i2]  at  k1*M1 + i2 + R3*k2
//   F3  B[k1][k2][.] --radix R3 over i2-->  X[k], k3 = 0..R3-1 in the registers of ONE thread
//       * m[k], conjugate, and straight on with the inverse transform's first pass (no exchange):
//   I3  c[k1][k2][.] --radix R3 over k3, * W_M1^(i2 k2)-->            C[k1][k2][i2]  at  k1*M1 + R3*k2 + i2
//   I2  C[k1][.][i2] --radix R2 over k2, * W_N^(i1 k1)-->             D[k1][i1]      at  k1*M1 + i2 + R3*j2
//   I1  D[.][i1] --radix R1 over k1--> conj --> y[i1 + M1*j1]
// (the inverse is the transposed flow on the conjugated products: z[n] = sum_k conj(X m)[k] W_N^(nk), y = conj z).
// Every pass reads and writes the SAME positions per butterfly, so the tile needs no second buffer and one barrier per
// pass; I1 of one tile and F1 of the CTA's next tile touch the same thread-private positions, so no barrier separates
// the tiles.
// Shared layout [pos][line] with pos ^= (pos / R3) & (G - 1), G = butterflies per 128-byte wavefront: conflict-free in
// all six passes (pass 3 walks consecutive positions per thread; the others walk consecutive positions per lane).
// A thread owns LP adjacent lines of its butterfly (one 16-byte shared / global access per point for two float32
// lines; the twiddles are fetched once for both).  Global accesses are W*sizeof(complex) runs (32 B sectors for 4
// float32 lines), one per axis element; the CTA's NEXT tile is requested into L2 while the current one is transformed
// (ncu on the first version: 29 % issue-active, long-scoreboard 5.8 per issue — the single resident CTA exposed every
// DRAM round trip), and W_N^(i1 k1) lives in shared memory beside the tile (the streamed tile evicted it from L1).
#pragma once
#include "fast3_device.cuh"
#include "tmem_device.cuh"

namespace impulse {

template <typename T, int LP> struct alignas((LP * sizeof(cx<T>)) >= 16 ? 16 : 8) CwVec { cx<T> v[LP]; };

// MODE: CW_CONV = the convolution above; CW_FWD / CW_BWD = the plain transform of the axis (passes F1, F2, F3 and a store
// straight from the last butterfly's registers; backward = conj(FFT(conj x))): ONE pass over the data for strided
// power-of-two axes that the four-step split serves in two.
enum { CW_CONV = 0, CW_FWD = 1, CW_BWD = 2 };
// GV: the LP lines of a thread are read / written as one vector in global memory (the launcher checks alignment)
// TM: the CTA's NEXT tile is loaded from HBM while the current one is transformed — half of a thread's points during
// pass F2, half during I2 (F3 in the plain modes) — and parked in TENSOR MEMORY until pass F1 picks it up: the single
// resident CTA no longer waits for its loads at every tile boundary (the tile itself fills the shared memory).
template <typename T, int R1, int R2, int R3, int W, int LP, int TT, bool GV, int MODE = CW_CONV, bool TM = false>
__global__ void __launch_bounds__(TT, 1)
colconvw_kernel(const __grid_constant__ LineJob J) {
  constexpr int N = R1 * R2 * R3, M1 = R2 * R3, PW = W / LP, TB = TT / PW;
  constexpr int NB1 = (N / R1) / TB, NB2 = (N / R2) / TB, NB3 = (N / R3) / TB;
  constexpr int GW = 128 / (W * (int)sizeof(cx<T>)), G = GW < 1 ? 1 : GW;
  static_assert(W % LP == 0 && TT % PW == 0 && (N / R1) % TB == 0 && (N / R2) % TB == 0 && (N / R3) % TB == 0, "whole-axis shape");
  static_assert((R3 & (R3 - 1)) == 0 && (G & (G - 1)) == 0 && G <= R3 && TB % G == 0, "swizzle needs powers of two");
  using Vec = CwVec<T, LP>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cx<T> *S = reinterpret_cast<cx<T> *>(smem_raw);
  cx<T> *s_tw1 = S + (size_t)N * W;   // [R1][M1]: W_N^(i1 k1)
  cx<T> *s_tw2 = s_tw1 + N;           // [R2][R3]: W_M1^(i2 k2)
  const int t = threadIdx.x, lp = t % PW, tb = t / PW;
  // tensor-memory staging: WV words per Vec, R1 * WV words per thread, (TT / 128) warps per lane quarter side by side
  constexpr int WV = (int)(sizeof(Vec) / 4), TM_COLS = R1 * WV * (TT / 128), H1 = R1 / 2;
  static_assert(!TM || (NB1 == 1 && TT % 128 == 0 && WV % 4 == 0 && (H1 * WV) % 8 == 0 && TM_COLS <= 512 && TM_COLS >= 32 &&
                        (TM_COLS & (TM_COLS - 1)) == 0), "tensor-memory staging shape");
  uint32_t tm_addr = 0, tm_base = 0;
  (void)tm_base;
  if constexpr (TM) {
    tm_base = cw_tmem_acquire<TM_COLS>(reinterpret_cast<uint32_t *>(s_tw2 + R2 * R3), t);
    tm_addr = tm_base + ((uint32_t)(((t >> 5) & 3) * 32) << 16) + (uint32_t)(t >> 7) * (R1 * WV);
  }

  const cx<T> *um = reinterpret_cast<const cx<T> *>(J.umul);
  for (int idx = t; idx < N; idx += TT) s_tw1[idx] = reinterpret_cast<const cx<T> *>(J.f3_tw1)[idx];
  for (int idx = t; idx < R2 * R3; idx += TT) s_tw2[idx] = reinterpret_cast<const cx<T> *>(J.f3_tw2)[idx];
  auto at = [&](const uint32_t pos) -> Vec & {
    return *reinterpret_cast<Vec *>(S + (pos ^ ((pos / R3) & (G - 1))) * W + lp * LP);
  };

  // tile -> (group of W adjacent lines, outer indices); the dimension along which the multiplier repeats (bdim[1])
  // runs fastest, so that one block of multipliers serves every image in turn from L2
  // J.n_load = GF: GF adjacent groups run fastest of all, i.e. on CTAs that execute side by side, so that the 32-byte
  // pieces of one 128-byte line are requested together (DRAM row hits instead of one activate per piece)
  const uint32_t g0n = (uint32_t)((J.bdim[0] + W - 1) / W), d1 = (uint32_t)J.bdim[1];
  const uint32_t gf = J.n_load ? J.n_load : 1u, ghn = (g0n + gf - 1) / gf;
  const uint32_t ntiles = ghn * gf * d1 * (uint32_t)J.bdim[2];
  struct Tile { int64_t in0, out0; uint64_t umoff; int nv; };   // offsets of the group's first line; valid lines of this thread
  auto tile_of = [&](const uint32_t id) {
    const uint32_t ra = id / gf, glo = id - ra * gf, r = ra / d1, i1 = ra - r * d1, i2 = r / ghn, g0 = (r - i2 * ghn) * gf + glo;
    Tile q;
    q.in0 = (int64_t)(g0 * W) + (int64_t)i1 * J.bs_in[1] + (int64_t)i2 * J.bs_in[2];
    q.out0 = (int64_t)(g0 * W) + (int64_t)i1 * J.bs_out[1] + (int64_t)i2 * J.bs_out[2];
    q.umoff = MODE == CW_CONV ? (uint64_t)(q.out0 + lp * LP) % J.umul_mod : 0;
    const int left = (int)J.bdim[0] - (int)(g0 * W) - lp * LP;   // (groups past the end of a ragged last block: nothing valid)
    q.nv = left >= LP ? LP : (left > 0 ? left : 0);
    return q;
  };
  // R1 points (stride `step` elements) of the thread's lines <-> global memory; the vector / ragged choice is made
  // once per butterfly, not per point
  auto gget = [&](const cx<T> *p, const int nv, const int64_t step, Vec (&x)[R1]) {
    if (GV && nv == LP) {
#pragma unroll
      for (int j = 0; j < R1; ++j) { x[j] = *reinterpret_cast<const Vec *>(p); p += step; }
    } else {
#pragma unroll
      for (int j = 0; j < R1; ++j) {
#pragma unroll
        for (int l = 0; l < LP; ++l) x[j].v[l] = l < nv ? p[l] : mk<T>((T)0, (T)0);
        p += step;
      }
    }
    if (MODE == CW_BWD) {
#pragma unroll
      for (int j = 0; j < R1; ++j)
#pragma unroll
        for (int l = 0; l < LP; ++l) x[j].v[l].y = -x[j].v[l].y;
    }
  };
  auto gput = [&](cx<T> *p, const int nv, const int64_t step, const auto &x) {   // x: Vec[R1] or Vec[R3]
    constexpr int R = (int)(sizeof(x) / sizeof(Vec));
    if (GV && nv == LP) {
#pragma unroll
      for (int j = 0; j < R; ++j) { *reinterpret_cast<Vec *>(p) = x[j]; p += step; }
    } else {
#pragma unroll
      for (int j = 0; j < R; ++j) {
#pragma unroll
        for (int l = 0; l < LP; ++l) if (l < nv) p[l] = x[j].v[l];
        p += step;
      }
    }
  };
  // R-point transforms of the LP lines held as x[point].v[line], in place
  auto fftR1 = [&](Vec (&x)[R1]) {
#pragma unroll
    for (int l = 0; l < LP; ++l) {
      cx<T> y[R1];
#pragma unroll
      for (int j = 0; j < R1; ++j) y[j] = x[j].v[l];
      RegFFT<T, R1>::run(y);
#pragma unroll
      for (int j = 0; j < R1; ++j) x[j].v[l] = y[j];
    }
  };
  auto fftR2 = [&](Vec (&x)[R2]) {
#pragma unroll
    for (int l = 0; l < LP; ++l) {
      cx<T> y[R2];
#pragma unroll
      for (int j = 0; j < R2; ++j) y[j] = x[j].v[l];
      RegFFT<T, R2>::run(y);
#pragma unroll
      for (int j = 0; j < R2; ++j) x[j].v[l] = y[j];
    }
  };
  auto fftR3 = [&](Vec (&x)[R3]) {
#pragma unroll
    for (int l = 0; l < LP; ++l) {
      cx<T> y[R3];
#pragma unroll
      for (int j = 0; j < R3; ++j) y[j] = x[j].v[l];
      RegFFT<T, R3>::run(y);
#pragma unroll
      for (int j = 0; j < R3; ++j) x[j].v[l] = y[j];
    }
  };
  // rows [j0, j0 + H1) of the thread's butterfly of tile q: HBM -> registers, and registers -> tensor memory
  auto tm_fetch = [&](const Tile &q, const int j0, Vec (&pre)[H1]) {
    const cx<T> *p = reinterpret_cast<const cx<T> *>(J.in) + q.in0 + lp * LP + (int64_t)(tb + M1 * j0) * J.es_in;
    const int64_t step = (int64_t)M1 * J.es_in;
    if (GV && q.nv == LP) {
#pragma unroll
      for (int j = 0; j < H1; ++j) { pre[j] = *reinterpret_cast<const Vec *>(p); p += step; }
    } else {
#pragma unroll
      for (int j = 0; j < H1; ++j) {
#pragma unroll
        for (int l = 0; l < LP; ++l) pre[j].v[l] = l < q.nv ? p[l] : mk<T>((T)0, (T)0);
        p += step;
      }
    }
  };
  auto tm_park = [&](const int j0, const Vec (&pre)[H1]) {
    const uint32_t *w = reinterpret_cast<const uint32_t *>(&pre[0]);
#pragma unroll
    for (int c = 0; c < H1 * WV; c += 8) cw_tmem_st8(tm_addr + (uint32_t)(j0 * WV + c), w + c);
  };
  auto f1 = [&](const Tile &q, const int i1, const bool parked = false) {
    Vec x[R1];
    if (TM && parked) {   // the tile was loaded ahead and waits in tensor memory
      cw_tmem_wait_st();
      uint32_t *w = reinterpret_cast<uint32_t *>(&x[0]);
#pragma unroll
      for (int c = 0; c < R1 * WV; c += 8) cw_tmem_ld8(tm_addr + (uint32_t)c, w + c);
      cw_tmem_wait_ld();
      if (MODE == CW_BWD) {
#pragma unroll
        for (int j = 0; j < R1; ++j)
#pragma unroll
          for (int l = 0; l < LP; ++l) x[j].v[l].y = -x[j].v[l].y;
      }
    } else {
      gget(reinterpret_cast<const cx<T> *>(J.in) + q.in0 + lp * LP + (int64_t)i1 * J.es_in, q.nv, (int64_t)M1 * J.es_in, x);
    }
    fftR1(x);
#pragma unroll
    for (int k = 1; k < R1; ++k) {
      const cx<T> w = s_tw1[k * M1 + i1];
#pragma unroll
      for (int l = 0; l < LP; ++l) x[k].v[l] = cmul(x[k].v[l], w);
    }
#pragma unroll
    for (int k = 0; k < R1; ++k) at((uint32_t)(k * M1 + i1)) = x[k];
  };
  // ask L2 for a tile (J.n_store: 0 = no, 1 = one prefetch.global.L2 per 32-byte piece of every W-line run, 2 = one bulk
  // prefetch per run: the TMA unit's queue instead of the load / store unit's; needs 16-byte aligned runs = GV)
  auto prefetch_tile = [&](const Tile &q) {
#if defined(__CUDA_ARCH__)
    const char *base = reinterpret_cast<const char *>(reinterpret_cast<const cx<T> *>(J.in) + q.in0);
    const int64_t step = J.es_in * (int64_t)sizeof(cx<T>);
    if (GV && J.n_store == 2) {
      for (int r = t; r < N; r += TT) prefetch_l2_bulk(base + (int64_t)r * step, (uint32_t)(W * sizeof(cx<T>)));
    } else if (J.n_store) {
      constexpr int PIECES = (W * (int)sizeof(cx<T>) + 31) / 32;
      for (int r = t; r < N * PIECES; r += TT)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(base + (int64_t)(r / PIECES) * step + (r % PIECES) * 32));
    }
#else
    (void)q;
#endif
  };

  auto tm_release = [&]() {   // every thread of the CTA comes here once, after its last tensor-memory access
    if constexpr (TM) cw_tmem_release<TM_COLS>(tm_base, t);
  };
  uint32_t tile = blockIdx.x;
  if (tile >= ntiles) { tm_release(); return; }
  Tile cur = tile_of(tile);
  __syncthreads();   // the tables
  if (MODE == CW_CONV) {
#pragma unroll 1
    for (int m = 0; m < NB1; ++m) f1(cur, tb + TB * m);
  }
  const T f = (T)J.fct;
  bool parked = false;   // plain modes: the current tile waits in tensor memory (every tile but the CTA's first)
  for (;;) {
    const uint32_t next = tile + gridDim.x;
    const bool more = next < ntiles;
    Tile nxt = cur;
    if (more) { nxt = tile_of(next); prefetch_tile(nxt); }
    if (MODE != CW_CONV) {
#pragma unroll 1
      for (int m = 0; m < NB1; ++m) f1(cur, tb + TB * m, parked);
    }
    __syncthreads();
    // ---------------- F2 (with TM: the first half of the next tile's points travels HBM -> registers -> tensor memory meanwhile)
    Vec pre[TM ? H1 : 1];
    if constexpr (TM) { if (more) tm_fetch(nxt, 0, pre); }
#pragma unroll 1
    for (int m = 0; m < NB2; ++m) {
      const int b = tb + TB * m, i2 = b % R3, k1 = b / R3;
      const uint32_t base = (uint32_t)(k1 * M1 + i2);
      Vec y[R2];
#pragma unroll
      for (int j = 0; j < R2; ++j) y[j] = at(base + R3 * j);
      fftR2(y);
#pragma unroll
      for (int k = 1; k < R2; ++k) {
        const cx<T> w = s_tw2[k * R3 + i2];
#pragma unroll
        for (int l = 0; l < LP; ++l) y[k].v[l] = cmul(y[k].v[l], w);
      }
#pragma unroll
      for (int k = 0; k < R2; ++k) at(base + R3 * k) = y[k];
    }
    if constexpr (TM) { if (more) tm_park(0, pre); }
    __syncthreads();
    if constexpr (MODE != CW_CONV) {
      if constexpr (TM) { if (more) tm_fetch(nxt, H1, pre); }
      // ---------------- F3 + store: X[k1 + R1 k2 + R1 R2 k3] leaves from the butterfly's registers
#pragma unroll 1
      for (int m = 0; m < NB3; ++m) {
        const int b = tb + TB * m, k2 = b % R2, k1 = b / R2;
        const uint32_t base = (uint32_t)(k1 * M1 + R3 * k2);
        Vec y[R3];
#pragma unroll
        for (int j = 0; j < R3; ++j) y[j] = at(base + j);
        fftR3(y);
        const T fy = MODE == CW_BWD ? -f : f;
#pragma unroll
        for (int k3 = 0; k3 < R3; ++k3)
#pragma unroll
          for (int l = 0; l < LP; ++l) y[k3].v[l] = mk<T>(y[k3].v[l].x * f, y[k3].v[l].y * fy);
        gput(reinterpret_cast<cx<T> *>(J.out) + cur.out0 + lp * LP + (int64_t)(k1 + R1 * k2) * J.es_out, cur.nv,
             (int64_t)(R1 * R2) * J.es_out, y);
      }
      if (!more) break;
      if constexpr (TM) { tm_park(H1, pre); parked = true; }
      __syncthreads();   // every pass-3 read before the next tile's pass-1 writes
      cur = nxt;
      tile = next;
      continue;
    }
    // ---------------- F3, multiply, I3
#pragma unroll 1
    for (int m = 0; m < NB3; ++m) {
      const int b = tb + TB * m, k2 = b % R2, k1 = b / R2;
      const uint32_t base = (uint32_t)(k1 * M1 + R3 * k2);
      Vec mreg[R3];   // requested first: their L2 latency overlaps the forward butterfly
      {
        const uint64_t o0 = cur.umoff + (uint64_t)((int64_t)(k1 + R1 * k2) * J.es_out), ostep = (uint64_t)((int64_t)(R1 * R2) * J.es_out);
        if (GV && cur.nv == LP && o0 + (R3 - 1) * ostep + LP <= J.umul_mod) {   // no wrap-around within this butterfly
          const cx<T> *p = um + o0;
#pragma unroll
          for (int k3 = 0; k3 < R3; ++k3) { mreg[k3] = *reinterpret_cast<const Vec *>(p); p += ostep; }
        } else {
          uint64_t o = o0;
#pragma unroll
          for (int k3 = 0; k3 < R3; ++k3) {
#pragma unroll
            for (int l = 0; l < LP; ++l) {
              uint64_t ol = o + l;
              if (ol >= J.umul_mod) ol %= J.umul_mod;
              mreg[k3].v[l] = l < cur.nv ? __ldg(um + ol) : mk<T>((T)0, (T)0);
            }
            o += ostep;
          }
        }
      }
      Vec y[R3];
#pragma unroll
      for (int j = 0; j < R3; ++j) y[j] = at(base + j);
      fftR3(y);
#pragma unroll
      for (int k3 = 0; k3 < R3; ++k3)
#pragma unroll
        for (int l = 0; l < LP; ++l) { y[k3].v[l] = cmul(y[k3].v[l], mreg[k3].v[l]); y[k3].v[l].y = -y[k3].v[l].y; }
      fftR3(y);
#pragma unroll
      for (int i2 = 1; i2 < R3; ++i2) {
        const cx<T> w = s_tw2[k2 * R3 + i2];
#pragma unroll
        for (int l = 0; l < LP; ++l) y[i2].v[l] = cmul(y[i2].v[l], w);
      }
#pragma unroll
      for (int i2 = 0; i2 < R3; ++i2) at(base + i2) = y[i2];
    }
    __syncthreads();
    // ---------------- I2 (with TM: the second half of the next tile's points)
    if constexpr (TM) { if (more) tm_fetch(nxt, H1, pre); }
#pragma unroll 1
    for (int m = 0; m < NB2; ++m) {
      const int b = tb + TB * m, i2 = b % R3, k1 = b / R3;
      const uint32_t base = (uint32_t)(k1 * M1 + i2);
      Vec y[R2];
#pragma unroll
      for (int k = 0; k < R2; ++k) y[k] = at(base + R3 * k);
      fftR2(y);
#pragma unroll
      for (int j = 0; j < R2; ++j) {
        const cx<T> w = s_tw1[base + R3 * j];
#pragma unroll
        for (int l = 0; l < LP; ++l) y[j].v[l] = cmul(y[j].v[l], w);
      }
#pragma unroll
      for (int j = 0; j < R2; ++j) at(base + R3 * j) = y[j];
    }
    if constexpr (TM) { if (more) tm_park(H1, pre); }
    __syncthreads();
    // ---------------- I1 + store, then F1 of this CTA's next tile on the same thread-private positions
#pragma unroll 1
    for (int m = 0; m < NB1; ++m) {
      const int i1 = tb + TB * m;
      {
        Vec y[R1];
#pragma unroll
        for (int k = 0; k < R1; ++k) y[k] = at((uint32_t)(k * M1 + i1));
        fftR1(y);
#pragma unroll
        for (int j = 0; j < R1; ++j)
#pragma unroll
          for (int l = 0; l < LP; ++l) y[j].v[l] = mk<T>(y[j].v[l].x * f, -y[j].v[l].y * f);
        gput(reinterpret_cast<cx<T> *>(J.out) + cur.out0 + lp * LP + (int64_t)i1 * J.es_out, cur.nv, (int64_t)M1 * J.es_out, y);
      }
      if (more) f1(nxt, i1, TM);
    }
    if (!more) break;
    cur = nxt;
    tile = next;
  }
  tm_release();
}

}  // namespace impulse
