// Launcher of the fused column transform (colfuse2_kernel, colfuse_device.cuh): both launches of a four-step split
// in one persistent kernel with the intermediate in an L2-resident ring.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>

#include "colfuse_device.cuh"
#include "fast_common.h"
#include "fft_kernels.h"

namespace impulse {

namespace {
template <typename T, int RA1, int RA2, int RB1, int RB2, int LPC>
int launch_colfuse(const FuseJob &F, int sm_count, cudaStream_t s) {
  constexpr int NA = RA1 * RA2, NB = RB1 * RB2, NMAX = NA > NB ? NA : NB;
  const size_t smem = sizeof(cx<T>) * (size_t)NMAX * LPC * (sizeof(cx<T>) == 16 ? 2 : 1) + 3 * sizeof(FuseItem);
  const bool bwd = (F.A.flags & F_CONJ_SEQ) != 0;
  auto kf = colfuse2_kernel<T, RA1, RA2, RB1, RB2, LPC, false>;
  auto kb = colfuse2_kernel<T, RA1, RA2, RB1, RB2, LPC, true>;
  static PerDeviceFlag flag;
  static int ctas_per_sm[kMaxDevices] = {};
  bool &configured = flag.here();
  if (!configured) {
    for (auto k : {kf, kb}) {
      cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return (int)e;
      e = cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
      if (e != cudaSuccess) return (int)e;
    }
    int nb = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kf, LPC * RA2, smem);
    if (e != cudaSuccess) return (int)e;
    ctas_per_sm[cur_dev()] = nb > 0 ? nb : 1;
    configured = true;
  }
  const uint64_t total = (uint64_t)F.tiles * (F.itemsA + F.itemsB);
  if (total == 0) return 0;
  if (total > 0xfff00000ull) return (int)cudaErrorInvalidValue;
  // every CTA of the grid must be resident (a waiting CTA relies on the CTAs holding earlier items to be running)
  uint64_t grid = (uint64_t)sm_count * ctas_per_sm[cur_dev()];
  if (grid > total) grid = total;
  static thread_local char name[96];
  snprintf(name, sizeof(name), "colfuse2_kernel<%s,%dx%d,%d>", sizeof(T) == 8 ? "double" : "float", NA, NB, LPC);
  g_last_kernel = name;
  (bwd ? kb : kf)<<<(unsigned)grid, LPC * RA2, smem, s>>>(F);
  return (int)cudaGetLastError();
}
}  // namespace

// Queue lag and ring size (tiles): the B items of a tile are queued `lag` rounds after its A items, `lag` rounds holding
// more items than the grid keeps in flight (three per CTA: in progress, staged, claimed); the ring holds
// 2*lag tiles so that a slot is rewritten only after its previous tile was consumed a window ago.  IMPULSE_FFT_FUSE_LAG
// overrides (A/B runs).
static uint32_t fuse_lag(uint32_t tiles, uint32_t per_round, int sm_count) {
  static const int env = [] { const char *e = getenv("IMPULSE_FFT_FUSE_LAG"); return e ? atoi(e) : 0; }();
  uint32_t lag = env > 0 ? (uint32_t)env : (uint32_t)((3ull * (uint64_t)sm_count * 3 + per_round - 1) / per_round) + 1;
  if (lag > tiles) lag = tiles;
  return lag ? lag : 1;
}
static uint32_t fuse_slots(uint32_t tiles, uint32_t lag) { return tiles < 2 * lag ? tiles : 2 * lag; }

// scratch the fused launch needs: the ring + the control words (which must be zero at launch)
size_t colfuse_scratch_bytes(const LineJob &a, const LineJob &b, uint32_t tiles, size_t *ctrl_off, size_t *ctrl_bytes) {
  const size_t csz = a.dtype == 1 ? 16 : 8;
  const uint32_t slots = fuse_slots(tiles ? tiles : 1, fuse_lag(tiles ? tiles : 1, a.n_fft + b.n_fft, 148));
  size_t ring = (size_t)slots * a.n_fft * b.n_fft * 16 * csz;
  ring = (ring + 255) & ~(size_t)255;
  const size_t cb = sizeof(unsigned int) * (2 + 2 * (size_t)tiles);
  if (ctrl_off) *ctrl_off = ring;
  if (ctrl_bytes) *ctrl_bytes = cb;
  return ring + cb;
}

int launch_colfuse_pair(const LineJob &a, const LineJob &b, uint32_t tiles, uint32_t g0n, void *scratch, int sm_count, void *stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  FuseJob F;
  F.A = a; F.B = b;
  size_t ctrl_off = 0;
  colfuse_scratch_bytes(a, b, tiles, &ctrl_off, nullptr);
  F.ring = scratch;
  F.ctrl = reinterpret_cast<unsigned int *>(static_cast<unsigned char *>(scratch) + ctrl_off);
  F.lag = fuse_lag(tiles ? tiles : 1, a.n_fft + b.n_fft, 148);   // (sized for a full B200: a smaller grid only waits less)
  F.ring_slots = fuse_slots(tiles ? tiles : 1, F.lag);
  F.tiles = tiles; F.g0n = g0n;
  F.itemsA = b.n_fft;   // one A item per n2
  F.itemsB = a.n_fft;   // one B item per k1
  const uint32_t ia = a.fast_id, ib = b.fast_id;
  if (ia == COL2_64_F64 && ib == COL2_64_F64) return launch_colfuse<double, 8, 8, 8, 8, 16>(F, sm_count, s);
  if (ia == COL2_64_F64 && ib == COL2_128_F64) return launch_colfuse<double, 8, 8, 16, 8, 16>(F, sm_count, s);
  if (ia == COL2_128_F64 && ib == COL2_128_F64) return launch_colfuse<double, 16, 8, 16, 8, 16>(F, sm_count, s);
  if (ia == COL2_64_F32 && ib == COL2_64_F32) return launch_colfuse<float, 8, 8, 8, 8, 16>(F, sm_count, s);
  if (ia == COL2_64_F32 && ib == COL2_128_F32) return launch_colfuse<float, 8, 8, 16, 8, 16>(F, sm_count, s);
  if (ia == COL2_128_F32 && ib == COL2_128_F32) return launch_colfuse<float, 16, 8, 16, 8, 16>(F, sm_count, s);
  return -1;
}

bool colfuse_pair_supported(uint32_t a, uint32_t b) {
  return (a == COL2_64_F64 && (b == COL2_64_F64 || b == COL2_128_F64)) || (a == COL2_128_F64 && b == COL2_128_F64) ||
         (a == COL2_64_F32 && (b == COL2_64_F32 || b == COL2_128_F32)) || (a == COL2_128_F32 && b == COL2_128_F32);
}

}  // namespace impulse
