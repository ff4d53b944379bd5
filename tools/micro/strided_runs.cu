// Micro-benchmark (tools, not product): what does HBM deliver for the whole-axis column tile pattern?
// Array [B][H][P] of 8-byte elements; a tile = all H rows x RUN bytes of one image; a CTA reads its tile (16 x 16-byte
// loads per thread in flight), writes it back in place, takes the next tile (image fastest).  RUN = 32 ... 512 bytes.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
template <int RUN, int TT>
__global__ void __launch_bounds__(TT) tile_copy(float4 *a, int B, int H, long pitch16 /* row pitch in 16-byte units */, int groups, int dummy) {
  extern __shared__ float4 sm[];
  constexpr int V = RUN / 16;              // 16-byte pieces per run
  constexpr int RPI = TT / V;              // rows per iteration
  const int t = threadIdx.x, v = t % V, r0 = t / V;
  const long ntiles = (long)groups * B;
  for (long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int img = (int)(tile % B), g = (int)(tile / B);
    float4 *base = a + ((long)img * H) * pitch16 + (long)g * V + v;
    for (int r = r0; r < H; r += RPI * 16) {
      float4 x[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) { const int rr = r + RPI * j; if (rr < H) x[j] = base[(long)rr * pitch16]; }
#pragma unroll
      for (int j = 0; j < 16; ++j) { const int rr = r + RPI * j; if (rr < H) { x[j].x += 1.f; base[(long)rr * pitch16] = x[j]; } }
    }
    if (dummy) sm[t] = base[0];
  }
}
// the same bytes as RUN = 64, but as two instructions of 32-byte runs (adjacent sectors requested back to back by one warp)
template <int TT>
__global__ void __launch_bounds__(TT) tile_copy_split(float4 *a, int B, int H, long pitch16, int groups) {
  constexpr int V = 2, RPI = TT / V;
  const int t = threadIdx.x, v = t % V, r0 = t / V;
  const long ntiles = (long)groups * B;
  for (long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int img = (int)(tile % B), g = (int)(tile / B);
    float4 *base = a + ((long)img * H) * pitch16 + (long)g * 4 + v;
    for (int r = r0; r < H; r += RPI * 8) {
      float4 x[8], y[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) { const int rr = r + RPI * j; if (rr < H) { x[j] = base[(long)rr * pitch16]; y[j] = base[(long)rr * pitch16 + 2]; } }
#pragma unroll
      for (int j = 0; j < 8; ++j) { const int rr = r + RPI * j; if (rr < H) { x[j].x += 1.f; base[(long)rr * pitch16] = x[j]; base[(long)rr * pitch16 + 2] = y[j]; } }
    }
  }
}
void run_split(float4 *a, int B, int H, long pitch16, int sms) {
  constexpr int TT = 512;
  const int groups = (int)(pitch16 * 16 / 64);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int it = 0; it < 2; ++it) tile_copy_split<TT><<<sms, TT>>>(a, B, H, pitch16, groups);
  cudaEventRecord(e0);
  for (int it = 0; it < 5; ++it) tile_copy_split<TT><<<sms, TT>>>(a, B, H, pitch16, groups);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
  const double bytes = 2.0 * B * H * (double)groups * 64;
  printf("run=2x32 B (two instructions, adjacent)  %.3f ms  %.0f GB/s  (%s)\n", ms, bytes / ms * 1e-6, cudaGetErrorString(cudaGetLastError()));
}
template <int RUN> void run(float4 *a, int B, int H, long pitch16, int smem_kb, int sms) {
  constexpr int TT = 512;
  const int groups = (int)(pitch16 * 16 / RUN);
  auto k = tile_copy<RUN, TT>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_kb * 1024);
  int nb = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k, TT, smem_kb * 1024);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int it = 0; it < 2; ++it) k<<<sms * nb, TT, smem_kb * 1024>>>(a, B, H, pitch16, groups, 0);
  cudaEventRecord(e0);
  for (int it = 0; it < 5; ++it) k<<<sms * nb, TT, smem_kb * 1024>>>(a, B, H, pitch16, groups, 0);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
  const double bytes = 2.0 * B * H * (double)groups * RUN;
  printf("run=%4d B  smem/CTA=%3d KB  CTAs/SM=%d  %.3f ms  %.0f GB/s  (%s)\n", RUN, smem_kb, nb, ms, bytes / ms * 1e-6, cudaGetErrorString(cudaGetLastError()));
}
int main() {
  const int B = 64, H = 4096; const long P = 2052, pitch16 = P * 8 / 16;
  float4 *a; cudaMalloc(&a, (size_t)B * H * P * 8); cudaMemset(a, 0, (size_t)B * H * P * 8);
  cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
  const int sms = pr.multiProcessorCount;
  run_split(a, B, H, pitch16, sms);
  for (int kb : {100}) {
    run<32>(a, B, H, pitch16, kb, sms); run<64>(a, B, H, pitch16, kb, sms); run<128>(a, B, H, pitch16, kb, sms);
    run<256>(a, B, H, pitch16, kb, sms); run<512>(a, B, H, pitch16, kb, sms);
  }
  return 0;
}
