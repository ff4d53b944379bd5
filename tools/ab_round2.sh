#!/bin/bash
# A/B of the variants prepared at the end of round 1 (env-gated, emulation-validated, not yet measured):
#   IMPULSE_FFT_F3_DB=1      second exchange buffer in the three-pass kernels (two barriers per row instead of four)
#   IMPULSE_FFT_FAST4=1      c2c rows of 8192 points on the four-pass 512-thread core (config 4's row pass)
#   IMPULSE_FFT_ZEROCOPY=1     host-pointer calls on pinned memory run the kernel directly on the mapped host buffers (e2e)
#   IMPULSE_FFT_MORE_SHAPES=1  three-pass shapes for 1536 / 2000 / 4000 complex points (today: generic engine)
#   IMPULSE_FFT_BLUE_F32=1   fused Bluestein in float32 (today: generic engine, 3 % of the roofline at 4099)
#   IMPULSE_FFT_BLUE_FOUR=1  fused Bluestein on the four-pass core (512 threads x 16 points: 16 warps per SM instead of 8)
# Parity under the flag first, then throughput on the config 1 / 3 / 5 workloads and a length sweep.
#   gpurun --timeout 400 -- 'bash tools/ab_round2.sh'
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
IMPULSE_FFT_F3_DB=1 timeout 120 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "register_kernels or config1_and_3 or r2c_c2r_hermitian or fft_filter2d" 2>&1 | tail -n 3
for mode in 0 1; do
  for wl in r2c_1024x4096_f64 r2c_16384x1000_f64 c2r_16384x1000_f64 r2c_16384x3888_f64 c2r_16384x3888_f64 filter2d_64x4096x4096_f32; do
    IMPULSE_FFT_F3_DB=$mode timeout 120 python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu --workload $wl 2>/dev/null | \
      python -c "import sys,json; d=json.loads(sys.stdin.read()); print('db=$mode', '$wl', d['value'], d['ms_per_step'], d['roofline']['kernel'])" | tee -a gpurun_out/ab_db.txt
  done
  IMPULSE_FFT_F3_DB=$mode timeout 120 python tools/size_sweep.py --kinds c2c,r2c,c2r --dtypes f64,f32 --lengths 1000,2048,3888,4096 2>&1 | sed "s/^/db=$mode /" | tee -a gpurun_out/ab_db.txt
done
IMPULSE_FFT_BLUE_FOUR=1 timeout 120 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "fused_bluestein or config1_and_3 or real_roundtrip" 2>&1 | tail -n 3
for mode in 0 1; do
  for wl in r2c_16384x4099_f64 c2r_16384x4099_f64; do
    IMPULSE_FFT_BLUE_FOUR=$mode timeout 120 python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu --workload $wl 2>/dev/null | \
      python -c "import sys,json; d=json.loads(sys.stdin.read()); print('four=$mode', '$wl', d['value'], d['ms_per_step'], d['roofline']['kernel'])" | tee -a gpurun_out/ab_blue4.txt
  done
done
IMPULSE_FFT_FAST4=1 timeout 120 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "register_kernels or config4 or nd_and_strided" 2>&1 | tail -n 3
for mode in 0 1; do
  for wl in c2c_8192x8192_c128 fft2_8192x8192_c128; do
    IMPULSE_FFT_FAST4=$mode timeout 120 python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu --workload $wl 2>/dev/null | \
      python -c "import sys,json; d=json.loads(sys.stdin.read()); print('fast4=$mode', '$wl', d['value'], d['ms_per_step'], d['roofline']['kernel'])" | tee -a gpurun_out/ab_fast4.txt
  done
done
IMPULSE_FFT_BLUE_F32=1 timeout 120 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "fused_bluestein or randomized" 2>&1 | tail -n 3
for mode in 0 1; do
  IMPULSE_FFT_BLUE_F32=$mode timeout 120 python tools/size_sweep.py --kinds c2c,r2c,c2r --dtypes f32 --lengths 1021,2051,4099 2>&1 | sed "s/^/blue_f32=$mode /" | tee -a gpurun_out/ab_blue_f32.txt
done
IMPULSE_FFT_MORE_SHAPES=1 timeout 120 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "register_kernels or c2c_lengths" 2>&1 | tail -n 3
for mode in 0 1; do
  IMPULSE_FFT_MORE_SHAPES=$mode timeout 120 python tools/size_sweep.py --kinds c2c,r2c,c2r --dtypes f64 --lengths 1536,2000,3072,4000,8000 2>&1 | sed "s/^/more_shapes=$mode /" | tee -a gpurun_out/ab_more_shapes.txt
done
IMPULSE_FFT_ZEROCOPY=1 timeout 120 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "host_staging or highlevel or golden" 2>&1 | tail -n 3
for mode in 0 1; do
  IMPULSE_FFT_ZEROCOPY=$mode timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu 2>/dev/null | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('zerocopy=$mode e2e', d['e2e'])" | tee -a gpurun_out/ab_zerocopy.txt
done
