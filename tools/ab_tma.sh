#!/bin/bash
# A/B of the staged (TMA) three-pass kernels: IMPULSE_FFT_F3_TMA = 0 (direct loads) / 1 (staged) / 2 (staged + second exchange buffer)
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
for mode in 1 2; do
  IMPULSE_FFT_F3_TMA=$mode timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "register_kernels or config1_and_3 or r2c_c2r_hermitian or fft_filter2d or randomized" 2>&1 | tail -n 3
done
for mode in 0 1 2; do
  for wl in r2c_1024x4096_f64 r2c_16384x1000_f64 c2r_16384x1000_f64 r2c_16384x3888_f64 c2r_16384x3888_f64 filter2d_64x4096x4096_f32; do
    IMPULSE_FFT_F3_TMA=$mode timeout 120 python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu --no-configs --workload $wl 2>/dev/null | \
      python -c "import sys,json; d=json.loads(sys.stdin.read()); print('tma=$mode', '$wl', d['value'], d['ms_per_step'], d['roofline']['kernel'])" | tee -a gpurun_out/ab_tma.txt
  done
  IMPULSE_FFT_F3_TMA=$mode timeout 120 python tools/size_sweep.py --kinds c2c,r2c,c2r --dtypes f64,f32 --lengths 1000,2048,3888,4096 2>&1 | sed "s/^/tma=$mode /" | tee -a gpurun_out/ab_tma.txt
done
