#!/bin/bash
# whole-axis kernel with the next tile staged in tensor memory (IMPULSE_FFT_CONVW_TMEM): parity under the flag, then timings
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
out=gpurun_out/ab_convw_tmem.txt; : > $out
IMPULSE_FFT_CONVW_TMEM=1 timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "filter2d or convolve_axis or config5 or strided_axis_whole" 2>&1 | tail -n 4 | tee -a $out
IMPULSE_FFT_CONVW_TMEM=1 timeout 200 python tests/sanitizer_cases.py --whole 2>&1 | tail -n 1 | cut -c1-300 | tee -a $out
for w in 0 1 0 1; do
  for shape in "128 2048 2048 f32" "256 1024 1024 f32" "1024 512 512 f32" "64 2048 2048 f64" "128 1024 1024 f64" "512 512 512 f64"; do
    IMPULSE_FFT_CONVW_TMEM=$w timeout 120 python tools/time_filter.py $shape 2>&1 | tail -n 1 | sed "s/^/tmem=$w /" | tee -a $out
  done
  IMPULSE_FFT_CONVW_TMEM=$w timeout 200 python tools/nd_sweep.py 1024 2>&1 | grep "f64" | sed "s/^/tmem=$w /" | tee -a $out
  IMPULSE_FFT_CONV_WHOLE=2 IMPULSE_FFT_CONVW_TMEM=$w timeout 120 python tools/time_filter.py 64 4096 4096 f32 2>&1 | tail -n 1 | sed "s/^/tmem=$w whole=2 /" | tee -a $out
done
