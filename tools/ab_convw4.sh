#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
out=gpurun_out/ab_convw4.txt; : > $out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_signal.py -x -q -m gpu -k "filter2d or convolve_axis or config5 or signal" 2>&1 | tail -n 3 | tee -a $out
for w in 0 1 2; do
  for shape in "64 4096 4096 f32" "256 1024 1024 f32" "1024 512 512 f32" "128 1024 1024 f64" "512 512 512 f64" "128 2048 2048 f32" "64 2048 2048 f64"; do
    IMPULSE_FFT_CONV_WHOLE=$w timeout 120 python tools/time_filter.py $shape 2>&1 | tail -n 1 | sed "s/^/whole=$w /" | tee -a $out
  done
done
