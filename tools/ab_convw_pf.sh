#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
out=gpurun_out/ab_convw_pf.txt; : > $out
for w in 0 1 2; do
  for shape in "128 2048 2048 f32" "256 1024 1024 f32" "1024 512 512 f32" "64 2048 2048 f64" "128 1024 1024 f64" "512 512 512 f64"; do
    IMPULSE_FFT_CONVW_PF=$w timeout 120 python tools/time_filter.py $shape 2>&1 | tail -n 1 | sed "s/^/pf=$w /" | tee -a $out
  done
  IMPULSE_FFT_CONVW_PF=$w timeout 300 python tools/nd_sweep.py 1024 2>&1 | grep "f64 (128" | sed "s/^/pf=$w /" | tee -a $out
done
