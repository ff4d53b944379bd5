#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
out=gpurun_out/ab_colw2.txt; : > $out
for w in 0 1; do
  IMPULSE_FFT_CW_NARROW=$w timeout 300 python tools/nd_sweep.py 2>&1 | grep "1024, 1024" | sed "s/^/narrow=$w /" | tee -a $out
  for shape in "256 1024 1024 f32" "128 1024 1024 f64"; do
    IMPULSE_FFT_CW_NARROW=$w timeout 120 python tools/time_filter.py $shape 2>&1 | tail -n 1 | sed "s/^/narrow=$w /" | tee -a $out
  done
done
