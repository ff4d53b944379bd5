#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
out=gpurun_out/ab_colw3.txt; : > $out
for w in 0 1; do
  IMPULSE_FFT_COL_WHOLE=$w timeout 300 python tools/nd_sweep.py 2048 2>&1 | sed "s/^/colwhole=$w /" | tee -a $out
  IMPULSE_FFT_COL_WHOLE=$w timeout 300 python tools/nd_sweep.py 1024 2>&1 | sed "s/^/colwhole=$w /" | tee -a $out
done
