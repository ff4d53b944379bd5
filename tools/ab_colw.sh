#!/bin/bash
# A/B of the whole-axis strided c2c (colconvw_kernel, plain mode) against the two-launch split; full GPU test suite first
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
out=gpurun_out/ab_colw.txt; : > $out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -n 4 | tee -a $out
for w in 0 1; do IMPULSE_FFT_COL_WHOLE=$w timeout 300 python tools/nd_sweep.py 2>&1 | sed "s/^/colwhole=$w /" | tee -a $out; done
