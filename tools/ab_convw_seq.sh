#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
out=gpurun_out/ab_convw_seq.txt; : > $out
for cfg in "1 0" "2 0" "2 1"; do
  set -- $cfg
  for shape in "64 4096 4096 f32" "128 2048 2048 f32" "64 2048 2048 f64"; do
    IMPULSE_FFT_CONV_WHOLE=2 IMPULSE_FFT_CONVW_GF=$1 IMPULSE_FFT_CONVW_SEQ=$2 timeout 120 python tools/time_filter.py $shape 2>&1 | tail -n 1 | sed "s/^/gf=$1 seq=$2 /" | tee -a $out
  done
done
