#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/sweep.txt
run() { # label, env..., workload
  label=$1; shift; wl=$1; shift
  out=$(env "$@" timeout 120 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu --workload $wl 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['kernel'])")
  echo "$wl $label $out" >> gpurun_out/sweep.txt
}
for wl in c2c_65536x1024_c128; do
  run nofast_default $wl IMPULSE_FFT_NO_FAST=1
  run nofast_L8 $wl IMPULSE_FFT_NO_FAST=1 IMPULSE_FFT_LINES=8
  run nofast_L2 $wl IMPULSE_FFT_NO_FAST=1 IMPULSE_FFT_LINES=2
  run nofast_L4_T128 $wl IMPULSE_FFT_NO_FAST=1 IMPULSE_FFT_LINES=4 IMPULSE_FFT_THREADS=128
done
for wl in r2c_1024x4096_f64 r2c_16384x1000_f64 r2c_16384x3888_f64 c2c_16384x4096_c128 r2c_16384x4099_f64 c2c_131072x1024_c64 fft2_8192x8192_c128; do
  run default $wl A=1
  run L1 $wl IMPULSE_FFT_LINES=1
  run L2 $wl IMPULSE_FFT_LINES=2
  run L8 $wl IMPULSE_FFT_LINES=8
  run T128 $wl IMPULSE_FFT_THREADS=128
done
cat gpurun_out/sweep.txt
