#!/bin/bash
# FFT(b)/M in tensor memory (IMPULSE_FFT_BLUE_TMEM) on the four-pass fused Bluestein kernel: config 3c shapes
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
out=gpurun_out/ab_blue_tmem.txt
for w in 0 1 0 1; do
  for wl in r2c_16384x4099_f64 c2r_16384x4099_f64; do
    IMPULSE_FFT_BLUE_TMEM=$w timeout 200 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu --no-configs --workload $wl 2>gpurun_out/err.txt | \
      python -c "import sys,json; d=json.loads(sys.stdin.read()); print('blue_tmem=$w $wl', d['value'], 'GB/s', d['ms_per_step'], 'ms', d['roofline'].get('kernel'), d.get('accuracy',{}).get('pass'))" 2>&1 | tail -n 1 | tee -a $out
  done
done
tail -n 3 gpurun_out/err.txt
