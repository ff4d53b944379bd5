#!/bin/bash
# Config 1 (1024 rows of 4096 real points, fp64) under the staging / exchange-buffer switches of the three-pass kernel:
# a short batch (3.5 rows per CTA) may prefer the variant with more resident CTAs.
# Usage (GPU box): bash tools/ab_cfg1_variants.sh > gpurun_out/r02_ab_cfg1_variants.txt
for rep in 1 2; do
  for v in "" "IMPULSE_FFT_F3_TMA=1" "IMPULSE_FFT_F3_TMA=0 IMPULSE_FFT_F3_DB=1" "IMPULSE_FFT_F3_TMA=0 IMPULSE_FFT_F3_DB=0"; do
    env $v timeout 60 python bench.py --workload r2c_1024x4096_f64 --steps 200 --warmup 20 --no-e2e --no-cpu --no-configs 2>/dev/null |
      python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('[$v]', d['config']['workload'], d['ms_per_step'], 'ms', d['value'], d['unit'], d['roofline'].get('kernel'))"
  done
done
