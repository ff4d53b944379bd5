#!/bin/bash
# A/B of IMPULSE_FFT_F3_GRID (even split of short batches over the CTAs) on config 1 and two neighbours.
# Usage (GPU box): bash tools/ab_f3grid.sh > gpurun_out/r02_ab_f3grid.txt
for rep in 1 2 3; do
  for m in 0 1; do
    for w in r2c_1024x4096_f64 r2c_16384x3888_f64; do
      IMPULSE_FFT_F3_GRID=$m timeout 120 python bench.py --workload $w --steps 200 --warmup 20 --no-e2e --no-cpu --no-configs 2>/dev/null |
        python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('grid_mode=$m', d['config']['workload'], d['ms_per_step'], 'ms', d['value'], d['unit'], (d.get('accuracy') or {}).get('pass'))"
    done
  done
done
