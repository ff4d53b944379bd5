#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
for lag in 2 3 4 5 6 8 12; do
  IMPULSE_FFT_FUSE_LAG=$lag timeout 120 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu --no-configs --workload fft2_8192x8192_c128 2>/dev/null | \
      python -c "import sys,json; d=json.loads(sys.stdin.read()); print('lag=$lag', d['value'], d['ms_per_step'])" | tee -a gpurun_out/ab_lag.txt
done
