#!/bin/bash
# A/B of the column-kernel launch options on the filter workload (per-kernel times via ncu launch list)
for pf in 1 0; do for g1 in 0 1; do
  export IMPULSE_FFT_COL_PREFETCH=$pf IMPULSE_FFT_COL_GRID1D=$g1
  echo "== prefetch=$pf grid1d=$g1"
  python bench.py --workload filter2d_64x4096x4096_f32 --steps 10 --warmup 3 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('filter ms', d['ms_per_step'])"
  python bench.py --workload fft2_8192x8192_c128 --steps 20 --warmup 3 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('fft2 ms', d['ms_per_step'])"
  ncu --metrics gpu__time_duration.sum --clock-control none -c 30 --csv --log-file gpurun_out/ab_$pf$g1.csv python bench.py --workload filter2d_64x4096x4096_f32 --steps 2 --warmup 1 > /dev/null 2>&1
  python tools/launch_table.py gpurun_out/ab_$pf$g1.csv 2>&1 | tail -4 | cut -c1-120
done; done
