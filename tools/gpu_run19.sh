#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fast3 -s 3 -c 1 -o gpurun_out/prof_fast3_1944 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --workload r2c_16384x3888_f64 > gpurun_out/ncu1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fast3 -s 3 -c 1 -o gpurun_out/prof_fast3_8192 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --workload c2c_8192x8192_c128 > gpurun_out/ncu2.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:colfast2 -s 2 -c 2 -o gpurun_out/prof_col_fft2 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --workload fft2_8192x8192_c128 > gpurun_out/ncu3.log 2>&1
ls gpurun_out/*.ncu-rep
