#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fastblue -s 3 -c 1 -o gpurun_out/prof_fastblue_4099 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --workload r2c_16384x4099_f64 > gpurun_out/ncu1.log 2>&1
ls gpurun_out/prof_fastblue_4099.ncu-rep
