#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fast3 -s 3 -c 1 -o gpurun_out/prof_fast3_4096 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --workload c2c_16384x4096_c128 > gpurun_out/ncu1.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"fft|fast|cmul|copy2d" -s 12 -c 8 --csv --log-file gpurun_out/launches_fft2.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --workload fft2_8192x8192_c128 > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"fft|fast|cmul|copy2d" -s 28 -c 14 --csv --log-file gpurun_out/launches_filter.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --workload filter2d_64x4096x4096_f32 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:line_fft -s 9 -c 2 -o gpurun_out/prof_generic_col python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --workload fft2_8192x8192_c128 > gpurun_out/ncu2.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
