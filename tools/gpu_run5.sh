#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --durations=5 > gpurun_out/tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/tests.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu --workload fft2_8192x8192_c128 > gpurun_out/bench_fft2.json 2> gpurun_out/bench_fft2.err
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:impulse -c 12 --csv --log-file gpurun_out/launches_fft2.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --workload fft2_8192x8192_c128 > /dev/null 2>&1
tail -12 gpurun_out/tests.log; cut -c1-420 gpurun_out/bench_fft2.json; tail -3 gpurun_out/bench_fft2.err; grep -E "impulse" gpurun_out/launches_fft2.csv | awk -F'","' '{print $5, $13, $15}' | tail -12
