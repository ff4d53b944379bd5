#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x --durations=6 > gpurun_out/tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/tests.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --workload filter2d_64x4096x4096_f32 > gpurun_out/bench_filter.json 2> gpurun_out/bench_filter.err
timeout 300 python bench.py --steps 100 --warmup 10 > gpurun_out/bench_main.json 2> gpurun_out/bench_main.err
tail -14 gpurun_out/tests.log; cut -c1-600 gpurun_out/bench_filter.json; tail -3 gpurun_out/bench_filter.err; cut -c1-300 gpurun_out/bench_main.json
