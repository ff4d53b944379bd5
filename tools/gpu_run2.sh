#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --durations=5 > gpurun_out/tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/tests.log
timeout 600 python bench.py --steps 100 --warmup 10 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fast2 -s 3 -c 1 -o gpurun_out/prof_fast2_1024 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_full.log 2>&1
tail -12 gpurun_out/tests.log; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
