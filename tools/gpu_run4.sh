#!/bin/bash
mkdir -p gpurun_out
timeout 180 python bench.py --steps 200 --warmup 10 --no-e2e --no-cpu > gpurun_out/bench_dyn.json 2> gpurun_out/bench_dyn.err; echo "rc=$?" >> gpurun_out/bench_dyn.err
timeout 600 python -m pytest tests -m gpu -q -x -k "c2c_lengths or config2 or t2_known or golden" > gpurun_out/tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/tests.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fast2p -s 3 -c 1 -o gpurun_out/prof_fast2p_dyn python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_full.log 2>&1
tail -4 gpurun_out/tests.log; cat gpurun_out/bench_dyn.json | cut -c1-400; tail -2 gpurun_out/bench_dyn.err
