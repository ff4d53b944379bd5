#!/bin/bash
mkdir -p gpurun_out
timeout 180 python bench.py --steps 100 --warmup 10 --no-e2e --no-cpu > gpurun_out/bench_tma.json 2> gpurun_out/bench_tma.err; echo "rc=$?" >> gpurun_out/bench_tma.err
IMPULSE_FFT_FAST_VARIANT=1 timeout 180 python bench.py --steps 100 --warmup 10 --no-e2e --no-cpu > gpurun_out/bench_notma.json 2> gpurun_out/bench_notma.err
timeout 900 python -m pytest tests -m gpu -q -x --durations=5 > gpurun_out/tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/tests.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fast2p -s 3 -c 1 -o gpurun_out/prof_fast2p_1024 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_full.log 2>&1
for w in r2c_1024x4096_f64 r2c_16384x1000_f64 r2c_16384x3888_f64 r2c_16384x4099_f64 c2r_16384x3888_f64 c2c_16384x4096_c128 c2c_8192x8192_c128 c2c_131072x1024_c64; do
  timeout 120 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu --workload $w >> gpurun_out/bench_others.json 2>> gpurun_out/bench_others.err
done
tail -8 gpurun_out/tests.log; cat gpurun_out/bench_tma.json gpurun_out/bench_notma.json | cut -c1-260; tail -2 gpurun_out/bench_tma.err; cut -c1-330 gpurun_out/bench_others.json
