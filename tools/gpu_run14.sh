#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/e2e2.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 30 --csv --log-file gpurun_out/launches_bench_default.csv python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_l.log 2>&1
for mb in 8 16 32 64 128; do
  IMPULSE_FFT_STAGE_MB=$mb timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$mb', d['e2e']['ms_per_step'], d['e2e']['value'])" >> gpurun_out/e2e2.txt
done
timeout 300 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
timeout 300 python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
cat gpurun_out/e2e2.txt; cat gpurun_out/bench_default.json; cut -c1-500 gpurun_out/bench_ref.json
