#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/sweep3.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "register_kernels or r2c_c2r" > gpurun_out/tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/tests.log
run() { label=$1; shift; wl=$1; shift
  out=$(env "$@" timeout 120 python bench.py --steps 50 --warmup 5 --no-e2e --no-cpu --workload $wl 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['kernel'])")
  echo "$wl $label $out" >> gpurun_out/sweep3.txt; }
for wl in r2c_1024x4096_f64 c2c_16384x4096_c128 c2c_8192x8192_c128; do run fast3 $wl A=1; done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fast3 -s 3 -c 1 -o gpurun_out/prof_fast3_4096b python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --workload c2c_16384x4096_c128 > gpurun_out/ncu1.log 2>&1
tail -4 gpurun_out/tests.log; cat gpurun_out/sweep3.txt
