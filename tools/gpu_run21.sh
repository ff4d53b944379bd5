#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/sweep8.txt
timeout 600 python -m pytest tests -m gpu -q -x -k "fused_bluestein" > gpurun_out/tests_fb.log 2>&1; echo "tests rc=$?" >> gpurun_out/tests_fb.log
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/tests.log
run() { label=$1; shift; wl=$1; shift
  out=$(env "$@" timeout 120 python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu --workload $wl 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['kernel'])")
  echo "$wl $label $out" >> gpurun_out/sweep8.txt; }
for wl in r2c_16384x4099_f64 c2r_16384x4099_f64; do run default $wl A=1; run nofb $wl IMPULSE_FFT_NO_FASTBLUE=1; done
tail -12 gpurun_out/tests_fb.log; tail -5 gpurun_out/tests.log; cat gpurun_out/sweep8.txt
