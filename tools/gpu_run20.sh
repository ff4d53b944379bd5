#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/sweep7.txt
timeout 900 python -m pytest tests -m gpu -q -x -k "register_kernels or column or four_step or config4 or filter or r2c_c2r" > gpurun_out/tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/tests.log
run() { label=$1; shift; wl=$1; shift
  out=$(env "$@" timeout 120 python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu --workload $wl 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['kernel'])")
  echo "$wl $label $out" >> gpurun_out/sweep7.txt; }
for wl in r2c_16384x3888_f64 c2r_16384x3888_f64 c2c_8192x8192_c128 fft2_8192x8192_c128 filter2d_64x4096x4096_f32; do run default $wl A=1; done
tail -4 gpurun_out/tests.log; cat gpurun_out/sweep7.txt
