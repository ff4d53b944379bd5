#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/sweep2.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "register_kernels or r2c_c2r or c2c_lengths or config1" > gpurun_out/tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/tests.log
run() { label=$1; shift; wl=$1; shift
  out=$(env "$@" timeout 120 python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu --workload $wl 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['kernel'])")
  echo "$wl $label $out" >> gpurun_out/sweep2.txt; }
for wl in r2c_1024x4096_f64 c2c_16384x4096_c128 c2c_8192x8192_c128 fft2_8192x8192_c128 filter2d_64x4096x4096_f32; do
  run fast3 $wl A=1
  run nofast3 $wl IMPULSE_FFT_NO_FAST3=1
done
timeout 120 python bench.py --steps 20 --warmup 3 --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('e2e', d['e2e'])" >> gpurun_out/sweep2.txt
tail -8 gpurun_out/tests.log; cat gpurun_out/sweep2.txt
