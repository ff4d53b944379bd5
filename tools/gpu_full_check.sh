#!/bin/bash
# Everything the round-end driver runs, plus the per-config sweep.  Usage (from the repo root):
#   gpurun --timeout 2400 -- 'bash tools/gpu_full_check.sh'
mkdir -p gpurun_out; rm -f gpurun_out/sweep.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -q -x --durations=5 > gpurun_out/tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/tests.log
timeout 300 python bench.py --impl reference > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
timeout 300 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
for wl in r2c_1024x4096_f64 r2c_16384x1000_f64 c2r_16384x1000_f64 r2c_16384x3888_f64 c2r_16384x3888_f64 r2c_16384x4099_f64 \
          c2r_16384x4099_f64 c2c_16384x4096_c128 c2c_8192x8192_c128 c2c_131072x1024_c64 fft2_8192x8192_c128 filter2d_64x4096x4096_f32 fftconvolve_4096x16384_k257_f64; do
  timeout 120 python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu --workload $wl 2>/dev/null | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$wl', d['value'], d['ms_per_step'], d['roofline']['kernel'])" >> gpurun_out/sweep.txt
done
timeout 300 python tools/size_sweep.py > gpurun_out/size_sweep.txt 2>&1
tail -3 gpurun_out/smoke.log; tail -8 gpurun_out/tests.log; cat gpurun_out/bench_default.json; cat gpurun_out/sweep.txt
