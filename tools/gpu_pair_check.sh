#!/bin/bash
# One GPU call for the in-register pair variants of the real three-pass kernels (r2c post-twiddle in pass 3,
# c2r pre-twiddle in pass 1): parity with the pair kernels on (default) and off, A/B throughput, the headline
# bench, ncu captures, then the full GPU suite with what time is left.  Everything lands in gpurun_out/.
#   gpurun --timeout 930 -- 'bash tools/gpu_pair_check.sh'
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
T0=$(date +%s); el() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a gpurun_out/progress.log; }
el start
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_signal.py -x -q -m gpu \
  -k "register_kernels or config1_and_3 or r2c_c2r_hermitian or fft_filter2d or randomized or signal or golden or readme" \
  > gpurun_out/t_pair_on.log 2>&1; el "targeted tests (pair on) rc=$? $(tail -1 gpurun_out/t_pair_on.log)"
WLS="r2c_1024x4096_f64 r2c_16384x1000_f64 c2r_16384x1000_f64 r2c_16384x3888_f64 c2r_16384x3888_f64 filter2d_64x4096x4096_f32"
for mode in 1 0; do
  for wl in $WLS; do
    IMPULSE_FFT_R2C_PAIR=$mode IMPULSE_FFT_C2R_PAIR=$mode timeout 120 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu --workload $wl 2>/dev/null | \
      python -c "import sys,json; d=json.loads(sys.stdin.read()); print('pair=$mode', '$wl', d['value'], d['ms_per_step'], d['roofline']['kernel'])" >> gpurun_out/ab_pair.txt
  done
  IMPULSE_FFT_R2C_PAIR=$mode IMPULSE_FFT_C2R_PAIR=$mode timeout 120 python tools/size_sweep.py --kinds r2c,c2r --lengths 512,1000,2048,3888,4096 2>&1 | sed "s/^/pair=$mode /" >> gpurun_out/ab_pair.txt
done
el "A/B done"
timeout 200 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; el "bench default rc=$?"
# ncu: --set full of the pair kernels on config 3b (r2c and c2r 16384 x 3888) and config 1's shape (r2c 4096)
timeout 120 ncu --set full --clock-control none --import-source on -k regex:fast3 -s 2 -c 1 -f -o gpurun_out/prof_r2c3888_pair python tools/run_one.py r2c f64 3888 16384 > gpurun_out/ncu1.log 2>&1
timeout 120 ncu --set full --clock-control none --import-source on -k regex:fast3 -s 3 -c 1 -f -o gpurun_out/prof_c2r3888_pair python tools/run_one.py c2r f64 3888 16384 > gpurun_out/ncu3.log 2>&1
el "ncu done"
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; el "smoke rc=$?"
timeout ${FULL_TIMEOUT:-420} python -m pytest tests -m gpu -q -x --durations=8 > gpurun_out/tests_full.log 2>&1; el "full gpu suite rc=$? $(tail -1 gpurun_out/tests_full.log)"
IMPULSE_FFT_R2C_PAIR=0 IMPULSE_FFT_C2R_PAIR=0 timeout 200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu \
  -k "register_kernels or config1_and_3" > gpurun_out/t_pair_off.log 2>&1; el "targeted tests (pair off) rc=$? $(tail -1 gpurun_out/t_pair_off.log)"
timeout 200 python bench.py --impl reference > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; el "bench reference rc=$?"
cat gpurun_out/progress.log; cat gpurun_out/ab_pair.txt; cat gpurun_out/bench_default.json; tail -5 gpurun_out/t_pair_on.log
