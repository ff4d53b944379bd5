#!/bin/bash
# Final GPU call of the round: everything the driver runs (smoke, full GPU suite, both bench arms), the per-config
# and per-length sweeps for profiles/, the Bluestein chirp-table A/B, and two ncu captures reduced to JSON on the box.
#   gpurun --timeout 720 -- 'bash tools/gpu_final_check.sh'
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
T0=$(date +%s); el() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a gpurun_out/progress3.log; }
el start
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; el "smoke rc=$? $(tail -n 1 gpurun_out/smoke.log)"
timeout 300 python -m pytest tests -m gpu -q -x --durations=8 > gpurun_out/tests_full.log 2>&1; el "full gpu suite rc=$? $(tail -n 1 gpurun_out/tests_full.log)"
IMPULSE_FFT_BLUE_BK_SMEM=1 timeout 120 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "fused_bluestein or config1_and_3 or real_roundtrip" \
  > gpurun_out/t3_bks.log 2>&1; el "bluestein tests (bk in smem) rc=$? $(tail -n 1 gpurun_out/t3_bks.log)"
for mode in 0 1; do
  for wl in r2c_16384x4099_f64 c2r_16384x4099_f64; do
    IMPULSE_FFT_BLUE_BK_SMEM=$mode timeout 120 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu --workload $wl 2>/dev/null | \
      python -c "import sys,json; d=json.loads(sys.stdin.read()); print('bk_smem=$mode', '$wl', d['value'], d['ms_per_step'], d['roofline']['kernel'])" >> gpurun_out/ab_blue.txt
  done
done
el "bluestein A/B done"
timeout 200 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; el "bench default rc=$?"
timeout 200 python bench.py --impl reference > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; el "bench reference rc=$?"
for wl in r2c_1024x4096_f64 r2c_16384x1000_f64 c2r_16384x1000_f64 r2c_16384x3888_f64 c2r_16384x3888_f64 r2c_16384x4099_f64 \
          c2r_16384x4099_f64 c2c_16384x4096_c128 c2c_8192x8192_c128 c2c_131072x1024_c64 fft2_8192x8192_c128 filter2d_64x4096x4096_f32 fftconvolve_4096x16384_k257_f64; do
  timeout 120 python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu --workload $wl 2>/dev/null | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$wl', d['value'], d['ms_per_step'], d['roofline']['kernel'])" >> gpurun_out/sweep.txt
done
el "workload sweep done"
cap() {  # name skip kind dtype n rows
  timeout 120 ncu --set full --clock-control none --import-source on -k regex:fast3 -s $2 -c 1 -f -o /tmp/$1 python tools/run_one.py $3 $4 $5 $6 > gpurun_out/ncu_$1.log 2>&1
  python tools/ncu_summary.py /tmp/$1.ncu-rep gpurun_out/$1 >> gpurun_out/ncu_$1.log 2>&1
}
cap r01_fast3_c2r3888_pair 3 c2r f64 3888 16384
cap r01_fast3_c2c2048_pf 2 c2c f64 2048 16384
el "ncu done"
timeout 200 python tools/size_sweep.py --kinds c2c,r2c,c2r > gpurun_out/size_sweep.txt 2>&1; el "size sweep rc=$?"
cat gpurun_out/progress3.log; cat gpurun_out/ab_blue.txt; cat gpurun_out/sweep.txt; cat gpurun_out/bench_default.json; tail -n 12 gpurun_out/tests_full.log
