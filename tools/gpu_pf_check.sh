#!/bin/bash
# Second GPU call of the round: ncu summaries of the pair kernels (reduced to JSON on the box: the reports
# themselves are too large to travel back), A/B of the register-prefetch variants (IMPULSE_FFT_F3_PF), parity
# with them on, and the float32 instances of the 500/1000/1944-point shapes.
#   gpurun --timeout 600 -- 'bash tools/gpu_pf_check.sh'
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
T0=$(date +%s); el() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a gpurun_out/progress2.log; }
el start
timeout 200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "register_kernels or config1_and_3 or r2c_c2r_hermitian or randomized" \
  > gpurun_out/t2_default.log 2>&1; el "targeted tests (default) rc=$? $(tail -1 gpurun_out/t2_default.log)"
IMPULSE_FFT_F3_PF=1 timeout 200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "register_kernels or config1_and_3 or fft_filter2d" \
  > gpurun_out/t2_pf.log 2>&1; el "targeted tests (PF=1) rc=$? $(tail -1 gpurun_out/t2_pf.log)"
WLS="r2c_1024x4096_f64 r2c_16384x1000_f64 c2r_16384x1000_f64 r2c_16384x3888_f64 c2r_16384x3888_f64 c2c_16384x4096_c128 filter2d_64x4096x4096_f32"
for mode in 0 1; do
  for wl in $WLS; do
    IMPULSE_FFT_F3_PF=$mode timeout 120 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu --workload $wl 2>/dev/null | \
      python -c "import sys,json; d=json.loads(sys.stdin.read()); print('pf=$mode', '$wl', d['value'], d['ms_per_step'], d['roofline']['kernel'])" >> gpurun_out/ab_pf.txt
  done
  IMPULSE_FFT_F3_PF=$mode timeout 120 python tools/size_sweep.py --kinds c2c,r2c,c2r --dtypes f64 --lengths 2048,4096 2>&1 | sed "s/^/pf=$mode /" >> gpurun_out/ab_pf.txt
done
timeout 120 python tools/size_sweep.py --kinds c2c,r2c,c2r --dtypes f32 --lengths 500,1000,1944,2000,3888 2>&1 | sed "s/^/f32 /" >> gpurun_out/ab_pf.txt
el "A/B done"
cap() {  # name skip kind dtype n rows
  timeout 120 ncu --set full --clock-control none --import-source on -k regex:fast3 -s $2 -c 1 -f -o /tmp/$1 python tools/run_one.py $3 $4 $5 $6 > gpurun_out/ncu_$1.log 2>&1
  python tools/ncu_summary.py /tmp/$1.ncu-rep gpurun_out/$1 >> gpurun_out/ncu_$1.log 2>&1
  ncu -i /tmp/$1.ncu-rep --page source --csv 2>/dev/null | head -400 > gpurun_out/$1.source.csv
}
cap r01_fast3_r2c3888_pair 2 r2c f64 3888 16384
cap r01_fast3_c2r3888_pair 3 c2r f64 3888 16384
cap r01_fast3_r2c4096_pair 2 r2c f64 4096 8192
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r01_launches_r2c_1024x4096.csv python bench.py --workload r2c_1024x4096_f64 --steps 5 --warmup 3 --no-e2e --no-cpu > /dev/null 2>&1
el "ncu done"
cat gpurun_out/progress2.log; cat gpurun_out/ab_pf.txt; tail -3 gpurun_out/t2_default.log gpurun_out/t2_pf.log; ls -la gpurun_out | head -40
