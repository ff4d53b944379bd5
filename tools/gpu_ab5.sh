#!/bin/bash
# A/B of three launch variants (all env-gated, defaults untouched) + an ncu summary of the Bluestein kernel with the
# chirp table in shared memory.
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
T0=$(date +%s); el() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a gpurun_out/progress5.log; }
el start
IMPULSE_FFT_F3_512P=1 IMPULSE_FFT_F3_MINB4=1 timeout 100 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "register_kernels or r2c_c2r_hermitian or config1_and_3" \
  > gpurun_out/t5_variants.log 2>&1; el "parity with the variants on rc=$? $(tail -n 1 gpurun_out/t5_variants.log)"
for v in "" "IMPULSE_FFT_F3_512P=1 IMPULSE_FFT_F3_MINB4=1"; do
  env $v timeout 60 python tools/size_sweep.py --kinds c2c,r2c,c2r --dtypes f64,f32 --lengths 1024,2048,4096 2>&1 | sed "s/^/[${v:-default}] /" >> gpurun_out/ab5.txt
  env $v timeout 60 python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu --workload r2c_1024x4096_f64 2>/dev/null | \
      python -c "import sys,json; d=json.loads(sys.stdin.read()); print('[${v:-default}]', 'r2c_1024x4096_f64', d['value'], d['ms_per_step'], d['roofline']['kernel'])" >> gpurun_out/ab5.txt
done
el "A/B done"
timeout 100 ncu --set full --clock-control none --import-source on -k regex:fastblue -s 2 -c 1 -f -o /tmp/blue python tools/run_one.py r2c f64 4099 16384 > gpurun_out/ncu_blue.log 2>&1
python tools/ncu_summary.py /tmp/blue.ncu-rep gpurun_out/r01_fastblue_r2c4099_bk_smem >> gpurun_out/ncu_blue.log 2>&1
el "ncu done"
cat gpurun_out/progress5.log gpurun_out/ab5.txt
