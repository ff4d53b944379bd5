#!/bin/bash
# host staging: ring pipeline (one stream per engine, 3 buffer pairs) vs the two-stream form; chunk sizes; whole-span staging
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
out=gpurun_out/ab_e2e2.txt; : > $out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "host or pipeline or staging or highlevel or nd_and_strided or errors" 2>&1 | tail -n 3 | tee -a $out
e2e() {  # label, env..., workload
  python bench.py --steps 5 --warmup 3 --no-cpu --no-configs --workload $1 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$2', '$1', 'e2e', d['e2e']['value'], 'GB/s', d['e2e']['ms_per_step'], 'ms')"
}
for ring in 0 1; do for mb in 16 32 64 128; do
  IMPULSE_FFT_STAGE_RING=$ring IMPULSE_FFT_STAGE_MB=$mb e2e c2c_65536x1024_c128 "ring=$ring mb=$mb" | tee -a $out
done; done
for ring in 0 1; do
  for w in r2c_16384x3888_f64 c2r_16384x4099_f64 r2c_1024x4096_f64 fft2_8192x8192_c128; do
    IMPULSE_FFT_STAGE_RING=$ring e2e $w "ring=$ring" | tee -a $out
  done
done
