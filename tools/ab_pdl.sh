#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "register_kernels or config1_and_3 or r2c_c2r_hermitian or fft_filter2d or config4 or highlevel" 2>&1 | tail -n 3
for mode in 0 1 0 1; do
  for wl in r2c_1024x4096_f64 r2c_16384x1000_f64 c2r_16384x1000_f64 r2c_16384x3888_f64; do
    IMPULSE_FFT_PDL=$mode timeout 120 python bench.py --steps 50 --warmup 5 --no-e2e --no-cpu --no-configs --workload $wl 2>/dev/null | \
      python -c "import sys,json; d=json.loads(sys.stdin.read()); print('pdl=$mode', '$wl', d['value'], d['ms_per_step'], d['roofline']['kernel'])" | tee -a gpurun_out/ab_pdl.txt
  done
done
