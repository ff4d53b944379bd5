#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
out=gpurun_out/ab_tiny.txt; : > $out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tiny_rows or c2c_lengths or r2c_c2r_hermitian or roundtrip or randomized" 2>&1 | tail -n 3 | tee -a $out
for w in 1 0; do
  IMPULSE_FFT_NO_FAST=$w timeout 200 python tools/size_sweep.py --kinds c2c,r2c,c2r --dtypes f64,f32 --lengths 4,8,16 2>&1 | sed "s/^/no_fast=$w /" | tee -a $out
done
