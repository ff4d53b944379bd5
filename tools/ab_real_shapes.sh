#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
out=gpurun_out/ab_real_shapes.txt; : > $out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "r2c_c2r_hermitian or roundtrip or randomized or highlevel" 2>&1 | tail -n 3 | tee -a $out
for w in 0 1; do
  IMPULSE_FFT_MORE_SHAPES=$w timeout 300 python tools/size_sweep.py --kinds r2c,c2r --dtypes f64,f32 --lengths 3072,4000,4374,6000,8000,13122 2>&1 | sed "s/^/more_shapes=$w /" | tee -a $out
done
