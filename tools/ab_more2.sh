#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
out=gpurun_out/ab_more_two_pass.txt; : > $out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "c2c_lengths or randomized or nd_and_strided or highlevel" 2>&1 | tail -n 3 | tee -a $out
for w in 0 1; do
  IMPULSE_FFT_MORE_SHAPES=$w timeout 300 python tools/size_sweep.py --kinds c2c --dtypes f64,f32 --lengths 50,72,81,96,192,200,400,576,729,900 2>&1 | sed "s/^/more_shapes=$w /" | tee -a $out
done
