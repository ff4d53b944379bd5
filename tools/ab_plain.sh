#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "filter2d or column_kernels or convolve_axis or config5 or nd_and_strided or four_step or randomized or long_lines or cols_from_parts" 2>&1 | tail -n 3
for w in 0 1 0 1; do
  IMPULSE_FFT_COL_PLAIN=$w timeout 120 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu --no-configs --workload filter2d_64x4096x4096_f32 2>/dev/null | \
      python -c "import sys,json; d=json.loads(sys.stdin.read()); print('plain=$w filter2d', d['value'], d['ms_per_step'])" | tee -a gpurun_out/ab_plain.txt
done
for w in 0 1; do IMPULSE_FFT_COL_PLAIN=$w timeout 200 python tools/nd_sweep.py 2>&1 | sed "s/^/plain=$w /" | tee -a gpurun_out/ab_plain.txt; done
for w in 0 1; do IMPULSE_FFT_COL_PLAIN=$w timeout 200 python tools/size_sweep.py --kinds c2c --dtypes f64,f32 --lengths 16384,65536,262144,1048576 2>&1 | sed "s/^/plain=$w /" | tee -a gpurun_out/ab_plain.txt; done
