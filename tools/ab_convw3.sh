#!/bin/bash
# tile-order A/B of the whole-axis convolution kernel: GF adjacent groups on neighbouring CTAs
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
out=gpurun_out/ab_convw_gf.txt; : > $out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "filter2d or convolve_axis or config5" 2>&1 | tail -n 3 | tee -a $out
for gf in 1 2 4 8 16; do
  IMPULSE_FFT_CONVW_GF=$gf timeout 120 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu --no-configs --workload filter2d_64x4096x4096_f32 2>/dev/null | \
      python -c "import sys,json; d=json.loads(sys.stdin.read()); print('gf=$gf filter2d', d['value'], d['ms_per_step'])" | tee -a $out
done
