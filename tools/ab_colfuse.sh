#!/bin/bash
# fused column transform (both launches of the four-step split in one kernel, intermediate in L2): parity, then A/B
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "four_step or config4 or nd_and_strided or column_kernels or randomized or convolve_axis or filter2d or long_lines" 2>&1 | tail -n 5
for mode in 1 0; do
  for wl in fft2_8192x8192_c128 filter2d_64x4096x4096_f32; do
    IMPULSE_FFT_NO_COLFUSE=$mode timeout 120 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu --no-configs --workload $wl 2>/dev/null | \
      python -c "import sys,json; d=json.loads(sys.stdin.read()); print('nofuse=$mode', '$wl', d['value'], d['ms_per_step'], d['gpu_launches'], d['roofline']['kernel'])" | tee -a gpurun_out/ab_colfuse.txt
  done
  IMPULSE_FFT_NO_COLFUSE=$mode timeout 200 python tools/nd_sweep.py 2>&1 | sed "s/^/nofuse=$mode /" | tee -a gpurun_out/ab_colfuse.txt
done
