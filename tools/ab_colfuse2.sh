#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "four_step or config4 or nd_and_strided or column_kernels or randomized" 2>&1 | tail -n 3
for mode in 0 1; do
  IMPULSE_FFT_NO_COLFUSE=$mode timeout 120 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu --no-configs --workload fft2_8192x8192_c128 2>/dev/null | \
      python -c "import sys,json; d=json.loads(sys.stdin.read()); print('nofuse=$mode', d['value'], d['ms_per_step'], d['gpu_launches'], d['roofline']['kernel'])" | tee -a gpurun_out/ab_colfuse2.txt
done
bash tools/gpu_ncu_cmd.sh r02_colfuse_fft2_v6 colfuse2 2 -- python tools/run_fft2.py 8192
