#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
out=gpurun_out/ab_cfg1.txt; : > $out
for v in 0 1 2; do for rep in 1 2; do
  IMPULSE_FFT_F3_TMA=$v timeout 200 python bench.py --steps 200 --warmup 20 --no-e2e --no-cpu --no-configs --workload r2c_1024x4096_f64 2>/dev/null | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('f3_tma=$v', d['value'], 'GB/s', d['ms_per_step'], 'ms', d['roofline'].get('kernel'), d.get('cuda_graph_200_iters'))" 2>&1 | tail -n 1 | tee -a $out
done; done
