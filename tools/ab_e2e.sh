#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "register_kernels or c2c_lengths or host_staging or highlevel or golden" 2>&1 | tail -n 3
for ramp in 1 0 1 0; do
  IMPULSE_FFT_STAGE_RAMP=$ramp timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu --no-configs 2>/dev/null | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('ramp=$ramp e2e', d['e2e']['value'], d['e2e']['ms_per_step'])" | tee -a gpurun_out/ab_e2e.txt
done
timeout 200 python tools/size_sweep.py --kinds c2c --dtypes f32 --lengths 1536,2000,2187,3000,4000,6561 2>&1 | tee -a gpurun_out/ab_e2e.txt
