#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
out=gpurun_out/ab_e2e_n2.txt; : > $out
for ring in 0 1 0 1; do
IMPULSE_FFT_STAGE_RING=$ring timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2957$ring bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu --no-configs 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ring=$ring n=2 e2e', d['e2e']['value'], 'GB/s', d['e2e']['ms_per_step'], 'ms')" | tee -a $out
done
