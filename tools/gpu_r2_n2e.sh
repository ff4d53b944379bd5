#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
timeout 1200 python -m pytest tests/test_gpu_dist.py -x -q -m gpu 2>&1 | tail -n 3
for cfg in "1 32" "4 32" "4 64" "4 148" "8 64" "2 64"; do
set -- $cfg
IMPULSE_FFT_SLAB_COPY_CTAS=$2 IMPULSE_FFT_SLAB_PULL_BENCH=$1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --steps 20 --warmup 5 --no-e2e --no-cpu > gpurun_out/r02_bench_n2_pull.json 2> gpurun_out/r02_bench_n2.err
python - <<PY | tee -a gpurun_out/ab_slab_pull_n2.txt
import json
d=json.loads(open('gpurun_out/r02_bench_n2_pull.json').read().strip().splitlines()[-1])
s=d['configs']['4_fft2_8192x8192_slab']
print('J=$1 ctas=$2', s['single_gpu_ms'], {k:(v['ms_per_step'], v['accuracy']['pass']) for k,v in s.items() if isinstance(v,dict)})
PY
done
