#!/bin/bash
# 2 GPUs: multi-GPU tests (fused and pipelined slab exchange) and the slab timings
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
timeout 1200 python -m pytest tests/test_gpu_dist.py tests/test_gpu_dist_cabi.py -x -q -m gpu 2>&1 | tail -n 15 | tee gpurun_out/r02_pytest_gpu_n2.txt
for j in 1 2 4 8; do
IMPULSE_FFT_SLAB_PULL_BENCH=$j timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2954$j bench.py --gpus 2 --steps 20 --warmup 5 --no-e2e --no-cpu > gpurun_out/r02_bench_n2_pull$j.json 2> gpurun_out/r02_bench_n2.err
echo "bench rc=$?"; tail -c 300 gpurun_out/r02_bench_n2.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench_n2_pull$j.json').read().strip().splitlines()[-1])
s=d['configs']['4_fft2_8192x8192_slab']
print('J=$j', s['single_gpu_ms'], {k:(v['ms_per_step'], v['accuracy']['pass']) for k,v in s.items() if isinstance(v,dict)})
PY
done
