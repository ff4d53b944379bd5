#!/bin/bash
# 2 GPUs: the multi-GPU tests (torchrun worker + single-process C ABI driver) and the bench line under torchrun
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
timeout 1200 python -m pytest tests/test_gpu_dist.py tests/test_gpu_dist_cabi.py -x -q -m gpu 2>&1 | tail -n 4 | tee gpurun_out/r02_pytest_gpu_n2.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err
echo "bench rc=$?"; tail -c 300 gpurun_out/r02_bench_n2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_n2.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value','n_gpus','ms_per_step','scaling')}, d.get('accuracy',{}).get('pass'))
for k in ('strong','fft2_slab','configs'):
    if k in d: print(k, json.dumps(d[k])[:900])
PY
