#!/bin/bash
# 2-GPU check: bench.py under torchrun (weak line + configs + strong c2c + slab fft2 with accuracy), the dist tests
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_gpu_dist.py -x -q -m gpu 2>&1 | tail -n 3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
echo "bench n2 rc=$?"; tail -c 1500 gpurun_out/bench_n2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n2.json').read().strip().splitlines()[-1])
print('main', d['value'], d['ms_per_step'], d['scaling'], d.get('e2e',{}).get('value'), d.get('accuracy'))
for k,v in d.get('configs',{}).items():
    print(k, json.dumps(v)[:700])
print('all pass', d.get('configs_accuracy_all_pass'))
PY
