#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_gpu_dist_cabi.py tests/test_gpu_dist.py -x -q -m gpu 2>&1 | tail -n 15
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
echo "bench n2 rc=$?"; tail -c 800 gpurun_out/bench_n2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n2.json').read().strip().splitlines()[-1])
print('main', d['value'], d['ms_per_step'], d['scaling'], d.get('e2e',{}).get('value'))
for k in ('2_c2c_65536x1024_strong','4_fft2_8192x8192_slab'):
    print(k, json.dumps(d['configs'][k])[:1800])
print('all pass', d.get('configs_accuracy_all_pass'))
PY
