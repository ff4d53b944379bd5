#!/bin/bash
# round 2, first check on the B200: full-size every-row parity, then the default bench line with the configs block
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "config or fused_bluestein or register_kernels" 2>&1 | tail -n 5
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
echo "bench rc=$?"; tail -c 600 gpurun_out/bench_default.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_default.json'))
print('main', d['value'], d['ms_per_step'], d['roofline']['frac'], d.get('e2e',{}).get('value'), d.get('accuracy'))
for k,v in d.get('configs',{}).items():
    print(k, {x:v.get(x) for x in ('ms_per_step','GB/s','frac_8TBps','kernel')}, 'acc', v.get('accuracy',{}).get('pass'), v.get('accuracy',{}).get('max_rel_l2_vs_oracle'), 'e2e', (v.get('e2e') or {}).get('value'), 'cpu', (v.get('cpu_baseline') or {}).get('value'), v.get('cuda_graph_200_iters'))
PY
