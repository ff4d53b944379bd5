#!/bin/bash
# round 2 record run on one B200: full GPU test suite, default bench line (with the configs block), reference arm, size sweep, launch list
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
timeout 2400 python -m pytest tests -x -q -m gpu 2>&1 | tail -n 6 | tee gpurun_out/r02_pytest_gpu.txt
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_bench_reference_arm.json 2>/dev/null
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err
echo "bench rc=$?"; tail -c 400 gpurun_out/r02_bench_default.err
timeout 900 python tools/size_sweep.py > gpurun_out/r02_size_sweep.txt 2>&1
timeout 300 python tools/nd_sweep.py > gpurun_out/r02_nd_sweep.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 0 -c 400 --csv --log-file gpurun_out/r02_launches_bench_default.csv python bench.py --steps 2 --warmup 3 --no-cpu > /dev/null 2>&1
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_default.json'))
print('main', d['value'], d['ms_per_step'], d['roofline']['frac'], d.get('e2e',{}).get('value'), d['cpu_baseline']['value'], d['accuracy']['pass'])
for k,v in d.get('configs',{}).items():
    print(k, v.get('ms_per_step'), v.get('GB/s'), v.get('frac_8TBps'), v.get('kernel'), 'acc', v.get('accuracy',{}).get('pass'), 'e2e', (v.get('e2e') or {}).get('value'), 'cpu', (v.get('cpu_baseline') or {}).get('value'), v.get('cuda_graph_200_iters'))
PY
