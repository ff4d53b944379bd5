#!/bin/bash
# round 2 (second session) record run on one B200: full GPU test suite, default bench line with the configs block, reference arm,
# sweeps, launch lists, ncu captures of the whole-axis convolution kernel and the fused middle pass of config 5
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
timeout 2400 python -m pytest tests -x -q -m gpu 2>&1 | tail -n 6 | tee gpurun_out/r02_pytest_gpu.txt
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_bench_reference_arm.json 2>/dev/null
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err
echo "bench rc=$?"; tail -c 400 gpurun_out/r02_bench_default.err
timeout 900 python tools/size_sweep.py > gpurun_out/r02_size_sweep.txt 2>&1
timeout 300 python tools/nd_sweep.py > gpurun_out/r02_nd_sweep.txt 2>&1
for shape in "64 4096 4096 f32" "128 2048 2048 f32" "256 1024 1024 f32" "1024 512 512 f32" "64 2048 2048 f64" "128 1024 1024 f64" "512 512 512 f64"; do
  timeout 120 python tools/time_filter.py $shape 2>&1 | tail -n 1
done | tee gpurun_out/r02_filter_sizes.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 0 -c 400 --csv --log-file gpurun_out/r02_launches_bench_default.csv python bench.py --steps 2 --warmup 3 --no-cpu > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r02_filter_launches.csv python tools/run_filter.py 64 > /dev/null 2>&1
bash tools/gpu_ncu_cmd.sh r02_colconvw_1024_f32 colconvw 1 -- python tools/time_filter.py 256 1024 1024 f32
bash tools/gpu_ncu_cmd.sh r02_colconvw_1024_f64 colconvw 1 -- python tools/time_filter.py 128 1024 1024 f64
bash tools/gpu_ncu_cmd.sh r02_colconv2_mid_config5 colconv2 1 -- python tools/run_filter.py 64
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_default.json'))
print('main', d['value'], d['ms_per_step'], d['roofline']['frac'], d.get('e2e',{}).get('value'), d['cpu_baseline']['value'], d['accuracy']['pass'])
for k,v in d.get('configs',{}).items():
    print(k, v.get('ms_per_step'), v.get('GB/s'), v.get('frac_8TBps'), v.get('kernel'), 'acc', v.get('accuracy',{}).get('pass'), 'e2e', (v.get('e2e') or {}).get('value'), 'cpu', (v.get('cpu_baseline') or {}).get('value'), v.get('cuda_graph_200_iters'))
PY
