#!/bin/bash
# parity + timing + one ncu capture of the whole-axis convolution kernel:  tools/ab_convw2.sh <capture name>
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
name=${1:-r02_colconvw}
out=gpurun_out/$name.ab.txt; : > $out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "filter2d or convolve_axis or config5" 2>&1 | tail -n 3 | tee -a $out
for w in 1 1; do
  IMPULSE_FFT_CONV_WHOLE=$w timeout 120 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu --no-configs --workload filter2d_64x4096x4096_f32 2>/dev/null | \
      python -c "import sys,json; d=json.loads(sys.stdin.read()); print('whole=$w filter2d', d['value'], d['ms_per_step'])" | tee -a $out
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/$name.launches.csv python tools/run_filter.py 64 > /dev/null 2>&1
python - <<PY | tee -a $out
import csv
rows=[r for r in csv.reader(open("gpurun_out/$name.launches.csv")) if len(r)>5 and r[0].isdigit()]
for r in rows[-3:]:
    print("  ", r[4][:80], r[-1], r[-2])
PY
bash tools/gpu_ncu_cmd.sh $name colconvw 1 -- python tools/run_filter.py 64
