#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
for g in 0 2 4 8; do
  IMPULSE_FFT_COL_PIPE_F32=$g timeout 120 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu --no-configs --workload filter2d_64x4096x4096_f32 2>/dev/null | \
      python -c "import sys,json; d=json.loads(sys.stdin.read()); print('pipe32=$g', d['value'], d['ms_per_step'], d['gpu_launches'], d['roofline']['kernel'])" | tee -a gpurun_out/ab_pipe32.txt
done
IMPULSE_FFT_COL_PIPE_F32=4 timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "filter2d or column_kernels or nd_and_strided or randomized" 2>&1 | tail -n 3
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -s 15 -c 5 --csv --log-file gpurun_out/filter_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-configs --workload filter2d_64x4096x4096_f32 > /dev/null 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/filter_launches.csv')))
h=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]
H=rows[h]
for r in rows[h+1:]:
    if len(r)==len(H): print(r[H.index('Kernel Name')][:70], r[H.index('Metric Name')], r[H.index('Metric Value')])
PY
