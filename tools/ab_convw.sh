#!/bin/bash
# A/B of the whole-axis convolution kernel (colconvw_kernel) against the three-launch scheme on config 5.
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
out=gpurun_out/ab_convw.txt; : > $out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_signal.py -x -q -m gpu -k "filter2d or convolve_axis or config5 or signal or fftconvolve" 2>&1 | tail -n 5 | tee -a $out
for w in 0 1 0 1; do
  IMPULSE_FFT_CONV_WHOLE=$w timeout 120 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu --no-configs --workload filter2d_64x4096x4096_f32 2>/dev/null | \
      python -c "import sys,json; d=json.loads(sys.stdin.read()); print('whole=$w filter2d', d['value'], d['ms_per_step'], d['roofline'].get('kernel'))" | tee -a $out
done
for w in 0 1; do
  IMPULSE_FFT_CONV_WHOLE=$w timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/convw_launches_$w.csv python tools/run_filter.py 64 > /dev/null 2>&1
  python - <<PY | tee -a $out
import csv
rows=[r for r in csv.reader(open("gpurun_out/convw_launches_$w.csv")) if len(r)>5 and r[0].isdigit()]
print("whole=$w launch list (last step):")
for r in rows[-7:]:
    print("  ", r[4][:70], r[-1], r[-2])
PY
done
