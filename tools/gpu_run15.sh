#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/sweep5.txt
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/tests.log
run() { label=$1; shift; wl=$1; shift
  out=$(env "$@" timeout 120 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu --workload $wl 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['kernel'])")
  echo "$wl $label $out" >> gpurun_out/sweep5.txt; }
for wl in fft2_8192x8192_c128 filter2d_64x4096x4096_f32; do run default $wl A=1; run nocol $wl IMPULSE_FFT_NO_COLFAST=1; done
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"fft|fast|cmul|copy2d" -s 12 -c 8 --csv --log-file gpurun_out/launches_fft2.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --workload fft2_8192x8192_c128 > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"fft|fast|cmul|copy2d" -s 28 -c 10 --csv --log-file gpurun_out/launches_filter.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --workload filter2d_64x4096x4096_f32 > /dev/null 2>&1
tail -6 gpurun_out/tests.log; cat gpurun_out/sweep5.txt; python tools/launch_table.py gpurun_out/launches_fft2.csv | head -4; python tools/launch_table.py gpurun_out/launches_filter.csv | head -11
