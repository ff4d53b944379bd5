#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -n 3 | tee gpurun_out/r02_pytest_gpu.txt
bash tools/gpu_ncu.sh r02_fastblue_r2c4099_tmem r2c f64 4099 16384 fastblue
grep -E "duration|dram__bytes_read|dram__bytes_write|fp64_cycles|issue_active|stalled_long|mio_throttle|registers" gpurun_out/r02_fastblue_r2c4099_tmem.md
out=gpurun_out/r02_compute_sanitizer.txt; : > $out
for tool in memcheck racecheck; do
  echo "== $tool: compute-sanitizer --tool $tool python tests/sanitizer_cases.py --new" >> $out
  timeout 1500 compute-sanitizer --tool $tool python tests/sanitizer_cases.py --new 2>&1 | grep -E "COMPUTE-SANITIZER|sanitizer cases|ERROR SUMMARY|RACECHECK SUMMARY|Error|error|hazard|Hazard" | head -40 >> $out
done
cat $out | cut -c1-200
