#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -n 3 | tee gpurun_out/r02_pytest_gpu.txt
bash tools/gpu_sanitizer2.sh 2>&1 | grep -E "SUMMARY|==" 
bash tools/gpu_ncu_cmd.sh r02_colconvw_1024_f64_tmem colconvw 1 -- python tools/nd_sweep.py 1024
grep -E "duration|issue_active|dram__bytes.sum.per|stalled_long|lg_throttle|warps_active" gpurun_out/r02_colconvw_1024_f64_tmem.md
