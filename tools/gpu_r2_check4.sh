#!/bin/bash
# full GPU test suite + ncu captures of the staged three-pass kernels and the four-pass Bluestein
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
timeout 2400 python -m pytest tests -x -q -m gpu 2>&1 | tail -n 12
cap() {  # name kind dtype n rows regex
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$6 -s 2 -c 1 -f -o gpurun_out/$1 python tools/run_one.py $2 $3 $4 $5 > gpurun_out/$1.log 2>&1
  tail -n 2 gpurun_out/$1.log
}
cap r02_fast3_r2c3888_tma r2c f64 3888 16384 fast3_kernel
cap r02_fast3_r2c4096_tma_db r2c f64 4096 8192 fast3_kernel
cap r02_fast3_c2r4096_tma c2r f64 4096 8192 fast3_kernel
cap r02_fast3_r2c1000_tma_db r2c f64 1000 16384 fast3_kernel
cap r02_fastblue_r2c4099_four r2c f64 4099 16384 fastblue
