#!/bin/bash
# ncu --set full captures summarised ON the GPU box (the .ncu-rep files with sources exceed what gpurun brings back):
#   tools/gpu_ncu.sh name kind dtype n rows kernel-regex [more sets of 6 ...]
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
while [ $# -ge 6 ]; do
  name=$1; kind=$2; dt=$3; n=$4; rows=$5; rx=$6; shift 6
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$rx -s 2 -c 1 -f -o /tmp/$name python tools/run_one.py $kind $dt $n $rows > gpurun_out/$name.log 2>&1
  tail -n 1 gpurun_out/$name.log
  python tools/ncu_summary.py /tmp/$name.ncu-rep gpurun_out/$name > /dev/null 2>&1
  # hottest source lines by stall samples (needs -lineinfo): file:line, samples, top stall reasons
  ncu -i /tmp/$name.ncu-rep --page source --csv > /tmp/$name.src.csv 2>/dev/null
  python tools/ncu_hot_lines.py /tmp/$name.src.csv > gpurun_out/$name.hot.txt 2>&1
  rm -f /tmp/$name.ncu-rep
done
