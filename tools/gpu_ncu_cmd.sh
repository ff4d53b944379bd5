#!/bin/bash
# tools/gpu_ncu_cmd.sh name kernel-regex skip -- command...   : one ncu --set full capture summarised on the box
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
name=$1; rx=$2; skip=$3; shift 4
timeout 400 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -f -o /tmp/$name "$@" > gpurun_out/$name.log 2>&1
tail -n 1 gpurun_out/$name.log
python tools/ncu_summary.py /tmp/$name.ncu-rep gpurun_out/$name > /dev/null 2>&1
ncu -i /tmp/$name.ncu-rep --page source --csv > /tmp/$name.src.csv 2>/dev/null
python tools/ncu_hot_lines.py /tmp/$name.src.csv > gpurun_out/$name.hot.txt 2>&1
rm -f /tmp/$name.ncu-rep
