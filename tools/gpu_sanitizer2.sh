#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
out=gpurun_out/r02_compute_sanitizer_whole.txt; : > $out
python tests/sanitizer_cases.py --whole 2>&1 | tail -n 2 >> $out
for tool in memcheck racecheck; do
  echo "== $tool: compute-sanitizer --tool $tool python tests/sanitizer_cases.py --whole" >> $out
  timeout 1500 compute-sanitizer --tool $tool python tests/sanitizer_cases.py --whole 2>&1 | grep -E "COMPUTE-SANITIZER|sanitizer cases|ERROR SUMMARY|RACECHECK SUMMARY|Error|error|hazard|Hazard" | head -40 >> $out
done
cat $out
