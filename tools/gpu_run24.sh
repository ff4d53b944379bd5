#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tests/sanitizer_cases.py > gpurun_out/san_plain.log 2>&1; echo "plain rc=$?" >> gpurun_out/san_plain.log
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python tests/sanitizer_cases.py > gpurun_out/san_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/san_memcheck.log
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python tests/sanitizer_cases.py > gpurun_out/san_racecheck.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/san_racecheck.log
tail -3 gpurun_out/san_plain.log; grep -E "ERROR SUMMARY|rc=|Invalid|hazard" gpurun_out/san_memcheck.log | sort | uniq -c | head; grep -E "RACECHECK SUMMARY|ERROR SUMMARY|rc=|hazard" gpurun_out/san_racecheck.log | sort | uniq -c | head
