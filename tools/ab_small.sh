#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "long_lines or all_lengths or randomized or four_step" 2>&1 | tail -3
timeout 300 python tools/size_sweep.py > gpurun_out/size_sweep.txt 2>&1; grep -E "n= +(32768|65536|100003|262144|1048576) " gpurun_out/size_sweep.txt
