#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "register_kernels or c2c_lengths or golden or highlevel or four_step or long_lines or column or config4 or nd_and_strided or r2c_c2r or config1" 2>&1 | tail -3
timeout 300 python tools/size_sweep.py > gpurun_out/size_sweep.txt 2>&1; grep -E "n= +(256|512|1024|2048|8192|16384|32768|65536|262144) " gpurun_out/size_sweep.txt
