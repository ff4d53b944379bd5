#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "register_kernels or golden or highlevel or r2c_c2r or config1 or filter or nd_and_strided" 2>&1 | tail -3
timeout 300 python tools/size_sweep.py > gpurun_out/size_sweep.txt 2>&1; grep -E "r2c f.. n= +(16|32|64|128|256|512|16384|32768) |c2c f32 n= +8192" gpurun_out/size_sweep.txt
timeout 300 compute-sanitizer --tool memcheck python tests/sanitizer_cases.py 2>&1 | tail -2
