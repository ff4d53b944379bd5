#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "register_kernels or config1 or config2 or r2c_c2r or filter or config5 or randomized" 2>&1 | tail -3
for wl in r2c_1024x4096_f64 r2c_16384x1000_f64 c2r_16384x1000_f64 r2c_16384x3888_f64 c2r_16384x3888_f64 filter2d_64x4096x4096_f32; do
  timeout 120 python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu --workload $wl 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$wl', d['value'], d['ms_per_step'])"
done
python tools/size_sweep.py 2>&1 | grep -E "r2c f.. n= +(512|1024|2048|4096|8192|16384) "
