#!/bin/bash
# A/B: TMA + dynamic-claim two-pass kernel (variant 0) against the plain one (variant 1)
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "register_kernels or c2c_lengths or golden or highlevel" 2>&1 | tail -3
for v in 0 1; do for wl in c2c_131072x1024_c64 c2c_131072x512_c128 c2c_262144x256_c128; do
  IMPULSE_FFT_FAST_VARIANT=$v timeout 120 python bench.py --steps 50 --warmup 5 --no-e2e --no-cpu --workload $wl 2>/dev/null | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('variant $v', '$wl', d['value'], d['ms_per_step'], d['roofline']['kernel'])"
done; done
