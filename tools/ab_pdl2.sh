#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "register_kernels or config2 or c2c_lengths or t2_known or readme or host_staging or randomized" 2>&1 | tail -n 3
for mode in 0 1 0 1; do
  for wl in c2c_65536x1024_c128 c2c_262144x256_c128 c2c_131072x1024_c64; do
    IMPULSE_FFT_PDL=$mode timeout 120 python bench.py --steps 50 --warmup 5 --no-e2e --no-cpu --no-configs --workload $wl 2>/dev/null | \
      python -c "import sys,json; d=json.loads(sys.stdin.read()); print('pdl=$mode', '$wl', d['value'], d['ms_per_step'], d['roofline']['kernel'])" | tee -a gpurun_out/ab_pdl.txt
  done
done
IMPULSE_FFT_PDL=1 timeout 200 python tools/size_sweep.py --kinds c2c --dtypes f64 --lengths 16,32,64,100,128,243,256,512,625,1024 2>&1 | sed 's/^/pdl=1 /' | tee -a gpurun_out/ab_pdl.txt
IMPULSE_FFT_PDL=0 timeout 200 python tools/size_sweep.py --kinds c2c --dtypes f64 --lengths 16,32,64,100,128,243,256,512,625,1024 2>&1 | sed 's/^/pdl=0 /' | tee -a gpurun_out/ab_pdl.txt
