#!/bin/bash
# A/B of the complex128 column kernel: lines per CTA (8 / 16) x software pipelining (IMPULSE_FFT_COL_PIPE groups)
for l in 0 1; do for g in 0 2; do
  export IMPULSE_FFT_COL_PIPE=$g IMPULSE_FFT_COL_LPC16=$l
  a=$(python bench.py --workload fft2_8192x8192_c128 --steps 30 --warmup 5 --no-e2e --no-cpu 2>/dev/null | tail -1 | python -c "import json,sys; print(json.loads(sys.stdin.read())['ms_per_step'])")
  b=$(python tools/size_sweep.py 2>&1 | grep -E "c2c f64 n= +(16384|65536|262144) " | awk '{print $4"@"$5}' | tr '\n' ' ')
  echo "LPC16=$l G=$g fft2 ms $a | c2c 16384/65536/262144: $b"
done; done
