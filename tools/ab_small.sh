#!/bin/bash
for m in 0 100; do echo "carveout mode $m"; IMPULSE_FFT_CARVEOUT=$m timeout 300 python tools/size_sweep.py 2>&1 | grep -E "line_fft" | grep -E "n= +(100|243|625|1536|2000|3000|5000|6561|10000|12288|4099) "; done
