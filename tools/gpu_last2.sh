#!/bin/bash
# Closing GPU call: Bluestein with the multipliers requested early (A/B + parity under the flag), then smoke and the
# full GPU suite on the committed defaults (1024-point real rows on the paired 8x8x8 shape).
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
T0=$(date +%s); el() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a gpurun_out/progress6.log; }
el start
IMPULSE_FFT_BLUE_BF_EARLY=1 timeout 90 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "fused_bluestein or config1_and_3 or real_roundtrip" \
  > gpurun_out/t6_bfe.log 2>&1; el "bluestein tests (bf early) rc=$? $(tail -n 1 gpurun_out/t6_bfe.log)"
for mode in 0 1; do
  for wl in r2c_16384x4099_f64 c2r_16384x4099_f64; do
    IMPULSE_FFT_BLUE_BF_EARLY=$mode timeout 60 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu --workload $wl 2>/dev/null | \
      python -c "import sys,json; d=json.loads(sys.stdin.read()); print('bf_early=$mode', '$wl', d['value'], d['ms_per_step'], d['roofline']['kernel'])" >> gpurun_out/ab6.txt
  done
done
el "A/B done"
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke6.log 2>&1; el "smoke rc=$? $(tail -n 1 gpurun_out/smoke6.log)"
timeout 150 python -m pytest tests -m gpu -q -x > gpurun_out/tests_full6.log 2>&1; el "full gpu suite rc=$? $(tail -n 1 gpurun_out/tests_full6.log)"
cat gpurun_out/progress6.log gpurun_out/ab6.txt
