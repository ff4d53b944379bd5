#!/bin/bash
# Last GPU call of round 1: the defaults as committed (chirp table in shared memory on) through smoke + the full GPU
# suite, compute-sanitizer memcheck and racecheck over the new kernel variants, final numbers for config 3c.
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
T0=$(date +%s); el() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a gpurun_out/progress4.log; }
el start
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke4.log 2>&1; el "smoke rc=$? $(tail -n 1 gpurun_out/smoke4.log)"
timeout 200 python -m pytest tests -m gpu -q -x > gpurun_out/tests_full4.log 2>&1; el "full gpu suite rc=$? $(tail -n 1 gpurun_out/tests_full4.log)"
for wl in r2c_16384x4099_f64 c2r_16384x4099_f64; do
  timeout 120 python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu --workload $wl 2>/dev/null | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$wl', d['value'], d['ms_per_step'], d['roofline']['kernel'])" >> gpurun_out/sweep4.txt
done
timeout 60 python tools/size_sweep.py --kinds c2c --dtypes f64 --lengths 2048,4099 >> gpurun_out/sweep4.txt 2>&1
el "config 3c numbers done"
timeout 110 compute-sanitizer --tool memcheck --error-exitcode 9 python tests/sanitizer_cases.py --new > gpurun_out/sanitizer_memcheck_new.log 2>&1; el "memcheck rc=$? $(tail -n 2 gpurun_out/sanitizer_memcheck_new.log | tr '\n' ' ')"
timeout 100 compute-sanitizer --tool racecheck --error-exitcode 9 python tests/sanitizer_cases.py --new > gpurun_out/sanitizer_racecheck_new.log 2>&1; el "racecheck rc=$? $(tail -n 2 gpurun_out/sanitizer_racecheck_new.log | tr '\n' ' ')"
cat gpurun_out/progress4.log; cat gpurun_out/sweep4.txt
