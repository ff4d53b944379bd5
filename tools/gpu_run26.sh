#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/scale_fft2.txt
for n in 4 8; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2961$n bench.py --gpus $n --steps 30 --warmup 5 --workload fft2_8192x8192_c128 2>/dev/null | grep '^{' | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('fft2 p2p n=$n', d['value'], d['ms_per_step'])" >> gpurun_out/scale_fft2.txt
  IMPULSE_FFT_SLAB=nccl timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2962$n bench.py --gpus $n --steps 30 --warmup 5 --workload fft2_8192x8192_c128 2>/dev/null | grep '^{' | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('fft2 nccl n=$n', d['value'], d['ms_per_step'])" >> gpurun_out/scale_fft2.txt
done
timeout 300 python -m pytest tests/test_gpu_dist.py -m gpu -q > gpurun_out/tests_dist8.log 2>&1; tail -2 gpurun_out/tests_dist8.log >> gpurun_out/scale_fft2.txt
cat gpurun_out/scale_fft2.txt
