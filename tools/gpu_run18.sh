#!/bin/bash
# 8-GPU scaling: batch shard (weak) and slab fft2 (strong)
mkdir -p gpurun_out; rm -f gpurun_out/scale.txt
for n in 1 2 4 8; do
  if [ $n -eq 1 ]; then
    timeout 200 python bench.py --gpus 1 --steps 50 --warmup 5 --no-cpu --no-e2e 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c2c n=$n', d['value'], d['ms_per_step'])" >> gpurun_out/scale.txt
    timeout 200 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu --no-e2e --workload fft2_8192x8192_c128 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('fft2 n=$n', d['value'], d['ms_per_step'])" >> gpurun_out/scale.txt
  else
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 50 --warmup 5 --no-e2e 2>/dev/null | grep '^{' | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c2c n=$n', d['value'], d['ms_per_step'])" >> gpurun_out/scale.txt
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n bench.py --gpus $n --steps 10 --warmup 3 --workload fft2_8192x8192_c128 2>/dev/null | grep '^{' | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('fft2 n=$n', d['value'], d['ms_per_step'])" >> gpurun_out/scale.txt
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n bench.py --gpus $n --steps 5 --warmup 3 --workload filter2d_64x4096x4096_f32 2>/dev/null | grep '^{' | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('filter n=$n', d['value'], d['ms_per_step'])" >> gpurun_out/scale.txt
  fi
done
timeout 300 python -m pytest tests/test_gpu_dist.py -m gpu -q > gpurun_out/tests_dist8.log 2>&1; tail -2 gpurun_out/tests_dist8.log >> gpurun_out/scale.txt
cat gpurun_out/scale.txt
