#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -q -x > gpurun_out/tests_dist.log 2>&1; echo "rc=$?" >> gpurun_out/tests_dist.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 --workload fft2_8192x8192_c128 > gpurun_out/bench_fft2_p2p.json 2> gpurun_out/bench_fft2_p2p.err
IMPULSE_FFT_SLAB=nccl timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 3 --workload fft2_8192x8192_c128 > gpurun_out/bench_fft2_nccl.json 2> gpurun_out/bench_fft2_nccl.err
tail -25 gpurun_out/tests_dist.log | cut -c1-200; for f in bench_fft2_p2p bench_fft2_nccl; do echo == $f; grep '^{' gpurun_out/$f.json | cut -c1-330; tail -2 gpurun_out/$f.err | cut -c1-300; done
