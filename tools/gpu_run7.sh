#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt
timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -q -x > gpurun_out/tests_dist.log 2>&1; echo "rc=$?" >> gpurun_out/tests_dist.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --workload fft2_8192x8192_c128 > gpurun_out/bench_fft2_n2.json 2> gpurun_out/bench_fft2_n2.err
timeout 300 python bench.py --steps 10 --warmup 3 --workload fft2_8192x8192_c128 --no-e2e --no-cpu > gpurun_out/bench_fft2_n1.json 2> gpurun_out/bench_fft2_n1.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.err
tail -5 gpurun_out/tests_dist.log; for f in bench_n2 bench_fft2_n2 bench_fft2_n1 bench_ref_n2; do echo == $f; cut -c1-700 gpurun_out/$f.json; tail -2 gpurun_out/$f.err; done
