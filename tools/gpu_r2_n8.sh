#!/bin/bash
# 8-GPU record: dist C-ABI tests on all devices, bench under torchrun (weak line, configs, strong c2c, slab fft2 + accuracy)
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
nvidia-smi topo -m 2>/dev/null | head -14 > gpurun_out/r02_topo.txt
for d in /sys/bus/pci/devices/*/numa_node; do :; done; python - <<'PY' > gpurun_out/r02_numa.txt 2>&1
import torch, ctypes
import impulse_b200 as ib
for i in range(torch.cuda.device_count()):
    import os
    aff = os.sched_getaffinity(0)
    node = ib.bind_host_to_device(i)
    print(i, torch.cuda.get_device_properties(i).name, 'numa', node, 'cpus', len(os.sched_getaffinity(0)))
    os.sched_setaffinity(0, aff)
PY
cat gpurun_out/r02_numa.txt
timeout 900 python -m pytest tests/test_gpu_dist_cabi.py tests/test_gpu_dist.py -x -q -m gpu 2>&1 | tail -n 4
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02_bench_n8.json 2> gpurun_out/r02_bench_n8.err
echo "bench n8 rc=$?"; tail -c 600 gpurun_out/r02_bench_n8.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_n8.json').read().strip().splitlines()[-1])
print('main', d['value'], d['ms_per_step'], d['scaling'], 'e2e', d.get('e2e',{}).get('value'), d.get('e2e',{}).get('ms_per_step'), 'numa', d['config'].get('host_numa_node'))
for k in ('2_c2c_65536x1024_strong','4_fft2_8192x8192_slab'):
    print(k, json.dumps(d['configs'][k])[:1500])
for k,v in d['configs'].items():
    if 'GB/s' in v: print(k, v['ms_per_step'], v['GB/s'], v.get('accuracy',{}).get('pass'))
print('all pass', d.get('configs_accuracy_all_pass'))
PY
