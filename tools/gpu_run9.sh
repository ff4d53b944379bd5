#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/e2e.txt
for mb in 4 16 64 256; do
  IMPULSE_FFT_STAGE_MB=$mb timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$mb', d['e2e'])" >> gpurun_out/e2e.txt
done
python - <<'PY' >> gpurun_out/e2e.txt
import torch, time
x = torch.empty(1<<30, dtype=torch.uint8, pin_memory=True); d = torch.empty(1<<30, dtype=torch.uint8, device='cuda'); d2=torch.empty_like(d); y=torch.empty(1<<30, dtype=torch.uint8, pin_memory=True)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
for name in ('h2d','d2h','both'):
    torch.cuda.synchronize(); t=time.perf_counter()
    for _ in range(3):
        if name in ('h2d','both'):
            with torch.cuda.stream(s1): d.copy_(x, non_blocking=True)
        if name in ('d2h','both'):
            with torch.cuda.stream(s2): y.copy_(d2, non_blocking=True)
    torch.cuda.synchronize(); dt=(time.perf_counter()-t)/3
    print(name, 'GiB-step ms', round(dt*1e3,2), 'GB/s per dir', round((1<<30)/dt/1e9,1))
PY
cat gpurun_out/e2e.txt
