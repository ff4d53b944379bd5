#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
python tools/run_slab_dist.py 8192 2 | tee gpurun_out/r02_slab_dist_n2.txt
timeout 300 ncu --metrics gpu__time_duration.sum,nvlrx__bytes.sum,nvltx__bytes.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_srcunit_tex_aperture_peer.sum,lts__t_sectors_srcunit_tex_aperture_peer_op_read.sum --clock-control none -k regex:'colfast2|colpipe2|colfuse2' -s 12 -c 6 --csv --log-file gpurun_out/r02_cols_from_parts_nvlink.csv python tools/run_slab_dist.py 8192 2 > /dev/null 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r02_cols_from_parts_nvlink.csv')))
h=[i for i,r in enumerate(rows) if 'Kernel Name' in r]
if h:
    H=rows[h[0]]
    for r in rows[h[0]+1:]:
        if len(r)==len(H): print(r[H.index('ID')], r[H.index('Kernel Name')][:50], r[H.index('Metric Name')], r[H.index('Metric Value')], r[H.index('Metric Unit')])
else:
    print(open('gpurun_out/r02_cols_from_parts_nvlink.csv').read()[-1500:])
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/pcie_aggregate.py 2>/dev/null | tee gpurun_out/r02_pcie_aggregate_n2.txt
python tools/pcie_aggregate.py | tee gpurun_out/r02_pcie_aggregate_n1.txt
