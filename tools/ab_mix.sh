#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
M=smsp__thread_inst_executed.sum,sm__sass_thread_inst_executed_op_fp32_pred_on.sum,sm__sass_thread_inst_executed_op_integer_pred_on.sum,sm__sass_thread_inst_executed_op_control_pred_on.sum,sm__sass_thread_inst_executed_op_memory_pred_on.sum,sm__sass_thread_inst_executed_op_conversion_pred_on.sum,sm__sass_thread_inst_executed_op_misc_pred_on.sum,sm__sass_thread_inst_executed_op_uniform_pred_on.sum,sm__sass_thread_inst_executed_op_bit_pred_on.sum,sm__sass_thread_inst_executed_op_fp64_pred_on.sum,sm__sass_thread_inst_executed_op_inter_thread_communication_pred_on.sum
timeout 300 ncu --metrics $M --clock-control none -s 15 -c 5 --csv --log-file gpurun_out/filter_mix.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-configs --workload filter2d_64x4096x4096_f32 > /dev/null 2>&1
python - <<'PY'
import csv,collections
rows=list(csv.reader(open('gpurun_out/filter_mix.csv')))
h=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]
H=rows[h]
d=collections.OrderedDict()
for r in rows[h+1:]:
    if len(r)==len(H): d.setdefault((r[H.index('ID')],r[H.index('Kernel Name')][:60]),{})[r[H.index('Metric Name')].replace('sm__sass_thread_inst_executed_op_','').replace('_pred_on.sum','')]=float(r[H.index('Metric Value')].replace(',',''))
for k,v in d.items():
    tot=v.get('smsp__thread_inst_executed.sum',1)
    print(k[1], f"total={tot:.3g}", ' '.join(f"{a}={100*b/tot:.1f}%" for a,b in v.items() if a!='smsp__thread_inst_executed.sum' and b>0))
PY
