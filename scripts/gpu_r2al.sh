#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
REPS=1 ONLY=chan_stats,cbn_act_fwd,cbn_act_bwd,minmax_fwd,minmax_bwd timeout -k 10 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/elem_kernels_r2al.csv python scripts/prof_elem.py > gpurun_out/elem_kernels_r2al.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/elem_kernels_r2al.csv')) if len(r)>10 and r[0].isdigit()]
d={}
for r in rows:
    d.setdefault(r[0],{'name':r[4].replace('void fgc::','').split('(')[0][:48],'grid':r[8]})[r[12]]=float(r[14].replace(',',''))
for k,v in d.items():
    t=v.get('gpu__time_duration.sum',0)/1000
    b=(v.get('dram__bytes_read.sum',0)+v.get('dram__bytes_write.sum',0))
    if t>8: print("%-50s %-14s %8.1f us  %7.0f MB  %.2f TB/s"%(v['name'],v['grid'],t,b/1e6,b/t/1e6))
PY
