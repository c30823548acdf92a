#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/small_kernels_r2ap.csv python scripts/prof_small.py > gpurun_out/small_kernels_r2ap.log 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/small_kernels_r2ap.csv')) if len(r)>10 and r[0].isdigit()]
last=None
for r in rows:
    name=r[4].replace('void fgc::','').split('(')[0][:52]
    t=float(r[14].replace(',',''))/1000
    if t>20 and (name,r[8])!=last: print("%-54s %-14s %-12s %8.1f us"%(name,r[8],r[7],t)); last=(name,r[8])
PY
