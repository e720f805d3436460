#!/bin/bash
mkdir -p gpurun_out
ncu --metrics dram__bytes_read.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum --clock-control none -k regex:sa_scan --csv --log-file gpurun_out/l2exp.csv python tools/gpu_l2exp.py 4000000 1024 > gpurun_out/l2exp.log 2>&1
cat gpurun_out/l2exp.log | grep launch
python - <<'PY'
import csv
rows=list(csv.DictReader(open('gpurun_out/l2exp.csv')))
by={}
for r in rows:
    by.setdefault(r['ID'],{})[r['Metric Name']]=r['Metric Value']
for i,m in by.items(): print(i, m)
PY
