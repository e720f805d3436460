#!/bin/bash
# usage: gpu_l2exp.sh ROWS BATCH "settings"
mkdir -p gpurun_out
ncu --metrics dram__bytes_read.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum --clock-control none -k regex:sa_scan --csv --log-file gpurun_out/l2exp.csv python tools/gpu_l2exp.py "$1" "$2" "$3" > gpurun_out/l2exp.log 2>&1
python - <<'PY'
import csv
lab=[l.strip() for l in open('gpurun_out/l2exp.log') if l.startswith('launch')]
lines=[l for l in open('gpurun_out/l2exp.csv') if l.startswith('"')]
by={}
for r in csv.DictReader(lines):
    by.setdefault(int(r['ID']),{})[r['Metric Name']]=float(r['Metric Value'].replace(',',''))
for i,m in sorted(by.items()):
    print(lab[i] if i<len(lab) else i, f"dram {m['dram__bytes_read.sum']/1e9:.2f} GB  hit {m['lts__t_sector_hit_rate.pct']:.1f}%  time {m['gpu__time_duration.sum']/1e6:.3f} ms")
PY
