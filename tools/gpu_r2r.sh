#!/bin/bash
# Round 2, GPU call R (1 GPU): the ncu launch list of the bench command itself (headline only, short timed regions).
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 1 --no-cpu --no-extra --no-pipeline --min-timed-s 0.3 --preheat-max 0.5"
timeout 420 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:sa_ -c 400 --csv --log-file gpurun_out/r02b_launches_bench_n1.csv $CMD > gpurun_out/r2r_bench_under_ncu.json 2> gpurun_out/r2r_bench_under_ncu.err
echo "rc=$?"; tail -2 gpurun_out/r2r_bench_under_ncu.err | cut -c1-200; wc -l gpurun_out/r02b_launches_bench_n1.csv
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/r02b_launches_bench_n1.csv')) if len(r) > 10 and r[0].isdigit()]
t = collections.Counter(); c = collections.Counter()
for r in rows:
    name = r[4].split('(')[0][:60]; t[name] += float(r[-1]); c[name] += 1
tot = sum(t.values())
for k, v in t.most_common(): print(f"{k:60s} n={c[k]:4d}  total {v/1e6:9.3f} ms  share {v/tot:.4f}  avg {v/c[k]/1e3:9.1f} us")
PY
