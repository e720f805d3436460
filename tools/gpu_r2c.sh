#!/bin/bash
# Round 2, GPU call C: the driver's bench commands as they are (N = 1), own arm and reference arm.
mkdir -p gpurun_out
( free -g; cat /sys/fs/cgroup/memory.max 2>/dev/null ) > gpurun_out/r2c_host.txt 2>&1
( time timeout 1500 python bench.py --steps 20 --warmup 3 ) > gpurun_out/r2c_bench_default.json 2> gpurun_out/r2c_bench_default.err
tail -5 gpurun_out/r2c_bench_default.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r2c_bench_default.json'))
    print({k:d.get(k) for k in ('value','ms_per_step','wall_s')}, d['e2e']['value'], d['clocks'])
    print('cfg', d['config'])
    print('roof', {k:d['roofline'].get(k) for k in ('frac','achieved','launch_ms','scan_share_of_step','same_box')})
    print('recall', d.get('recall'))
    print('cpu', {k:d['cpu_baseline'].get(k) for k in ('value','cores','spread')} if d.get('cpu_baseline') else None, d.get('post_check_error'))
    for k,v in (d.get('extra_configs') or {}).items():
        if 'error' in v: print(k, v); continue
        print(k, round(v['value']), round(v['e2e']['value']), v['roofline']['bound'], round(v['roofline']['frac'],3), v['clocks'].get('sm_mhz'), v.get('recall'), v.get('streaming'))
except Exception as e:
    print('ERR', e)
PY
( time timeout 600 python bench.py --impl reference --steps 20 --warmup 3 ) > gpurun_out/r2c_bench_reference.json 2> gpurun_out/r2c_bench_reference.err
cut -c1-1200 gpurun_out/r2c_bench_reference.json; tail -3 gpurun_out/r2c_bench_reference.err
