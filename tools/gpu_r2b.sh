#!/bin/bash
# Round 2, second GPU call: L2-prefetch sweep, the rewritten bench (default command, N=1), ncu captures of the D=768 shape.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "search_matches_oracle or crowd or fallback" > gpurun_out/r2b_pytest.log 2>&1; tail -3 gpurun_out/r2b_pytest.log
timeout 900 python tools/gpu_sweep.py --opt l2_prefetch=0,4,8,16,24,48 --shapes cfg5,b128,b1024,cfg2,cfg4 --out gpurun_out/r2b_sweep_l2pf.json > gpurun_out/r2b_sweep_l2pf.log 2>&1
cut -c1-260 gpurun_out/r2b_sweep_l2pf.log
# the driver's command, as is
( time timeout 1200 python bench.py --steps 20 --warmup 3 ) > gpurun_out/r2b_bench_default.json 2> gpurun_out/r2b_bench_default.err
tail -5 gpurun_out/r2b_bench_default.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r2b_bench_default.json'))
    print({k:d.get(k) for k in ('value','ms_per_step','wall_s')}, d['e2e']['value'], d['clocks'])
    print('roof', {k:d['roofline'].get(k) for k in ('frac','achieved','launch_ms','scan_share_of_step','same_box')})
    print('recall', d.get('recall'))
    print('cpu', {k:d['cpu_baseline'].get(k) for k in ('value','cores','spread')} if d.get('cpu_baseline') else None, d.get('post_check_error'))
    for k,v in (d.get('extra_configs') or {}).items():
        if 'error' in v: print(k, v); continue
        print(k, round(v['value']), round(v['e2e']['value']), v['roofline']['bound'], round(v['roofline']['frac'],3), v['clocks'].get('sm_mhz'), v.get('recall'), v.get('streaming'))
except Exception as e:
    print('ERR', e)
PY
( time timeout 600 python bench.py --impl reference --steps 20 --warmup 3 ) > gpurun_out/r2b_bench_reference.json 2> gpurun_out/r2b_bench_reference.err
cut -c1-900 gpurun_out/r2b_bench_reference.json; tail -3 gpurun_out/r2b_bench_reference.err
# ncu: launch list + full capture of the config-5 shard shape (D = 768, B = 128, k = 5) -- VERDICT r01 item 3
CMD5="python bench.py --rows 6250000 --dim 768 --batch 128 --k 5 --seed 5678 --qseed 8765 --no-cpu --no-extra --data philox --steps 2 --warmup 1 --min-timed-s 0.02 --preheat-max 0.3"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:sa_ -c 60 --csv --log-file gpurun_out/r02_launches_cfg5shard.csv $CMD5 > gpurun_out/r2b_ncu_l5.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sa_scan -s 6 -c 1 -f -o gpurun_out/r02_scan_cfg5shard_cg1 $CMD5 > gpurun_out/r2b_ncu_f5.log 2>&1
tail -2 gpurun_out/r2b_ncu_f5.log
ls -la gpurun_out | grep r02_
