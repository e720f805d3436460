#!/bin/bash
# Round 2, GPU call H (1 GPU): sanitizer + parity on the per-warp fallback lists and the warp shard-merge, then the driver's
# own command lines once more at HEAD (evidence for profiles/).
mkdir -p gpurun_out; rm -f gpurun_out/sanitize_summary.txt
timeout 900 python -m pytest tests -m gpu -x -q -k "crowd or fallback or merge_shards or multi or golden or serve" > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2h_pytest.log; tail -4 gpurun_out/r2h_pytest.log
bash tools/gpu_sanitize.sh > gpurun_out/r2h_sanitize.log 2>&1; cat gpurun_out/sanitize_summary.txt
grep -c "fix_list_insert" gpurun_out/sanitize_racecheck_fixup.log
( time timeout 1500 python bench.py --steps 20 --warmup 3 ) > gpurun_out/r2h_bench_default.json 2> gpurun_out/r2h_bench_default.err
tail -4 gpurun_out/r2h_bench_default.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r2h_bench_default.json'))
    print({k:d.get(k) for k in ('value','ms_per_step','wall_s')}, d['e2e']['value'], d['clocks'])
    print('roof', {k:d['roofline'].get(k) for k in ('frac','achieved','launch_ms','scan_share_of_step','same_box')})
    print('recall', d.get('recall')); print('pipeline', d.get('e2e_pipeline'))
    print('cpu', {k:d['cpu_baseline'].get(k) for k in ('value','cores','spread')} if d.get('cpu_baseline') else None, d.get('post_check_error'))
    for k,v in (d.get('extra_configs') or {}).items():
        if 'error' in v: print(k, v); continue
        print(k, round(v['value']), round(v['e2e']['value']), v['roofline']['bound'], round(v['roofline']['frac'],3), v['clocks'].get('sm_mhz'), (v.get('recall') or {}).get('strict_order'), (v.get('streaming') or {}).get('value'))
except Exception as e:
    print('ERR', e)
PY
