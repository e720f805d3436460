#!/bin/bash
# Round 2, GPU call D (2 GPUs): multi-GPU tests through the C ABI, then the driver's N=2 bench command.
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2d_topo.txt 2>&1
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -x -q > gpurun_out/r2d_pytest.log 2>&1; tail -3 gpurun_out/r2d_pytest.log
( time timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 3 ) > gpurun_out/r2d_bench_n2.json 2> gpurun_out/r2d_bench_n2.err
tail -8 gpurun_out/r2d_bench_n2.err | cut -c1-400
python - <<'PY'
import json
try:
    d=json.loads([l for l in open('gpurun_out/r2d_bench_n2.json') if l.startswith('{')][-1])
    print({k:d.get(k) for k in ('value','ms_per_step','wall_s','n_gpus')}, d['e2e']['value'], d['clocks'])
    print('cfg', d['config'])
    print('roof', {k:d['roofline'].get(k) for k in ('frac','achieved','launch_ms','scan_share_of_step','same_box','search_ms_events')})
    print('recall', d.get('recall'))
    print('pipeline', d.get('e2e_pipeline'))
    for k,v in (d.get('extra_configs') or {}).items():
        if 'error' in v: print(k, v); continue
        print(k, round(v['value']), round(v['e2e']['value']), v['roofline']['bound'], round(v['roofline']['frac'],3), round(v['roofline']['scan_share_of_step'],3), v['clocks'].get('sm_mhz'), v.get('recall'), v.get('streaming'))
except Exception as e:
    print('ERR', e)
PY
