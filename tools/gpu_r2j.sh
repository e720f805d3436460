#!/bin/bash
# Round 2, GPU call J (8 GPUs): (1) headline only with NCCL's NVLink P2P transport disabled (exchange through host shared
# memory: is the NVLink power worth more than its latency at 8 ranks?), (2) the driver's full N=8 command at HEAD.
mkdir -p gpurun_out
NCCL_P2P_DISABLE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 3 --no-extra --no-cpu --no-pipeline > gpurun_out/r2j_bench_n8_nop2p.json 2> gpurun_out/r2j_nop2p.err
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 20 --warmup 3 ) > gpurun_out/r2j_bench_n8.json 2> gpurun_out/r2j_bench_n8.err
tail -4 gpurun_out/r2j_bench_n8.err | cut -c1-300
python - <<'PY'
import json
for f in ('gpurun_out/r2j_bench_n8_nop2p.json','gpurun_out/r2j_bench_n8.json'):
    try:
        d=json.loads([l for l in open(f) if l.startswith('{')][-1])
        print(f, {k:d.get(k) for k in ('value','ms_per_step','wall_s')}, 'e2e', round(d['e2e']['value']), d['clocks'].get('sm_mhz'), d['clocks'].get('power_w_median'))
        print('  roof', {k:d['roofline'].get(k) for k in ('frac','achieved','launch_ms','scan_share_of_step','same_box')})
        print('  recall', d.get('recall')); print('  pipeline', (d.get('e2e_pipeline') or {}).get('value'))
        for k,v in (d.get('extra_configs') or {}).items():
            if 'error' in v: print(k, v); continue
            print('  ', k, round(v['value']), round(v['e2e']['value']), v['roofline']['bound'], round(v['roofline']['frac'],3), round(v['roofline']['scan_share_of_step'],3), v['clocks'].get('sm_mhz'), (v.get('recall') or {}).get('strict_order'), (v.get('streaming') or {}).get('value'))
    except Exception as e:
        print(f, 'ERR', e)
PY
