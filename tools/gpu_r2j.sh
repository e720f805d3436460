#!/bin/bash
# Round 2, GPU call J (8 GPUs): NVLink P2P vs host-shared-memory transport of the one all-gather, same box, A/B/A order
# (headline only), then the driver's full N=8 command at HEAD.
mkdir -p gpurun_out
Q="--gpus 8 --steps 20 --warmup 3 --no-extra --no-cpu --no-pipeline"
T="timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
$T --master-port 29531 bench.py $Q > gpurun_out/r2j_a1_nvlink.json 2> gpurun_out/r2j_a1.err
NCCL_P2P_DISABLE=1 $T --master-port 29532 bench.py $Q > gpurun_out/r2j_b_shm.json 2> gpurun_out/r2j_b.err
$T --master-port 29533 bench.py $Q > gpurun_out/r2j_a2_nvlink.json 2> gpurun_out/r2j_a2.err
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 8 --steps 20 --warmup 3 ) > gpurun_out/r2j_bench_n8.json 2> gpurun_out/r2j_bench_n8.err
tail -4 gpurun_out/r2j_bench_n8.err | cut -c1-300
python - <<'PY'
import json
for f in ('gpurun_out/r2j_a1_nvlink.json','gpurun_out/r2j_b_shm.json','gpurun_out/r2j_a2_nvlink.json','gpurun_out/r2j_bench_n8.json'):
    try:
        d=json.loads([l for l in open(f) if l.startswith('{')][-1])
        print(f, {k:d.get(k) for k in ('value','wall_s')}, 'e2e', round(d['e2e']['value']), 'mhz', d['clocks'].get('sm_mhz'), 'W', d['clocks'].get('power_w_median'), 'region', d['config']['timed_region_s'], 'bps', d['config']['batches_per_step'])
        print('  roof', {k:d['roofline'].get(k) for k in ('frac','achieved','launch_ms','scan_share_of_step','same_box')})
        if d.get('recall'): print('  recall', d.get('recall'))
        if d.get('e2e_pipeline'): print('  pipeline', (d.get('e2e_pipeline') or {}).get('value'))
        for k,v in (d.get('extra_configs') or {}).items():
            if 'error' in v: print(k, v); continue
            print('  ', k, round(v['value']), round(v['e2e']['value']), v['roofline']['bound'], round(v['roofline']['frac'],3), round(v['roofline']['scan_share_of_step'],3), v['clocks'].get('sm_mhz'), (v.get('recall') or {}).get('strict_order'), (v.get('streaming') or {}).get('value'))
    except Exception as e:
        print(f, 'ERR', e)
PY
