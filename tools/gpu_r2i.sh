#!/bin/bash
# Round 2, GPU call I (2 GPUs): why is the scan ~10 % slower per clock inside multi-rank runs than on a 1-GPU lease?
#  a) one process on GPU 0, GPU 1 idle   b) two independent processes, one per GPU, no NCCL   c) the 2-rank bench
mkdir -p gpurun_out
CUDA_VISIBLE_DEVICES=0 timeout 300 python tools/gpu_prof.py --shapes n8shard,b1024 --iters 32 --preheat 1.5 --out gpurun_out/r2i_a_gpu0_alone.json > gpurun_out/r2i_a.log 2>&1
CUDA_VISIBLE_DEVICES=0 timeout 300 python tools/gpu_prof.py --shapes n8shard,b1024 --iters 32 --preheat 1.5 --out gpurun_out/r2i_b_gpu0.json > gpurun_out/r2i_b0.log 2>&1 &
CUDA_VISIBLE_DEVICES=1 timeout 300 python tools/gpu_prof.py --shapes n8shard,b1024 --iters 32 --preheat 1.5 --out gpurun_out/r2i_b_gpu1.json > gpurun_out/r2i_b1.log 2>&1 &
wait
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --steps 20 --warmup 3 --no-extra --no-cpu --no-pipeline > gpurun_out/r2i_c_bench_n2.json 2> gpurun_out/r2i_c.err
NCCL_P2P_DISABLE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus 2 --steps 20 --warmup 3 --no-extra --no-cpu --no-pipeline > gpurun_out/r2i_d_bench_n2_nop2p.json 2> gpurun_out/r2i_d.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2i_[ab]_*.json')):
    for r in json.load(open(f)):
        print(f.split('r2i_')[1][:-5], r['shape'], 'scan_ms %.4f' % r['scan_ms'], 'TF %.0f' % r['tflops'], 'tile %.0f' % r['cycles_per_tile'], 'mma_wait_full %.0f' % r['mma_wait_full'], 'epi_busy %.0f' % r['epi_busy'])
for f in ('gpurun_out/r2i_c_bench_n2.json','gpurun_out/r2i_d_bench_n2_nop2p.json'):
    try:
        d=json.loads([l for l in open(f) if l.startswith('{')][-1])
        print(f, round(d['value']), 'scan TF %.0f' % d['roofline']['achieved'], 'launch_ms %.3f' % d['roofline']['launch_ms'], d['clocks']['sm_mhz'], d['roofline']['same_box'])
    except Exception as e:
        print(f, 'ERR', e)
PY
