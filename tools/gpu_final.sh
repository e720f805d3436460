#!/bin/bash
# Final round-1 evidence at HEAD: tests, smoke, benches, ncu launch lists + full captures
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -x -q -m gpu ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
timeout 600 python bench.py --steps 10 --warmup 3 --batch 128 --no-cpu > gpurun_out/bench_b128.json 2> gpurun_out/bench_b128.err
timeout 600 python bench.py --steps 20 --warmup 3 --rows 1000000 --batch 256 --no-cpu > gpurun_out/bench_1m_b256.json 2> gpurun_out/bench_1m_b256.err
timeout 600 python bench.py --steps 20 --warmup 3 --rows 6250000 --dim 768 --batch 128 --k 5 --no-cpu > gpurun_out/bench_cfg5.json 2> gpurun_out/bench_cfg5.err
bash tools/gpu_ncu.sh > gpurun_out/ncu_final.log 2>&1
timeout 600 python tools/host_pipeline_bench.py 16384 --gpu > gpurun_out/pipeline_gpu.log 2>&1
tail -4 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/smoke.log; cat gpurun_out/bench_default.json gpurun_out/bench_reference.json; tail -2 gpurun_out/pipeline_gpu.log
for f in gpurun_out/bench_b128.json gpurun_out/bench_1m_b256.json gpurun_out/bench_cfg5.json; do python -c "
import json,sys; d=json.load(open('$f')); print('$f', round(d['value']), round(d['e2e']['value']), round(d['roofline']['achieved'],1), d['roofline']['unit'], round(d['roofline']['frac'],4))"; done
