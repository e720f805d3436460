#!/bin/bash
# Round check on one B200: GPU test-suite, headline bench, HBM-bound bench, variants.  Logs -> gpurun_out/
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
lscpu | head -20 >> gpurun_out/gpu.txt; free -g >> gpurun_out/gpu.txt
( time timeout 1500 python -m pytest tests -x -q -m gpu ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
timeout 600 python bench.py --steps 10 --warmup 3 --batch 128 --no-cpu > gpurun_out/bench_b128.json 2> gpurun_out/bench_b128.err
timeout 600 python bench.py --steps 10 --warmup 3 --cta-group 1 --no-cpu > gpurun_out/bench_cg1.json 2> gpurun_out/bench_cg1.err
timeout 600 python bench.py --steps 10 --warmup 3 --rows 1000000 --batch 256 --no-cpu > gpurun_out/bench_1m_b256.json 2> gpurun_out/bench_1m_b256.err
tail -5 gpurun_out/pytest_gpu.log; cat gpurun_out/bench_*.json; tail -3 gpurun_out/bench_*.err
