#!/bin/bash
bash tools/gpu_ncu.sh > gpurun_out/ncu_final.log 2>&1; tail -2 gpurun_out/ncu_final.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; cat gpurun_out/bench_default.json
