#!/bin/bash
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -x -q -m gpu ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
timeout 300 compute-sanitizer --tool racecheck --error-exitcode 9 python tests/harness/gpu_stage.py search 1 3000 256 100 10 > gpurun_out/sanitize_racecheck_cg1.log 2>&1; echo "racecheck cg1 rc=$?"; tail -2 gpurun_out/sanitize_racecheck_cg1.log
timeout 600 python tools/host_pipeline_bench.py 16384 --gpu 2>&1 | tail -2
timeout 600 python tests/harness/stream_bench.py 2>&1 | tail -2
timeout 600 python tests/harness/stream_bench.py --rows 2000000 --dim 1536 --batch 256 --k 10 --epoch-rows 100000 --steps 60 2>&1 | tail -2
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu 2>&1 | tail -1
