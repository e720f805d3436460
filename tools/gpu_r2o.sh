#!/bin/bash
# Round 2, GPU call O (1 GPU): full GPU suite with the window bound on by default, the adversarial orders with the
# window bound off / on (near-duplicate queries, the case that cost +7..12 %), and the N = 1 bench line at HEAD.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2o_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2o_pytest.log; tail -3 gpurun_out/r2o_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2o_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2o_smoke.log; tail -2 gpurun_out/r2o_smoke.log | cut -c1-200
timeout 600 python tools/gpu_worstcase.py --rows 4000000 --batches 128,1024 --option window_bound --presample 0,1 --noise 0.1 --out gpurun_out/r2o_worstcase_window.json > gpurun_out/r2o_worstcase_window.log 2>&1
cut -c1-240 gpurun_out/r2o_worstcase_window.log
timeout 900 python bench.py > gpurun_out/r2o_bench_n1.json 2> gpurun_out/r2o_bench_n1.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/r2o_bench_n1.json
