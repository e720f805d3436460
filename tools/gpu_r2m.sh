#!/bin/bash
# Round 2, GPU call M (1 GPU): regression at HEAD after the drift-control re-tune, the "two launches of 2 x 37 lanes"
# (grid 148) alternative measured rather than costed, and the N = 1 bench line.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2m_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2m_pytest.log; tail -3 gpurun_out/r2m_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2m_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2m_smoke.log; tail -2 gpurun_out/r2m_smoke.log | cut -c1-200
timeout 400 python tools/gpu_sweep.py --opt max_launch_qblocks=0,2 --shapes b1024 --rounds 3 --iters 24 --out gpurun_out/r2m_grid148.json > gpurun_out/r2m_grid148.log 2>&1
cut -c1-220 gpurun_out/r2m_grid148.log
timeout 900 python bench.py > gpurun_out/r2m_bench_n1.json 2> gpurun_out/r2m_bench_n1.err; echo "bench rc=$?"; cut -c1-1500 gpurun_out/r2m_bench_n1.json
