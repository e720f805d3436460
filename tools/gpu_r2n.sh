#!/bin/bash
# Round 2, GPU call N (1 GPU): the window bound (kKL/2 lanes' second bests) -- parity suite with it on by default, then
# A/B on time (option window_bound = 0 / 1) for the stock build and for the tournament-argmax build of the insertion
# round (-DSA_ARGMAX_TREE, libsa_b200_tree.so), and the role-cycle counters on the two short-scan shapes.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2n_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2n_pytest.log; tail -3 gpurun_out/r2n_pytest.log
SH=cfg2,n8shard,b1024,b128,cfg5
timeout 500 python tools/gpu_sweep.py --opt window_bound=0,1 --shapes $SH --rounds 3 --iters 24 --out gpurun_out/r2n_window_stock.json > gpurun_out/r2n_window_stock.log 2>&1
echo stock; cut -c1-200 gpurun_out/r2n_window_stock.log
SA_LIB_PATH=$PWD/quickstart-streaming-agents_b200/libsa_b200_tree.so timeout 500 python tools/gpu_sweep.py --opt window_bound=0,1 --shapes $SH --rounds 3 --iters 24 --out gpurun_out/r2n_window_tree.json > gpurun_out/r2n_window_tree.log 2>&1
echo tree; cut -c1-200 gpurun_out/r2n_window_tree.log
timeout 500 python tools/gpu_sweep.py --opt window_bound=0,1 --shapes $SH --rounds 3 --iters 24 --out gpurun_out/r2n_window_stock2.json > gpurun_out/r2n_window_stock2.log 2>&1
echo stock again; cut -c1-200 gpurun_out/r2n_window_stock2.log
for wb in 0 1; do
  timeout 300 python tools/gpu_prof.py --shapes cfg2,n8shard --iters 12 --preheat 0.5 --opts window_bound=$wb --out gpurun_out/r2n_roles_wb$wb.json > gpurun_out/r2n_roles_wb$wb.log 2>&1
  echo "roles window_bound=$wb"; tail -4 gpurun_out/r2n_roles_wb$wb.log | cut -c1-420
done
