#!/bin/bash
# Round 2, GPU call K (1 GPU): last regression run of the whole GPU suite at HEAD + the mbarrier wait-hint sweep.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2k_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2k_pytest.log; tail -4 gpurun_out/r2k_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2k_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2k_smoke.log; tail -2 gpurun_out/r2k_smoke.log | cut -c1-300
timeout 600 python tools/gpu_sweep.py --opt wait_hint_ns=0,1000,20000 --shapes b1024,n8shard,cfg5 --rounds 3 --out gpurun_out/r2k_sweep_wait_hint.json > gpurun_out/r2k_sweep_wait_hint.log 2>&1
cut -c1-230 gpurun_out/r2k_sweep_wait_hint.log
# Round 2, GPU call L (1 GPU): re-tune the drift control after the epilogue rewrite (pace_gain x max_drift), headline shapes.
mkdir -p gpurun_out
for md in 0 1 2; do
  timeout 400 python tools/gpu_sweep.py --opt pace_gain=0,8,16,32,64 --fixed max_drift=$md --shapes b1024,n8shard --rounds 2 --iters 24 --out gpurun_out/r2l_pace_md$md.json > gpurun_out/r2l_pace_md$md.log 2>&1
  echo "max_drift=$md"; cut -c1-200 gpurun_out/r2l_pace_md$md.log
done
