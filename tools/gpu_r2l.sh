#!/bin/bash
# Round 2, GPU call L (1 GPU): re-tune the drift control after the epilogue rewrite (pace_gain x max_drift), headline shapes.
mkdir -p gpurun_out
for md in 0 1 2; do
  timeout 400 python tools/gpu_sweep.py --opt pace_gain=0,8,16,32,64 --fixed max_drift=$md --shapes b1024,n8shard --rounds 2 --iters 24 --out gpurun_out/r2l_pace_md$md.json > gpurun_out/r2l_pace_md$md.log 2>&1
  echo "max_drift=$md"; cut -c1-200 gpurun_out/r2l_pace_md$md.log
done
