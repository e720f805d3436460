#!/bin/bash
# Round 2, GPU call P (1 GPU): ncu captures of the scan kernel at HEAD (window bound in, 227 registers): launch lists and
# --set full for the headline shape, the 1M / B=256 shape and the D = 768 shard shape.
mkdir -p gpurun_out
for shape in b1024 cfg2 cfg5; do
  CMD="python tools/gpu_prof.py --shapes $shape --iters 3 --preheat 0.2"
  timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:sa_ -c 40 --csv --log-file gpurun_out/r02b_launches_$shape.csv $CMD > gpurun_out/r2p_l_$shape.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:sa_scan -s 5 -c 1 -f -o gpurun_out/r02b_scan_$shape $CMD > gpurun_out/r2p_f_$shape.log 2>&1
  tail -1 gpurun_out/r2p_f_$shape.log | cut -c1-200
done
ls -la gpurun_out | grep r02b_
