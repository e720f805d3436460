#!/bin/bash
# ncu evidence for the scan kernel (one GPU; never under torchrun).  Outputs -> gpurun_out/
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 1 --no-cpu"
# launch lists of the bench command (cold-cache, serialised: compare shares)
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:sa_ -c 40 --csv --log-file gpurun_out/launches_b1024.csv $B > gpurun_out/ncu_l1.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:sa_ -c 40 --csv --log-file gpurun_out/launches_b128.csv $B --batch 128 > gpurun_out/ncu_l2.log 2>&1
# full captures of the dominant kernel
ncu --set full --clock-control none --import-source on -k regex:sa_scan -s 3 -c 1 -f -o gpurun_out/scan_b1024_cg2 $B > gpurun_out/ncu_f1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sa_scan -s 3 -c 1 -f -o gpurun_out/scan_b128_cg1 $B --batch 128 > gpurun_out/ncu_f2.log 2>&1
ls -la gpurun_out/*.ncu-rep; tail -3 gpurun_out/ncu_f1.log
