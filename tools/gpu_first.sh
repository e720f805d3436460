#!/bin/bash
# First hardware bring-up: staged, each under its own timeout; logs to gpurun_out/first.log
mkdir -p gpurun_out
L=gpurun_out/first.log
: > $L
run() { echo "=== $*" >> $L; timeout 300 python tests/harness/gpu_stage.py "$@" >> $L 2>&1; echo "rc=$?" >> $L; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv >> $L 2>&1
run dots 1 1024 1536 128
run dots 1 700 768 100
run search 1 20000 1536 200 10
run search 1 5000 768 37 5
run dots 2 1024 1536 256
run search 2 20000 1536 300 10
run perf 1 1000000 1536 128
run perf 1 1000000 1536 256
run perf 2 1000000 1536 256
tail -60 $L
