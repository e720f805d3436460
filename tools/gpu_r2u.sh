#!/bin/bash
# Round 2, call U (host CPU of the GPU box only): native decoder threads and the host search stage on the real host.
mkdir -p gpurun_out
( nproc
for t in 1 2 4 8; do SA_WIRE_THREADS=$t python tools/host_decode_bench.py; done
for t in 1 4; do SA_WIRE_THREADS=$t python tools/host_pipeline_bench.py 32768 | tail -1; done ) > gpurun_out/r2u_host.txt 2>&1
cat gpurun_out/r2u_host.txt
