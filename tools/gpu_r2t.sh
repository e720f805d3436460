#!/bin/bash
# Round 2, GPU call T (4 GPUs): the one rank count never exercised -- a short N = 4 bench (headline only).
mkdir -p gpurun_out
timeout 330 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 4 --steps 3 --warmup 1 --no-extra --no-pipeline --min-timed-s 0.5 --recall-queries 32 --cpu-sample-queries 32 --cpu-sample-rows 131072 > gpurun_out/r2t_bench_n4.json 2> gpurun_out/r2t_bench_n4.err; echo "bench rc=$?"
tail -2 gpurun_out/r2t_bench_n4.err | cut -c1-200; cut -c1-400 gpurun_out/r2t_bench_n4.json
