#!/bin/bash
# N-GPU validation: row-sharded bench through torchrun (NCCL all-gather + merge)
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_$N.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "rc=$?"; cat gpurun_out/bench_n$N.json; tail -5 gpurun_out/bench_n$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 --batch 4096 > gpurun_out/bench_n${N}_b4096.json 2> gpurun_out/bench_n${N}_b4096.err
cat gpurun_out/bench_n${N}_b4096.json; tail -3 gpurun_out/bench_n${N}_b4096.err
