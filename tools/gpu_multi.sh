#!/bin/bash
# N-GPU validation: row-sharded bench through torchrun (NCCL all-gather + merge)
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_$N.txt 2>&1
export NCCL_DEBUG=WARN
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "rc=$?"; cat gpurun_out/bench_n$N.json; grep -v "OMP_NUM\|\*\*\*\*" gpurun_out/bench_n$N.err | tail -5
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 --batch 4096 --no-cpu > gpurun_out/bench_n${N}_b4096.json 2> gpurun_out/bench_n${N}_b4096.err
cat gpurun_out/bench_n${N}_b4096.json; grep -v "OMP_NUM\|\*\*\*\*" gpurun_out/bench_n${N}_b4096.err | tail -3
# config 5 shape: 50M x 768 over N GPUs, batch 128, top-5
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 20 --warmup 3 --rows $((6250000*N)) --dim 768 --batch 128 --k 5 --no-cpu > gpurun_out/bench_n${N}_cfg5.json 2> gpurun_out/bench_n${N}_cfg5.err
cat gpurun_out/bench_n${N}_cfg5.json; grep -v "OMP_NUM\|\*\*\*\*" gpurun_out/bench_n${N}_cfg5.err | tail -3
