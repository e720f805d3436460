#!/bin/bash
# N-GPU validation: row-sharded bench through torchrun (NCCL all-gather + merge)
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_$N.txt 2>&1
P='import json,sys; d=json.loads(sys.stdin.read()); print(sys.argv[1], "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), d["roofline"]["unit"], round(d["roofline"]["achieved"],1), "frac", round(d["roofline"]["frac"],4), "ms", round(d["ms_per_step"],3), d.get("recall"))'
run() { name=$1; shift; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@" > gpurun_out/bench_n${N}_$name.json 2> gpurun_out/bench_n${N}_$name.err; tail -1 gpurun_out/bench_n${N}_$name.json | python -c "$P" "N=$N $name"; }
run default --steps 10 --warmup 3
run b4096 --steps 10 --warmup 3 --batch 4096 --no-cpu
run cfg5 --steps 20 --warmup 3 --rows $((6250000*N)) --dim 768 --batch 128 --k 5 --no-cpu
