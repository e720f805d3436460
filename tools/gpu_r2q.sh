#!/bin/bash
# Round 2, GPU call Q (2 GPUs): the multi-GPU tests and the N = 2 bench line at HEAD (window bound in).
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -x -q > gpurun_out/r2q_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2q_pytest.log; tail -3 gpurun_out/r2q_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 3 --no-extra > gpurun_out/r2q_bench_n2.json 2> gpurun_out/r2q_bench_n2.err; echo "bench rc=$?"
tail -3 gpurun_out/r2q_bench_n2.err | cut -c1-200; cut -c1-600 gpurun_out/r2q_bench_n2.json
