#!/bin/bash
# Round 2, GPU call B (no host-memory-hungry step in here): host probe, parity subset, L2-prefetch sweep, ncu of the D=768 shape.
mkdir -p gpurun_out
( nproc; free -g; cat /sys/fs/cgroup/memory.max 2>/dev/null; cat /sys/fs/cgroup/memory.current 2>/dev/null; df -h /dev/shm /tmp | cat; lscpu | grep -E "Model name|Socket|NUMA node\(s\)" ) > gpurun_out/r2b_host.txt 2>&1
cat gpurun_out/r2b_host.txt
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "search_matches_oracle or crowd or fallback" > gpurun_out/r2b_pytest.log 2>&1; tail -3 gpurun_out/r2b_pytest.log
timeout 900 python tools/gpu_sweep.py --opt l2_prefetch=0,4,8,16,24,48 --shapes cfg5,b128,b1024,cfg2,cfg4 --out gpurun_out/r2b_sweep_l2pf.json > gpurun_out/r2b_sweep_l2pf.log 2>&1
cut -c1-200 gpurun_out/r2b_sweep_l2pf.log
# ncu: launch list + full capture of the config-5 shard shape (D = 768, B = 128, k = 5) -- VERDICT r01 item 3
CMD5="python tools/gpu_prof.py --shapes cfg5 --iters 4 --preheat 0.2"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:sa_ -c 60 --csv --log-file gpurun_out/r02_launches_cfg5shard.csv $CMD5 > gpurun_out/r2b_ncu_l5.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sa_scan -s 6 -c 1 -f -o gpurun_out/r02_scan_cfg5shard_cg1 $CMD5 > gpurun_out/r2b_ncu_f5.log 2>&1
tail -2 gpurun_out/r2b_ncu_f5.log
ls -la gpurun_out | grep r02_
