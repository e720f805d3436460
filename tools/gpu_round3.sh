#!/bin/bash
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -x -q -m gpu ) 2>&1 | tail -3
P='import json,sys; d=json.loads(sys.stdin.read()); print(sys.argv[1], round(d["value"]), round(d["e2e"]["value"]), round(d["roofline"]["achieved"],1), round(d["roofline"]["frac"],4))'
for S in "" "--no-share"; do
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu --rows 1000000 --batch 256 $S 2>&1 | tail -1 | python -c "$P" "1M B=256 $S"
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu --rows 1250000 --batch 1024 $S 2>&1 | tail -1 | python -c "$P" "1.25M B=1024 $S"
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu --rows 6250000 --dim 768 --batch 128 --k 5 $S 2>&1 | tail -1 | python -c "$P" "cfg5 $S"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu $S 2>&1 | tail -1 | python -c "$P" "10M B=1024 $S"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --batch 128 $S 2>&1 | tail -1 | python -c "$P" "10M B=128 $S"
done
