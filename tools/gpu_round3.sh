#!/bin/bash
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -x -q -m gpu ) 2>&1 | tail -3
P='import json,sys; d=json.loads(sys.stdin.read()); print(sys.argv[1], round(d["value"]), round(d["e2e"]["value"]), round(d["roofline"]["achieved"],1), round(d["roofline"]["frac"],4), d["config"]["rows_per_gpu"])'
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "$P" "10M B=1024 qpu auto"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --batch 4096 2>&1 | tail -1 | python -c "$P" "10M B=4096 qpu auto"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --batch 2048 2>&1 | tail -1 | python -c "$P" "10M B=2048 qpu auto"
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu --rows 1250000 2>&1 | tail -1 | python -c "$P" "1.25M B=1024 qpu auto"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "$P" "10M B=1024 qpu auto (again)"
