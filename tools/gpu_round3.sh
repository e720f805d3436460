#!/bin/bash
P='import json,sys; d=json.loads(sys.stdin.read()); print(sys.argv[1], round(d["value"]), round(d["e2e"]["value"]), round(d["e2e"]["blocking_value"]), round(d["roofline"]["achieved"],1), round(d["roofline"]["frac"],4), d["clocks"]["sm_mhz"], d["clocks"]["power_w_max"])'
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "$P" headline
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu --batch 128 2>&1 | tail -1 | python -c "$P" b128
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu --rows 6250000 --dim 768 --batch 128 --k 5 2>&1 | tail -1 | python -c "$P" cfg5
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu --rows 6250000 --dim 768 --batch 128 --k 5 --preheat 0 2>&1 | tail -1 | python -c "$P" cfg5-nopreheat
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu --rows 1000000 --batch 256 2>&1 | tail -1 | python -c "$P" 1m-b256
