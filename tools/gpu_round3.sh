#!/bin/bash
mkdir -p gpurun_out
for i in 1 2; do
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('headline', round(d['value']), round(d['e2e']['value']), round(d['e2e']['blocking_value']), round(d['roofline']['frac'],4), d['clocks']['sm_mhz'])"
done
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu --batch 128 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('b128', round(d['value']), round(d['e2e']['value']), round(d['e2e']['blocking_value']), round(d['roofline']['frac'],4))"
