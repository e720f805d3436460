#!/bin/bash
# Round 2, first GPU call: parity tests, role profile of the new epilogue, quick bench lines.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/r2a_gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
tail -5 gpurun_out/r2a_pytest.log
timeout 600 python tools/gpu_prof.py --shapes cfg5,b128,b1024,cfg2 --out gpurun_out/r2a_prof.json > gpurun_out/r2a_prof.log 2>&1
cat gpurun_out/r2a_prof.log | cut -c1-600
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r2a_bench_b1024.json 2> gpurun_out/r2a_bench_b1024.err
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --rows 6250000 --dim 768 --batch 128 --k 5 > gpurun_out/r2a_bench_cfg5.json 2> gpurun_out/r2a_bench_cfg5.err
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --batch 128 > gpurun_out/r2a_bench_b128.json 2> gpurun_out/r2a_bench_b128.err
for f in gpurun_out/r2a_bench_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['clocks'])
    print({k:d['roofline'][k] for k in ('bound','achieved','frac','launch_ms','scan_share_of_step','achieved_gbs','achieved_tflops')})
except Exception as e:
    print('ERR', e); print(open(sys.argv[1].replace('.json','.err')).read()[-2000:])
PY
done
