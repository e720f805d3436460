#!/bin/bash
# Round 2, GPU call E (1 GPU): full parity suite on the new epilogue / pre-pass, A/B of the insertion walk, worst-case
# orders, ncu captures of the headline shapes.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2e_pytest.log; tail -6 gpurun_out/r2e_pytest.log
# A/B: extract-max rounds (default build) vs group-wise walk (experiment build), same shapes, interleaved by process
for rep in 1 2; do
  for lib in "" "quickstart-streaming-agents_b200/libsa_b200_groupwise.so"; do
    tag=rounds; [ -n "$lib" ] && tag=groupwise
    SA_LIB_PATH=$lib timeout 300 python tools/gpu_prof.py --shapes cfg2,cfg5,b1024,cfg4 --iters 32 --preheat 1.0 --out gpurun_out/r2e_ab_${tag}_${rep}.json > gpurun_out/r2e_ab_${tag}_${rep}.log 2>&1
  done
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2e_ab_*.json')):
    for r in json.load(open(f)):
        print(f.split('r2e_ab_')[1][:-5], r['shape'], 'scan_ms %.4f' % r['scan_ms'], 'TF %.0f' % r['tflops'], 'GB/s %.0f' % r['gbs'], 'epi_busy %.0f' % r['epi_busy'], 'tile %.0f' % r['cycles_per_tile'])
PY
timeout 600 python tools/gpu_sweep.py --opt presample=0,16,64 --shapes cfg2,cfg5,b1024,cfg4 --out gpurun_out/r2e_sweep_presample.json > gpurun_out/r2e_sweep_presample.log 2>&1
cut -c1-220 gpurun_out/r2e_sweep_presample.log
timeout 900 python tools/gpu_worstcase.py --rows 4000000 --batches 128,1024 --presample 0,16,64 --out gpurun_out/r2e_worstcase.json > gpurun_out/r2e_worstcase.log 2>&1
cut -c1-260 gpurun_out/r2e_worstcase.log
# ncu: headline shape (B = 1024) and the HBM-bound capture (B = 128) at HEAD
for shape in b1024 b128; do
  CMD="python tools/gpu_prof.py --shapes $shape --iters 3 --preheat 0.2"
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:sa_ -c 40 --csv --log-file gpurun_out/r02_launches_${shape}.csv $CMD > gpurun_out/r2e_ncu_l_${shape}.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:sa_scan -s 5 -c 1 -f -o gpurun_out/r02_scan_${shape} $CMD > gpurun_out/r2e_ncu_f_${shape}.log 2>&1
  tail -1 gpurun_out/r2e_ncu_f_${shape}.log
done
ls -la gpurun_out | grep "r02_"
