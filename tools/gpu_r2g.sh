#!/bin/bash
# Round 2, GPU call G (1 GPU): smoke, full parity suite on the decoupled-rings kernel, A/B against the one-ring build,
# compute-sanitizer, tight worst-case orders.
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2g_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2g_smoke.log; tail -2 gpurun_out/r2g_smoke.log | cut -c1-400
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2g_pytest.log; tail -5 gpurun_out/r2g_pytest.log
for rep in 1 2; do
  for lib in "" "quickstart-streaming-agents_b200/libsa_b200_onering.so"; do
    tag=tworings; [ -n "$lib" ] && tag=onering
    SA_LIB_PATH=$lib timeout 400 python tools/gpu_prof.py --shapes cfg2,cfg5,b128,n8shard,cfg4,b1024 --iters 32 --preheat 1.0 --out gpurun_out/r2g_ab_${tag}_${rep}.json > gpurun_out/r2g_ab_${tag}_${rep}.log 2>&1
  done
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2g_ab_*.json')):
    for r in json.load(open(f)):
        print(f.split('r2g_ab_')[1][:-5], r['shape'], 'scan_ms %.4f' % r['scan_ms'], 'TF %.0f' % r['tflops'], 'GB/s %.0f' % r['gbs'], 'tile %.0f' % r['cycles_per_tile'], 'prod_wait %.0f' % r['prod_wait_empty'], 'mma_wait_full %.0f' % r['mma_wait_full'], 'mma_wait_epi %.0f' % r['mma_wait_tempty'], 'epi_busy %.0f' % r['epi_busy'], r.get('cta_cycles_min_max'))
PY
bash tools/gpu_sanitize.sh > gpurun_out/r2g_sanitize.log 2>&1; cat gpurun_out/sanitize_summary.txt
timeout 600 python tools/gpu_worstcase.py --rows 4000000 --batches 128,1024 --presample 0,64 --noise 0.1 --out gpurun_out/r2g_worstcase_tight.json > gpurun_out/r2g_worstcase_tight.log 2>&1
cut -c1-260 gpurun_out/r2g_worstcase_tight.log
