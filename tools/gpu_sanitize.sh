#!/bin/bash
# compute-sanitizer passes over a small end-to-end search (memcheck, racecheck, synccheck, initcheck)
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool --error-exitcode 9 python tests/harness/gpu_stage.py search 2 3000 256 300 10 > gpurun_out/sanitize_$tool.log 2>&1
  echo "$tool rc=$?" | tee -a gpurun_out/sanitize_summary.txt
  tail -4 gpurun_out/sanitize_$tool.log
done
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python tests/harness/gpu_stage.py search 1 3000 256 100 10 > gpurun_out/sanitize_memcheck_cg1.log 2>&1
echo "memcheck cg1 rc=$?" | tee -a gpurun_out/sanitize_summary.txt; tail -3 gpurun_out/sanitize_memcheck_cg1.log
# the exact fallback scan (locks, last-CTA finalise): every (query, lane) forced through it
for tool in memcheck racecheck; do
  timeout 600 compute-sanitizer --tool $tool --error-exitcode 9 python tests/harness/gpu_stage.py search 2 3000 256 300 10 1 > gpurun_out/sanitize_${tool}_fixup.log 2>&1
  echo "$tool fixup rc=$?" | tee -a gpurun_out/sanitize_summary.txt; tail -3 gpurun_out/sanitize_${tool}_fixup.log
done
