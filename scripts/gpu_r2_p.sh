#!/bin/bash
# round 2, closing checks on the final tree: sanitizer (incl. the batch pipeline), 20,000-case soak
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python scripts/sanitize_case.py > gpurun_out/r2_sanitize_${tool}_final.txt 2>&1
  echo "$tool rc=$?" | tee -a gpurun_out/r2_sanitize_${tool}_final.txt
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|MISMATCH" gpurun_out/r2_sanitize_${tool}_final.txt | tail -3
done
timeout 900 python scripts/gpu_soak.py --cases 20000 --seed 424242 > gpurun_out/r2_soak_20000_final.txt 2>&1; tail -2 gpurun_out/r2_soak_20000_final.txt
