#!/bin/bash
# round 2: compute-sanitizer over every codec family + the new histogram / encoder kernels, full GPU suite, soak, N=1 bench
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python scripts/sanitize_case.py > gpurun_out/r2_sanitize_${tool}.txt 2>&1
  echo "$tool rc=$?" | tee -a gpurun_out/r2_sanitize_${tool}.txt
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|MISMATCH" gpurun_out/r2_sanitize_${tool}.txt | tail -3
done
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2_pytest_gpu_g.txt; tail -3 gpurun_out/r2_pytest_gpu_g.txt
timeout 600 python scripts/gpu_soak.py --cases 3000 --seed 77 > gpurun_out/r2_soak_3000.txt 2>&1; tail -3 gpurun_out/r2_soak_3000.txt
