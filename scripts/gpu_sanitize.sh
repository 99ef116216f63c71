#!/bin/bash
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  timeout 1200 compute-sanitizer --tool $tool --error-exitcode 9 python scripts/sanitize_case.py > gpurun_out/sanitize_${tool}.log 2>&1
  echo "$tool rc=$?" | tee -a gpurun_out/sanitize_${tool}.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard" gpurun_out/sanitize_${tool}.log | tail -3
done
: > gpurun_out/sweep_1g.log
for cfg in "64 10 0" "64 10 1" "64 11 0" "64 11 1" "64 12 0" "64 12 1" "64 13 0" "64 14 0" "32 15 0" "32 12 1" "32 10 0"; do
  set -- $cfg
  timeout 300 python bench.py --kernel-only --steps 10 --warmup 3 --states $1 --bits $2 --table $3 >> gpurun_out/sweep_1g.log 2>&1
done
grep kernel_only gpurun_out/sweep_1g.log | cut -c1-200
