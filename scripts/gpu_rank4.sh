#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_rank4.py tests/test_cpp_dropin.py -m gpu -x -q > gpurun_out/pytest_rank4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_rank4.log
tail -15 gpurun_out/pytest_rank4.log
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python scripts/sanitize_case.py > gpurun_out/sanitize_${tool}_rank4.log 2>&1
  echo "$tool rc=$?" | tee -a gpurun_out/sanitize_${tool}_rank4.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|MISMATCH" gpurun_out/sanitize_${tool}_rank4.log | tail -3
done
