#!/bin/bash
# iteration pass: parity tests, then kernel-only throughput for the headline config and a small sweep
mkdir -p gpurun_out
TAG=${1:-iter}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_${TAG}.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_${TAG}.log
tail -4 gpurun_out/pytest_${TAG}.log
: > gpurun_out/sweep_${TAG}.log
timeout 300 python bench.py --kernel-only --steps 20 --warmup 3 >> gpurun_out/sweep_${TAG}.log 2>&1
for cfg in "64 12 0" "64 12 1" "64 10 0" "64 13 0" "64 14 0" "32 15 0" "32 12 0" "32 12 1"; do
  set -- $cfg
  timeout 300 python bench.py --kernel-only --size 400000000 --steps 10 --warmup 3 --states $1 --bits $2 --table $3 >> gpurun_out/sweep_${TAG}.log 2>&1
done
grep kernel_only gpurun_out/sweep_${TAG}.log | cut -c1-330
