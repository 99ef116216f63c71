#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_tma.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_tma.log
tail -5 gpurun_out/pytest_tma.log
: > gpurun_out/sweep_tma.log
for cfg in "64 15 0" "64 12 0" "64 10 0" "32 15 0"; do
  set -- $cfg
  timeout 300 python bench.py --kernel-only --steps 10 --warmup 3 --states $1 --bits $2 --table $3 >> gpurun_out/sweep_tma.log 2>&1
done
grep kernel_only gpurun_out/sweep_tma.log | cut -c1-120
