#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-iter}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_${TAG}.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_${TAG}.log
tail -5 gpurun_out/pytest_${TAG}.log
timeout 900 python bench.py --extra > gpurun_out/bench_${TAG}.log 2>&1; tail -1 gpurun_out/bench_${TAG}.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(json.dumps({k:d[k] for k in ('value','e2e','single_recurrence')}, indent=1))"
