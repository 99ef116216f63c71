#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-iter}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_${TAG}.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_${TAG}.log
tail -4 gpurun_out/pytest_${TAG}.log
timeout 120 python scripts/pcie_probe.py > gpurun_out/pcie_${TAG}.log 2>&1; cat gpurun_out/pcie_${TAG}.log
timeout 900 python bench.py --extra > gpurun_out/bench_${TAG}.log 2>&1; tail -1 gpurun_out/bench_${TAG}.log
