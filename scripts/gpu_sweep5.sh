#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 1500 python scripts/sweep_config5.py > gpurun_out/config5_summary.log 2>&1; tail -2 gpurun_out/config5_summary.log
