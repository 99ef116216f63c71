#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_encoder.py -m gpu -x -q > gpurun_out/pytest_enc.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_enc.log
tail -25 gpurun_out/pytest_enc.log
