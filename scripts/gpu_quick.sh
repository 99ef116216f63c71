#!/bin/bash
# quick A/B: parity subset + kernel-only headline
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "prepared or golden or batch" > gpurun_out/pytest_quick.log 2>&1; tail -2 gpurun_out/pytest_quick.log
for i in 1 2; do timeout 300 python bench.py --kernel-only --steps 20 --warmup 3 2>&1 | tail -1 | cut -c1-110; done
