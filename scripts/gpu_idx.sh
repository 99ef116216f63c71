#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "parallel_device_index or prepared_stream" > gpurun_out/pytest_idx.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_idx.log
tail -15 gpurun_out/pytest_idx.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['index_ms'], d['value'])"
