#!/bin/bash
# first GPU pass: smoke, parity tests, short bench, full bench. Everything is logged under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; lscpu | grep -E "Model name|^CPU\(s\)|Thread|Socket" >> gpurun_out/gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python bench.py --size 100000000 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_100m.log 2>&1
timeout 600 python bench.py > gpurun_out/bench_1g.log 2>&1
tail -3 gpurun_out/smoke.log; tail -15 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/bench_100m.log; tail -2 gpurun_out/bench_1g.log
