#!/bin/bash
# round 2, first GPU pass: the whole -m gpu suite, then the default bench line (both arms)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/r2_gpu_box.txt 2>&1
nproc >> gpurun_out/r2_gpu_box.txt
timeout 1500 python -m pytest tests -m gpu -x -q -s 2>&1 | tail -40 > gpurun_out/r2_pytest_gpu.txt
tail -15 gpurun_out/r2_pytest_gpu.txt
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
tail -c 3000 gpurun_out/r2_bench_n1.json; tail -5 gpurun_out/r2_bench_n1.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err
cat gpurun_out/r2_bench_ref.json; tail -3 gpurun_out/r2_bench_ref.err
