#!/bin/bash
# end-of-iteration evidence: full GPU tests, bench with extras, ncu launch list + full capture of the dominant kernel
mkdir -p gpurun_out
TAG=${1:-final}
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_${TAG}.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_${TAG}.log
tail -4 gpurun_out/pytest_${TAG}.log
timeout 900 python bench.py --extra > gpurun_out/bench_${TAG}.log 2>&1; tail -1 gpurun_out/bench_${TAG}.log | cut -c1-300
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_${TAG}.log 2>&1; tail -1 gpurun_out/bench_ref_${TAG}.log | cut -c1-200
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --headline-only > gpurun_out/launches_${TAG}.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:units_n -s 3 -c 2 -f -o gpurun_out/prof_${TAG} \
    python bench.py --steps 2 --warmup 3 --kernel-only > gpurun_out/prof_${TAG}.log 2>&1
ls -la gpurun_out | grep ${TAG}
