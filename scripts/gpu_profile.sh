#!/bin/bash
# ncu evidence for the dominant kernel: launch list of the bench command + one full-set capture.
mkdir -p gpurun_out
TAG=${1:-v1}
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --headline-only > gpurun_out/launches_${TAG}.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:units_n -s 3 -c 2 -f -o gpurun_out/prof_${TAG} \
    python bench.py --steps 2 --warmup 3 --kernel-only > gpurun_out/prof_${TAG}.log 2>&1
ls -la gpurun_out/
