#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:units_n -s 3 -c 1 -f -o gpurun_out/prof_r2_b14 \
    python bench.py --steps 2 --warmup 3 --kernel-only --no-overlap --bits 14 > gpurun_out/prof_r2_b14.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:units_n -s 3 -c 1 -f -o gpurun_out/prof_r2_n32b15 \
    python bench.py --steps 2 --warmup 3 --kernel-only --no-overlap --bits 15 --states 32 > gpurun_out/prof_r2_n32b15.log 2>&1
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
ls -la gpurun_out | grep prof_r2
