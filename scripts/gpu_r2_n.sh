#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:units_n -s 1 -c 1 -f -o gpurun_out/prof_r2_lone_n64b12 \
    python scripts/lone_warp_once.py 20000000 64 12 0 > gpurun_out/prof_r2_lone.log 2>&1
tail -3 gpurun_out/prof_r2_lone.log
