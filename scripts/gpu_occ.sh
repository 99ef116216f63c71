#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/occ_sweep.log
for k in 4 8 12 16 0; do
  timeout 300 python bench.py --kernel-only --steps 10 --warmup 3 --ctas-per-sm $k >> gpurun_out/occ_sweep.log 2>&1
done
grep kernel_only gpurun_out/occ_sweep.log | cut -c1-170
