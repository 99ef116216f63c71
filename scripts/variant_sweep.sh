#!/bin/bash
# A/B of build variants over probability bits x state counts (400 MB mt_ streams); usage: variant_sweep.sh name1 name2 ...
mkdir -p gpurun_out
: > gpurun_out/variant_sweep.jsonl
for N in 64 32; do for b in 10 11 12 13 14 15; do
  python scripts/variant_bench.py "$@" --bits $b --states $N --size 400000000 --steps 30 --rounds 2 2>&1 | grep variant >> gpurun_out/variant_sweep.jsonl
done; done
cat gpurun_out/variant_sweep.jsonl | cut -c1-140
