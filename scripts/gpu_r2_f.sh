#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/e2e_trace.py > gpurun_out/r2_e2e_trace.txt 2>&1
grep "call" gpurun_out/r2_e2e_trace.txt
awk '/call 2/{f=1;next} /call 3/{f=0} f' gpurun_out/r2_e2e_trace.txt | head -150
timeout 300 python -m pytest tests -m gpu -x -q -k "histogram or concurrent or hist" 2>&1 | tail -4
