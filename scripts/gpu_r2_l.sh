#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/batch_trace.py 0 64 12 > gpurun_out/r2_batch_trace.txt 2>&1
grep "call" gpurun_out/r2_batch_trace.txt
awk '/call 1/{f=1;next} /call 2/{f=0} f' gpurun_out/r2_batch_trace.txt | grep -v "h2d end" | head -40
awk '/call 1/{f=1;next} /call 2/{f=0} f' gpurun_out/r2_batch_trace.txt | grep "h2d end" | tail -3
