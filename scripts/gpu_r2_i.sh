#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/r2_variants_grp24.jsonl
timeout 200 python scripts/variant_bench.py --bits 15 --states 64 g16 g24 >> gpurun_out/r2_variants_grp24.jsonl 2>> gpurun_out/r2_variants_grp24.err
timeout 200 python scripts/variant_bench.py --bits 15 --states 32 g16 g24 >> gpurun_out/r2_variants_grp24.jsonl 2>> gpurun_out/r2_variants_grp24.err
cat gpurun_out/r2_variants_grp24.jsonl; tail -3 gpurun_out/r2_variants_grp24.err
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
