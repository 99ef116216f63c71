#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/r2_variants_unroll.jsonl
timeout 200 python scripts/variant_bench.py --bits 15 --states 32 u2 u4 u8 >> gpurun_out/r2_variants_unroll.jsonl 2>> gpurun_out/r2_variants_unroll.err
timeout 200 python scripts/variant_bench.py --bits 15 --states 64 u2 u4 u8 >> gpurun_out/r2_variants_unroll.jsonl 2>> gpurun_out/r2_variants_unroll.err
timeout 200 python scripts/variant_bench.py --bits 12 --states 32 u2 u4 u8 >> gpurun_out/r2_variants_unroll.jsonl 2>> gpurun_out/r2_variants_unroll.err
cut -c1-250 gpurun_out/r2_variants_unroll.jsonl; tail -3 gpurun_out/r2_variants_unroll.err
