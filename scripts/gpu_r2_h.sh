#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/r2_variants_bits.jsonl
for b in 14 12 10; do
  timeout 200 python scripts/variant_bench.py --bits $b r1 cur >> gpurun_out/r2_variants_bits.jsonl 2>> gpurun_out/r2_variants_bits.err
done
cat gpurun_out/r2_variants_bits.jsonl
tail -3 gpurun_out/r2_variants_bits.err
