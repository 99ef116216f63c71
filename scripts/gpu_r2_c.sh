#!/bin/bash
# round 2, third GPU pass: A/B of the loop-exit fix, histogram microbenchmark, default bench (both arms), ncu evidence
mkdir -p gpurun_out
timeout 300 python scripts/variant_bench.py r1 r2shfl redux > gpurun_out/r2_variants_exit.jsonl 2> gpurun_out/r2_variants_exit.err
cat gpurun_out/r2_variants_exit.jsonl; tail -3 gpurun_out/r2_variants_exit.err
timeout 300 ./scripts/ubench_hist.bin > gpurun_out/r2_ubench_hist.jsonl 2>&1
cat gpurun_out/r2_ubench_hist.jsonl | cut -c1-230
export HSR_BENCH_TRACE=1
timeout 420 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
echo "bench rc=$?"; tail -c 2500 gpurun_out/r2_bench_n1.json; tail -4 gpurun_out/r2_bench_n1.err
timeout 200 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 2 --warmup 3 --kernel-only > gpurun_out/r2_launches.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:units_n -s 3 -c 2 -f -o gpurun_out/r2_prof \
    python bench.py --steps 2 --warmup 3 --kernel-only --no-overlap > gpurun_out/r2_prof.log 2>&1
ls -la gpurun_out | tail -14
