#!/bin/bash
# round 2, second GPU pass: default bench line with stage trace, reference arm, ncu launch list + full capture of the headline kernel
mkdir -p gpurun_out
export HSR_BENCH_TRACE=1
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
echo "bench rc=$?"; tail -c 1500 gpurun_out/r2_bench_n1.json; tail -25 gpurun_out/r2_bench_n1.err
timeout 300 python bench.py --kernel-only --steps 20 --warmup 3 > gpurun_out/r2_bench_kernel_only.json 2> gpurun_out/r2_bench_kernel_only.err
echo "kernel-only rc=$?"; cat gpurun_out/r2_bench_kernel_only.json; tail -5 gpurun_out/r2_bench_kernel_only.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 2 --warmup 3 --kernel-only > gpurun_out/r2_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:units_n -s 3 -c 2 -f -o gpurun_out/r2_prof \
    python bench.py --steps 2 --warmup 3 --kernel-only --no-overlap > gpurun_out/r2_prof.log 2>&1
ls -la gpurun_out | tail -12
