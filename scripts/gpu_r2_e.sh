#!/bin/bash
# round 2, fifth GPU pass: host-pointer pipeline (ranges launched from inside the walk): tests + bench line
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2_pytest_gpu_e.txt; tail -5 gpurun_out/r2_pytest_gpu_e.txt
export HSR_BENCH_TRACE=1
timeout 400 python bench.py --steps 20 --warmup 3 --headline-only --no-cpu-baseline > gpurun_out/r2_bench_n1_e.json 2> gpurun_out/r2_bench_n1_e.err
echo "bench rc=$?"; python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r2_bench_n1_e.json"))
    print("value", d["value"], "serialized", d["value_serialized"], "e2e", d["e2e"])
except Exception as e:
    print("no bench json:", e)
PY
tail -3 gpurun_out/r2_bench_n1_e.err
