#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
export HSR_BENCH_TRACE=1
timeout 420 python bench.py --no-cpu-baseline > gpurun_out/bench_r2f.json 2> gpurun_out/bench_r2f.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_r2f.json"))
print("value", d["value"], "e2e", d["e2e"]["value"])
for k, v in d["other_configs"].items():
    if k.endswith("_batch"): print(k, v)
PY
