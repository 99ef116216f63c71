#!/bin/bash
# last visit of round 2: what the driver runs at round end on one GPU (suite, smoke, both bench arms), on the final tree
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r2_pytest_gpu_last.txt
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 200 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_last.json 2> gpurun_out/bench_ref_last.err; echo "ref rc=$?"
timeout 420 python bench.py > gpurun_out/bench_last.json 2> gpurun_out/bench_last.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_last.json")); r = json.load(open("gpurun_out/bench_ref_last.json"))
print("value", d["value"], "serialized", d["value_serialized"]["value"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["value"], "ref", r["value"], "e2e/ref", round(d["e2e"]["value"] / r["value"], 1), "launches", d["gpu_launches"])
PY
