#!/bin/bash
# round 2: regression sweep of BASELINE config 5 at 20 MB per case (all single-stream codecs x bits x entropy), 10k-case soak
mkdir -p gpurun_out
timeout 900 python scripts/sweep_config5.py --size 20000000 --out gpurun_out/r2_config5_sweep_20mb.jsonl > gpurun_out/r2_config5_sweep.log 2>&1
tail -3 gpurun_out/r2_config5_sweep.log
python - <<'PY'
import json
rows = [json.loads(l) for l in open("gpurun_out/r2_config5_sweep_20mb.jsonl") if l.startswith("{")]
print(len(rows), "cases,", sum(1 for r in rows if not r.get("bit_exact", r.get("ok", False))), "not bit-exact")
PY
timeout 600 python scripts/gpu_soak.py --cases 10000 --seed 2026 > gpurun_out/r2_soak_10000.txt 2>&1; tail -2 gpurun_out/r2_soak_10000.txt
