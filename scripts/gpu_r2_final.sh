#!/bin/bash
# round 2, evidence pass on the current tree: smoke, default bench (both arms), launch list + full ncu capture, bits sweep
TAG=${1:-r2e}
mkdir -p gpurun_out
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 420 python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; echo "bench rc=$?"
timeout 200 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_${TAG}.json 2> gpurun_out/bench_ref_${TAG}.err; echo "ref rc=$?"
python - <<PY
import json
d = json.load(open("gpurun_out/bench_${TAG}.json")); r = json.load(open("gpurun_out/bench_ref_${TAG}.json"))
print("value", d["value"], "serialized", d["value_serialized"]["value"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["value"], "ref", r["value"], "e2e/ref", round(d["e2e"]["value"] / r["value"], 1))
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --headline-only > gpurun_out/launches_${TAG}.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:units_n -s 3 -c 2 -f -o gpurun_out/prof_${TAG} \
    python bench.py --steps 2 --warmup 3 --kernel-only --no-overlap > gpurun_out/prof_${TAG}.log 2>&1
: > gpurun_out/sweep_${TAG}.jsonl
for cfg in "64 10" "64 11" "64 12" "64 13" "64 14" "32 15" "32 12" "32 10"; do
  set -- $cfg
  timeout 120 python bench.py --kernel-only --steps 20 --warmup 3 --states $1 --bits $2 >> gpurun_out/sweep_${TAG}.jsonl 2>/dev/null
done
cut -c1-260 gpurun_out/sweep_${TAG}.jsonl
