#!/bin/bash
# one GPU visit: tests, smoke, sanitizer (memcheck + racecheck on the small set), bench with extras, reference arm, ncu evidence
mkdir -p gpurun_out
TAG=${1:-v10}
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_${TAG}.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_${TAG}.log
tail -3 gpurun_out/pytest_${TAG}.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python scripts/sanitize_case.py > gpurun_out/sanitize_${tool}_${TAG}.log 2>&1
  echo "$tool rc=$?" | tee -a gpurun_out/sanitize_${tool}_${TAG}.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY" gpurun_out/sanitize_${tool}_${TAG}.log | tail -1
done
timeout 900 python bench.py --extra > gpurun_out/bench_${TAG}.log 2>&1; tail -1 gpurun_out/bench_${TAG}.log | cut -c1-400
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_${TAG}.log 2>&1; tail -1 gpurun_out/bench_ref_${TAG}.log | cut -c1-200
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --headline-only > gpurun_out/launches_${TAG}.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:units_n -s 3 -c 2 -f -o gpurun_out/prof_${TAG} \
    python bench.py --steps 2 --warmup 3 --kernel-only > gpurun_out/prof_${TAG}.log 2>&1
ls -la gpurun_out | grep ${TAG}
