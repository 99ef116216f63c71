#!/bin/bash
# round 2, fourth GPU pass: whole GPU suite on the new histogram kernels, bench --extra (encoder + histogram figures),
# launch list of the encoder / histogram kernels
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2_pytest_gpu_d.txt; tail -5 gpurun_out/r2_pytest_gpu_d.txt
export HSR_BENCH_TRACE=1
timeout 600 python bench.py --steps 20 --warmup 3 --extra > gpurun_out/r2_bench_n1_extra.json 2> gpurun_out/r2_bench_n1_extra.err
echo "bench rc=$?"; python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r2_bench_n1_extra.json"))
    oc = d["other_configs"]
    for k in ("observe_hist", "segment_hists_64k", "device_encoder_pw64k", "device_encoder_policy_pw64k", "device_encoder_iid", "device_encoder_policy_iid"):
        print(k, oc.get(k))
    print("value", d["value"], "e2e", d["e2e"]["value"])
except Exception as e:
    print("no bench json:", e)
PY
tail -3 gpurun_out/r2_bench_n1_extra.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_launches_encoder.csv \
    python scripts/gpu_encoder_once.py > gpurun_out/r2_launches_encoder.log 2>&1
cut -d, -f5,15 gpurun_out/r2_launches_encoder.csv | tail -30
