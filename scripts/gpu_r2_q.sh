#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "encoder or harness or smoke or golden" 2>&1 | tail -3
timeout 300 python scripts/gpu_soak.py --cases 3000 --seed 99 2>&1 | tail -1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_launches_encoder2.csv \
    python scripts/gpu_encoder_once.py > gpurun_out/r2_launches_encoder2.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open("gpurun_out/r2_launches_encoder2.csv")) if len(r)>5 and not r[0].startswith("==")]
h=rows[0]; ki=h.index("Kernel Name"); vi=h.index("Metric Value")
for r in rows[-16:]:
    print(r[ki][:50].ljust(52), round(float(r[vi].replace(",",""))/1e6,3), "ms")
PY
