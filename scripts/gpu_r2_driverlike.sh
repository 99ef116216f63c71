#!/bin/bash
# exactly what the driver runs at N GPUs: both arms through torchrun, default flags
N=${1:-8}
mkdir -p gpurun_out
T0=$(date +%s)
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --impl reference --gpus $N --steps 5 --warmup 1 > gpurun_out/driverlike_ref_n${N}.json 2> gpurun_out/driverlike_ref_n${N}.err
echo "reference arm rc=$? after $(( $(date +%s) - T0 )) s; stdout lines: $(wc -l < gpurun_out/driverlike_ref_n${N}.json)"
T0=$(date +%s)
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 \
    bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/driverlike_n${N}.json 2> gpurun_out/driverlike_n${N}.err
echo "our arm rc=$? after $(( $(date +%s) - T0 )) s; stdout lines: $(wc -l < gpurun_out/driverlike_n${N}.json)"
python - <<PY
import json
d = json.load(open("gpurun_out/driverlike_n${N}.json")); r = json.load(open("gpurun_out/driverlike_ref_n${N}.json"))
print("value", d["value"], "scaling", d["scaling"], "e2e", d["e2e"]["value"], "ref", r["value"], "clocks", d["clocks"])
PY
grep -v "^\[bench\|^W\|^\*\|OMP_NUM" gpurun_out/driverlike_n${N}.err | tail -5
