#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_encoder.py tests/test_gpu_parity.py -m gpu -x -q -k "encoder or hist or identical or decode_with or block_sizes" > gpurun_out/pytest_enc.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_enc.log
tail -4 gpurun_out/pytest_enc.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:enc_ -c 12 --csv --log-file gpurun_out/launches_enc.csv python - > gpurun_out/enc_prof.log 2>&1 <<PY
import sys, torch
sys.path.insert(0, ".")
import __graft_entry__ as g
pkg = g.load_package()
n = 1_000_000_000
data = pkg.synth_zipf(n, 1.0, 42, 65536)
d_in = torch.from_numpy(data).cuda()
bound = pkg.encode_mt_bound(64, n)
d_out = torch.empty(bound, dtype=torch.uint8, device="cuda")
for _ in range(3):
    c = pkg.encode_mt_device(64, 15, d_in.data_ptr(), n, d_out.data_ptr(), bound, 0, 0)
print("compressed", c)
PY
grep -E "enc_" gpurun_out/launches_enc.csv | awk -F'","' '{print $5, $(NF-0)}' | tail -4
