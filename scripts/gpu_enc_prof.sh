#!/bin/bash
# per-kernel times of the device encoder (fixed blocks and policy mode) on 1 GB
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:enc_ -c 40 --csv --log-file gpurun_out/launches_enc_v12.csv python - > gpurun_out/enc_prof_v12.log 2>&1 <<PY
import sys, torch
sys.path.insert(0, ".")
import __graft_entry__ as g
pkg = g.load_package()
n = 1_000_000_000
for seg in (65536, 0):
    data = pkg.synth_zipf(n, 1.0, 42, seg)
    d_in = torch.from_numpy(data).cuda()
    bound = pkg.encode_mt_bound(64, n)
    d_out = torch.empty(bound, dtype=torch.uint8, device="cuda")
    for _ in range(2):
        c = pkg.encode_mt_device(64, 15, d_in.data_ptr(), n, d_out.data_ptr(), bound, 0, 0)
    for _ in range(2):
        c2 = pkg.encode_mt_policy_device(64, 15, d_in.data_ptr(), n, d_out.data_ptr(), bound, 0, 0)
    print("compressed", seg, c, c2)
PY
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/launches_enc_v12.csv")) if len(r)>5 and not r[0].startswith("==")]
h=rows[0]; ki=h.index("Kernel Name"); vi=h.index("Metric Value")
for r in rows[1:]:
    print(r[ki][:44].ljust(46), round(float(r[vi].replace(",",""))/1e6,3), "ms")
PY
