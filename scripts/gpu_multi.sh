#!/bin/bash
# multi-GPU smoke: torchrun bench at N ranks + the one-process multi-device entry point
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/multi_gpus.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n${N}.log 2>&1
tail -1 gpurun_out/bench_n${N}.log | cut -c1-1800
timeout 600 python - > gpurun_out/multi_entry_n${N}.log 2>&1 <<PY
import sys, time, numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import __graft_entry__ as g, checkers as ck
pkg = g.load_package()
n = 400_000_000
data = pkg.synth_zipf(n, 1.0, 42, 65536)
stream = ck.ref_encode(2, 64, 15, data)
hin, hout = pkg.host_alloc(stream.size), pkg.host_alloc(n)
hin.array[:] = stream
for devs in ([0], list(range($N))):
    for rep in range(3):
        t0 = time.perf_counter()
        got, out = pkg.decode_mt_multi(64, 15, hin.array, n, devices=devs, out=hout.array)
        dt = time.perf_counter() - t0
    ok = got == n and np.array_equal(out, data)
    print({"devices": devs, "ok": bool(ok), "GBps": round(n / dt / 1e9, 2), "err": pkg.last_error()})
PY
cat gpurun_out/multi_entry_n${N}.log | tail -3
