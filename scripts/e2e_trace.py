"""Timeline of one host-pointer decode (hsr_decode, 1 GB mt_64x15 pw64k, pinned buffers): device timestamps of every
copy piece, launch and copy-out of the three-stream pipeline (HSR_TRACE_PIPELINE), plus the wall time of the call.
Development tool: python scripts/e2e_trace.py > gpurun_out/e2e_trace.txt 2>&1"""
import os, sys, time
os.environ["HSR_TRACE_PIPELINE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import __graft_entry__ as g
import checkers as ck
pkg = g.load_package()
n = 1_000_000_000
data = ck.synth_zipf(n, 1.0, seed=42, segment_bytes=65536)
stream = ck.ref_encode(2, 64, 15, data)
hin, hout = pkg.host_alloc(stream.size), pkg.host_alloc(n)
hin.array[:] = stream
lib = pkg.lib()
for it in range(4):
    t0 = time.perf_counter()
    got = lib.hsr_decode(2, 64, 15, hin.ptr, stream.size, hout.ptr, n)
    dt = time.perf_counter() - t0
    print(f"call {it}: {dt * 1e3:.3f} ms wall, returned {got}", file=sys.stderr, flush=True)
assert np.array_equal(hout.array, data)
