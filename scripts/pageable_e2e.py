"""hsr_decode end to end from PAGEABLE host buffers (what the reference harness allocates, src/main.cpp:649-650) next to
pinned ones: 1 GB mt_64x15 pw64k. Development tool."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import __graft_entry__ as g
import checkers as ck
pkg = g.load_package()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000_000
data = ck.synth_zipf(n, 1.0, seed=42, segment_bytes=65536)
stream = ck.ref_encode(2, 64, 15, data)
lib = pkg.lib()
hin, hout = pkg.host_alloc(stream.size), pkg.host_alloc(n)
hin.array[:] = stream
pin = np.array(stream, copy=True); pout = np.empty(n, np.uint8); pout[:] = 0
for label, i_ptr, o_ptr, o_arr in (("pinned", hin.ptr, hout.ptr, hout.array), ("pageable", pin.ctypes.data, pout.ctypes.data, pout)):
    for it in range(4):
        t0 = time.perf_counter()
        got = lib.hsr_decode(2, 64, 15, i_ptr, stream.size, o_ptr, n)
        dt = time.perf_counter() - t0
        print(f"{label} call {it}: {dt * 1e3:.2f} ms = {n / dt / 1e9:.2f} GB/s, returned {got}", flush=True)
    assert np.array_equal(o_arr, data)
