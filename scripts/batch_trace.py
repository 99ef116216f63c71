"""Timeline of one hsr_decode_batch call from pinned host buffers (2368 raw 64x12 streams of 400 KB): device timestamps
of the copy pieces, group launches and copy-outs (HSR_TRACE_PIPELINE), and the wall time of the bare C call.
Development tool."""
import os, sys, time, ctypes as C
os.environ["HSR_TRACE_PIPELINE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import __graft_entry__ as g
import checkers as ck
pkg = g.load_package()
from hypersonic_rans_b200.capi import BatchItem
k_streams, each = 2368, 400_000
data = ck.synth_zipf(k_streams * each, 1.0, seed=43, segment_bytes=0)
fam, states, bits = (int(x) for x in (sys.argv[1:4] if len(sys.argv) > 3 else (0, 64, 12)))
parts, items, pos = [], [], 0
for k in range(k_streams):
    stream = ck.ref_encode(fam, states, bits, data[k * each:(k + 1) * each])
    pad = (-pos) % 16
    parts.append(np.zeros(pad, np.uint8)); pos += pad
    items.append((pos, stream.size, k * each, each))
    parts.append(stream); pos += stream.size
in_base = np.concatenate(parts)
hin, hout = pkg.host_alloc(in_base.size), pkg.host_alloc(data.size)
hin.array[:] = in_base
arr = (BatchItem * k_streams)(*[BatchItem(*map(int, it)) for it in items])
lengths = np.zeros(k_streams, np.uint64)
lib = pkg.lib()
for it in range(3):
    t0 = time.perf_counter()
    ok = lib.hsr_decode_batch(fam, states, bits, hin.ptr, hout.ptr, arr, k_streams, lengths.ctypes.data)
    dt = time.perf_counter() - t0
    print(f"call {it}: {dt * 1e3:.3f} ms wall, {ok} streams", file=sys.stderr, flush=True)
assert np.array_equal(hout.array, data)
