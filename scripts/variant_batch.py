"""A/B timing of build variants on a batch of independent raw streams (one launch, one warp per stream) and on the device
mt_ encoder; usage: python scripts/variant_batch.py name1 name2 ...   (development tool, see variant_bench.py)"""
import ctypes as C, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import __graft_entry__ as entry
import checkers as ck
pkg = entry.load_package()

class Item(C.Structure):
    _fields_ = [("inOffset", C.c_uint64), ("inLength", C.c_uint64), ("outOffset", C.c_uint64), ("outCapacity", C.c_uint64)]

k_streams, each = 2368, 400_000
data = pkg.synth_zipf(k_streams * each, 1.0, seed=43, segment_bytes=0)
parts, items, pos = [], [], 0
one = {}
for k in range(k_streams):
    stream = ck.ref_encode(0, 64, 12, data[k * each:(k + 1) * each])
    pad = (-pos) % 16
    parts.append(np.zeros(pad, np.uint8)); pos += pad
    items.append((pos, stream.size, k * each, each))
    parts.append(stream); pos += stream.size
in_base = np.concatenate(parts)
arr = (Item * k_streams)(*[Item(*map(int, it)) for it in items])
total = k_streams * each
out = torch.empty(total + 64, dtype=torch.uint8, device="cuda")
d_ref = torch.from_numpy(data).cuda()
st = torch.cuda.current_stream().cuda_stream
n_enc = 1_000_000_000
enc_in = torch.from_numpy(pkg.synth_zipf(n_enc, 1.0, seed=42, segment_bytes=65536)).cuda()
for name in sys.argv[1:]:
    lib = C.CDLL(os.path.join(ROOT, "variants", f"libhsr_{name}.so"))
    lib.hsr_stream_upload_batch.restype = C.c_void_p
    lib.hsr_stream_upload_batch.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t]
    lib.hsr_stream_decode_async.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint, C.c_void_p]
    lib.hsr_stream_status.restype = C.c_uint; lib.hsr_stream_status.argtypes = [C.c_void_p]
    lib.hsr_stream_free.argtypes = [C.c_void_p]
    lib.hsr_encode_mt_bound.restype = C.c_size_t; lib.hsr_encode_mt_bound.argtypes = [C.c_int, C.c_size_t, C.c_size_t]
    lib.hsr_encode_mt_device.restype = C.c_size_t
    lib.hsr_encode_mt_device.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p]
    h = lib.hsr_stream_upload_batch(0, 64, 12, in_base.ctypes.data, arr, k_streams)
    assert h
    out.zero_()
    for _ in range(2):
        lib.hsr_stream_decode_async(h, out.data_ptr(), total, 0, st)
    torch.cuda.synchronize()
    ok = lib.hsr_stream_status(h) == 0 and bool(torch.equal(out[:total], d_ref))
    ms = []
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); lib.hsr_stream_decode_async(h, out.data_ptr(), total, 0, st); e1.record()
        torch.cuda.synchronize(); ms.append(e0.elapsed_time(e1))
    lib.hsr_stream_free(h)
    bound = lib.hsr_encode_mt_bound(64, n_enc, 0)
    enc_out = torch.empty(bound, dtype=torch.uint8, device="cuda")
    import time
    ts = []
    for _ in range(6):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        comp = lib.hsr_encode_mt_device(64, 15, enc_in.data_ptr(), n_enc, enc_out.data_ptr(), bound, 0, st)
        torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    del enc_out
    print(json.dumps({"variant": name, "raw64x12_batch_GBps": round(total / np.mean(ms) / 1e6, 1), "batch_ms": [round(x, 3) for x in ms[:4]], "bit_exact": ok,
                      "encode_ms_min": round(min(ts[1:]) * 1e3, 3), "encode_comp": int(comp)}), flush=True)
