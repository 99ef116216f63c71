"""A/B timing of kernel build variants (variants/libhsr_<name>.so, built with `make EXTRA=-D...`) on ONE reference-encoded
stream: the same compressed bytes, the same block index, the same launches, only the shared object differs.

    python scripts/variant_bench.py [--size N] [--bits 15] [--states 64] [--steps 20] name1 name2 ...

Prints one JSON line per variant (kernel-path decoded GB/s, bit-exact flag). Development tool, not part of the product.
"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def bind(path):
    lib = C.CDLL(path)
    lib.hsr_stream_upload.restype = C.c_void_p
    lib.hsr_stream_upload.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_int, C.c_int]
    lib.hsr_stream_decode_async.restype = C.c_int
    lib.hsr_stream_decode_async.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint, C.c_void_p]
    lib.hsr_stream_status.restype = C.c_uint
    lib.hsr_stream_status.argtypes = [C.c_void_p]
    lib.hsr_stream_free.restype = None
    lib.hsr_stream_free.argtypes = [C.c_void_p]
    lib.hsr_set_option.restype = C.c_int
    lib.hsr_set_option.argtypes = [C.c_char_p, C.c_long]
    return lib


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("names", nargs="+")
    ap.add_argument("--size", type=int, default=1_000_000_000)
    ap.add_argument("--bits", type=int, default=15)
    ap.add_argument("--states", type=int, default=64)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--rounds", type=int, default=2, help="each variant is timed this many times, interleaved")
    a = ap.parse_args()

    import torch
    import __graft_entry__ as entry
    import checkers as ck
    pkg = entry.load_package()
    data = pkg.synth_zipf(a.size, 1.0, seed=42, segment_bytes=65536)
    stream = ck.ref_encode(2, a.states, a.bits, data)
    n = data.size
    out = torch.empty(n + 64, dtype=torch.uint8, device="cuda")
    d_ref = torch.from_numpy(data).cuda()
    st = torch.cuda.current_stream().cuda_stream
    results = {}
    for rnd in range(a.rounds):
        for name in a.names:
            lib = bind(os.path.join(ROOT, "variants", f"libhsr_{name}.so"))
            h = lib.hsr_stream_upload(2, a.states, a.bits, stream.ctypes.data, stream.size, 0, 1)
            assert h, name
            out.zero_()
            for _ in range(3):
                assert lib.hsr_stream_decode_async(h, out.data_ptr(), n, 0, st) >= 0, name
            torch.cuda.synchronize()
            ok = lib.hsr_stream_status(h) == 0 and bool(torch.equal(out[:n], d_ref))
            times = []
            for _ in range(a.steps):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                lib.hsr_stream_decode_async(h, out.data_ptr(), n, 0, st)
                e1.record()
                torch.cuda.synchronize()
                times.append(e0.elapsed_time(e1))
            # the same launches back to back (no host sync in between): what consecutive decodes cost when they may overlap
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(a.steps):
                lib.hsr_stream_decode_async(h, out.data_ptr(), n, 0, st)
            e1.record()
            torch.cuda.synchronize()
            b2b = e0.elapsed_time(e1) / a.steps
            lib.hsr_stream_free(h)
            r = results.setdefault(name, {"ms": [], "ok": True, "b2b": []})
            r["b2b"].append(b2b)
            r["ms"] += times
            r["ok"] = r["ok"] and ok
    for name in a.names:
        r = results[name]
        ms = float(np.mean(r["ms"]))
        print(json.dumps({"variant": name, "bits": a.bits, "states": a.states, "decoded_GBps": round(n / ms / 1e6, 1),
                          "ms_mean": round(ms, 4), "ms_min": round(min(r["ms"]), 4), "bit_exact": r["ok"],
                          "back_to_back_ms": round(min(r["b2b"]), 4), "back_to_back_GBps": round(n / min(r["b2b"]) / 1e6, 1),
                          "compressed": int(stream.size)}), flush=True)


if __name__ == "__main__":
    main()
