"""BASELINE.json config 5: probability bits 10-15 x source entropy (Zipf s = 0 .. 3) x N in {32, 64} x {raw, block_},
100 MB each: GPU single-stream decode (one warp: these codecs are one recurrence) next to the reference's fastest
AVX2 decoder on one host core, every output checked byte for byte. Writes one JSON line per case.
    python scripts/sweep_config5.py [--size 100000000] [--out gpurun_out/config5_sweep.jsonl]"""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as g
import checkers as ck
import torch

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=100_000_000)
ap.add_argument("--out", default="gpurun_out/config5_sweep.jsonl")
a = ap.parse_args()
pkg = g.load_package()
n = a.size
out_dev = torch.empty(n + 64, dtype=torch.uint8, device="cuda")
st = torch.cuda.current_stream().cuda_stream
rows = []
with open(a.out, "w") as f:
    for s in (0.0, 0.5, 1.0, 1.5, 2.0, 3.0):
        data = pkg.synth_zipf(n, s, seed=42, segment_bytes=0)
        for states in (32, 64, 16):
            for bits in range(10, 16):
                # 16 states: rANS32x16_16w only; the 32blk layout rides along with the 32-state codecs
                fams = ((ck.RAW, "raw"),) if states == 16 else ((ck.RAW, "raw"), (ck.BLOCK, "block_")) + (((ck.RAW32BLK, "32blk"),) if states == 32 else ())
                for fam, label in fams:
                    try:
                        stream = ck.ref_encode(fam, states, bits, data)
                    except ck.RefEncoderOverflow:
                        continue
                    ps = pkg.PreparedStream.upload(fam, states, bits, stream)
                    ps.decode_async(out_dev.data_ptr(), n, st); torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(); ps.decode_async(out_dev.data_ptr(), n, st); e1.record(); torch.cuda.synchronize()
                    got = out_dev[:n].cpu().numpy()
                    status = ps.status()
                    ps.free()
                    t0 = time.perf_counter()
                    cn, cout = ck.ref_decode(fam, states, bits, stream, n, ck.IMPL_AVX2)
                    cpu_s = time.perf_counter() - t0
                    # truth = what the reference decoder makes of the stream (its 32blk encoder corrupts incompressible input)
                    ok = status == 0 and bool(np.array_equal(got, cout[:n]))
                    ref_round_trip = bool(np.array_equal(cout[:n], data))
                    row = {"zipf_s": s, "states": states, "bits": bits, "codec": label, "ratio": round(stream.size / n, 4),
                           "gpu_one_stream_GBps": round(n / e0.elapsed_time(e1) / 1e6, 3), "cpu_avx2_one_core_GBps": round(n / cpu_s / 1e9, 3),
                           "bit_exact": ok and cn == n, "reference_round_trip_ok": ref_round_trip}
                    f.write(json.dumps(row) + "\n"); f.flush()
                    rows.append(row)
print(json.dumps({"cases": len(rows), "all_bit_exact": all(r["bit_exact"] for r in rows),
                  "gpu_GBps_min_max": [min(r["gpu_one_stream_GBps"] for r in rows), max(r["gpu_one_stream_GBps"] for r in rows)],
                  "cpu_GBps_min_max": [min(r["cpu_avx2_one_core_GBps"] for r in rows), max(r["cpu_avx2_one_core_GBps"] for r in rows)]}))
