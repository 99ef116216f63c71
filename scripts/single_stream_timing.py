"""Single-stream decode times (one warp each) for the raw layouts, 100 MB Zipf(1) iid; development tool."""
import json, sys, time
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import torch
import __graft_entry__ as g, checkers as ck
pkg = g.load_package()
pkg.set_option("table", int(sys.argv[1]) if len(sys.argv) > 1 else 0)
n = 100_000_000
data = pkg.synth_zipf(n, 1.0, seed=42, segment_bytes=0)
out = torch.empty(n + 64, dtype=torch.uint8, device="cuda")
for label, fam, states, bits in (("rANS32x16_16w_12", 0, 16, 12), ("rANS32x32_32blk_16w_15", 3, 32, 15), ("rANS32x32_32blk_16w_11", 3, 32, 11),
                                 ("rANS32x32_16w_11", 0, 32, 11), ("rANS32x64_16w_12", 0, 64, 12), ("rANS32x32_16w_15", 0, 32, 15),
                                 ("rANS32x64_16w_15", 0, 64, 15), ("rANS32x64_16w_13", 0, 64, 13), ("block_rANS32x64_16w_15", 1, 64, 15),
                                 ("mt_rANS32x64_16w_15 (iid: few huge blocks)", 2, 64, 15)):
    stream = ck.ref_encode(fam, states, bits, data)
    ps = pkg.PreparedStream.upload(fam, states, bits, stream)
    st = torch.cuda.current_stream().cuda_stream
    ps.decode_async(out.data_ptr(), n, st); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ps.decode_async(out.data_ptr(), n, st); e1.record(); torch.cuda.synchronize()
    ok = ps.status() == 0 and bool(np.array_equal(out[:n].cpu().numpy(), data))
    ms = e0.elapsed_time(e1)
    rows = n / states
    print(json.dumps({"codec": label, "ms": round(ms, 2), "GBps": round(n / ms / 1e6, 4), "cycles_per_row": round(ms * 1e-3 * 1.965e9 / rows, 1), "bit_exact": ok, "table_option": pkg.get_option("table"), "units": int(ps.units)}), flush=True)
    ps.free()
