"""One single-stream decode (raw rANS32x64_16w 12-bit, 100 MB: BASELINE config 2) for an ncu capture of the lone warp."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import __graft_entry__ as g
import checkers as ck
pkg = g.load_package()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
states, bits, fam = (int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (64, 12, 0)
data = ck.synth_zipf(n, 1.0, seed=42, segment_bytes=0)
stream = ck.ref_encode(fam, states, bits, data)
ps = pkg.PreparedStream.upload(fam, states, bits, stream)
out = torch.empty(n + 64, dtype=torch.uint8, device="cuda")
for _ in range(2):
    ps.decode_async(out.data_ptr(), n, 0)
torch.cuda.synchronize()
assert ps.status() == 0 and np.array_equal(out[:n].cpu().numpy(), data)
print("ok")
