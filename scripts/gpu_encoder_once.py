"""One pass of the device encoder (fixed blocks, then policy mode) and the histogram entry points over 1 GB, for an ncu
launch list (scripts/gpu_r2_d.sh). Development tool."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
pkg = g.load_package()
n = 1_000_000_000
data = pkg.synth_zipf(n, 1.0, 42, 65536)
d_in = torch.from_numpy(data).cuda()
bound = pkg.encode_mt_bound(64, n)
d_out = torch.empty(bound, dtype=torch.uint8, device="cuda")
hist = torch.zeros(256, dtype=torch.int32, device="cuda")
counts = torch.zeros(((n + 65535) // 65536, 256), dtype=torch.int16, device="cuda")
for _ in range(2):
    c = pkg.encode_mt_device(64, 15, d_in.data_ptr(), n, d_out.data_ptr(), bound, 0, 0)
    c2 = pkg.encode_mt_policy_device(64, 15, d_in.data_ptr(), n, d_out.data_ptr(), bound, 0, 0)
    pkg.observe_hist_device(d_in.data_ptr(), n, hist.data_ptr(), 0)
    pkg.make_hist_segments_device(d_in.data_ptr(), n, 65536, 15, counts.data_ptr(), 0)
torch.cuda.synchronize()
print("compressed", c, c2)
