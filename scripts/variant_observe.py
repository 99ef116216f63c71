"""A/B of observe_hist builds: 1 GB of Zipf(1) bytes (iid and pw64k) and uniform bytes; checks counts against numpy."""
import ctypes as C, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import __graft_entry__ as entry
pkg = entry.load_package()
n = 1_000_000_000
inputs = {"zipf1_pw64k": pkg.synth_zipf(n, 1.0, seed=42, segment_bytes=65536), "zipf1_iid": pkg.synth_zipf(n, 1.0, seed=42, segment_bytes=0),
          "zipf3_iid": pkg.synth_zipf(n, 3.0, seed=42, segment_bytes=0), "uniform": pkg.synth_zipf(n, 0.0, seed=42, segment_bytes=0)}
hist = torch.zeros(256, dtype=torch.int32, device="cuda")
st = torch.cuda.current_stream().cuda_stream
for label, data in inputs.items():
    want = np.bincount(data, minlength=256)
    d = torch.from_numpy(data).cuda()
    for name in sys.argv[1:]:
        lib = C.CDLL(os.path.join(ROOT, "variants", f"libhsr_{name}.so"))
        lib.hsr_observe_hist_device.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
        lib.hsr_observe_hist_device(d.data_ptr() + 1, n - 1, hist.data_ptr(), st)  # unaligned start as well
        torch.cuda.synchronize()
        ok1 = bool(np.array_equal(hist.cpu().numpy(), np.bincount(data[1:], minlength=256)))
        lib.hsr_observe_hist_device(d.data_ptr(), n, hist.data_ptr(), st)
        torch.cuda.synchronize()
        ok = ok1 and bool(np.array_equal(hist.cpu().numpy(), want))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            lib.hsr_observe_hist_device(d.data_ptr(), n, hist.data_ptr(), st)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(json.dumps({"input": label, "variant": name, "ms": round(ms, 4), "GBps": round(n / ms / 1e6, 1), "exact": ok}), flush=True)
    del d
