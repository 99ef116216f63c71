"""Small decode set for compute-sanitizer (memcheck / racecheck / initcheck): every family, both N, bits 10/12/15."""
import sys
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import __graft_entry__ as g
pkg = g.load_package()
z = np.load("tests/golden/golden.npz")
bad = 0
for key in sorted(z.files):
    if not key.startswith("stream/"):
        continue
    _, name, fam, states, bits = key.split("/")
    if name not in ("multi", "runs", "tiny65", "small") or int(bits) not in (10, 12, 15):
        continue
    data = z[f"in/{name}"]
    for table in (0, 1, 3):
        pkg.set_option("table", table)
        n, out = pkg.decode(int(fam), int(states), int(bits), z[key], data.size)
        ok = n == data.size and np.array_equal(out[:n], data)
        bad += not ok
        print(key, "table", table, "ok" if ok else "MISMATCH")
z4 = np.load("tests/golden/golden_rank4.npz")  # rANS32x16_16w and rANS32x32_32blk_16w
for key in sorted(z4.files):
    if not key.startswith("stream/"):
        continue
    _, name, fam, states, bits = key.split("/")
    if name not in ("multi", "tiny65", "tiny33", "small") or int(bits) not in (10, 12, 15):
        continue
    data = z4[f"in/{name}"]
    n, out = pkg.decode(int(fam), int(states), int(bits), z4[key], data.size)
    ok = n == data.size and np.array_equal(out[:n], data)
    bad += not ok
    print(key, "ok" if ok else "MISMATCH")
cnt, cum = pkg.make_hist(z["in/multi"], 12)
print("hist ok", np.array_equal(cnt, z["hist/multi/12"][0]))
# round 2: per-range histograms (counter columns + one lane per histogram) on odd ranges, and the device encoders
import torch
data = z["in/multi"]
d = torch.from_numpy(np.ascontiguousarray(data)).cuda()
for seg, off in ((4099, 3), (65536, 0), (17, 1)):
    nbytes = min(data.size - off, seg * 37 + 5)
    nseg = (nbytes + seg - 1) // seg
    counts = torch.zeros((nseg, 256), dtype=torch.int16, device="cuda")
    rc = pkg.make_hist_segments_device(d.data_ptr() + off, nbytes, seg, 13, counts.data_ptr(), 0)
    torch.cuda.synchronize()
    print("segments", seg, off, "rc", rc)
    bad += rc <= 0
for states, bits in ((32, 10), (64, 15)):
    for fn in (pkg.encode_mt, pkg.encode_mt_policy):
        stream = fn(states, bits, data, 0)
        n, out = pkg.decode(2, states, bits, stream, data.size)
        ok = n == data.size and np.array_equal(out[:n], data)
        bad += not ok
        print("encoder", fn.__name__, states, bits, "ok" if ok else "MISMATCH")
# round 2: the pipelined batch entry point (groups, per-stream status in mapped host memory), one stream corrupted
for fam, states, bits in ((0, 64, 12), (1, 32, 10), (2, 64, 15)):
    key = f"stream/multi/{fam}/{states}/{bits}"
    streams = [z[key].copy() for _ in range(6)]
    streams[2][16 + (16 + 4 * states if fam == 2 else (4 * states + 8 if fam == 1 else 0)) + 9] ^= 0x40
    src = z["in/multi"]
    parts, items, pos = [], [], 0
    for k, st in enumerate(streams):
        pad = (-pos) % 16
        parts.append(np.zeros(pad, np.uint8)); pos += pad
        items.append((pos, st.size, k * (src.size + 3), src.size))
        parts.append(st); pos += st.size
    out_base = np.full(6 * (src.size + 3) + 64, 0xCC, np.uint8)
    pkg.set_option("batch_group_mb", 1)
    ok, lengths = pkg.decode_batch(fam, states, bits, np.concatenate(parts), out_base, items)
    pkg.set_option("batch_group_mb", 0)
    good = ok == 5 and lengths[2] == 0 and all(np.array_equal(out_base[k * (src.size + 3): k * (src.size + 3) + src.size], src) for k in (0, 1, 3, 4, 5))
    bad += not good
    print("batch", fam, states, bits, "ok" if good else "MISMATCH", ok)
sys.exit(1 if bad else 0)
