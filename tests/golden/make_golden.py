"""Generates tests/golden/golden.npz from the UNMODIFIED reference (oracle/_ref/libhsrans_ref.so).

Run in the build container (needs /root/reference compiled by `make -C oracle ref`):
    python tests/golden/make_golden.py
The reference ships no golden vectors of its own: its only check is round-trip identity. These
fixtures pin (a) the three stream formats byte for byte as its encoders write them, (b) its decoders' output and
(c) make_hist / normalize_hist, so the oracle and the CUDA path can be checked where /root/reference is absent.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import checkers as ck  # noqa: E402


def zipf_bytes(n, s, seed, segment=0):
    rng = np.random.default_rng(seed)
    p = 1.0 / np.arange(1, 257) ** s
    p /= p.sum()
    ranks = rng.choice(256, n, p=p)
    if segment == 0:
        return rng.permutation(256).astype(np.uint8)[ranks]
    out = np.empty(n, np.uint8)
    for o in range(0, n, segment):
        out[o:o + segment] = rng.permutation(256).astype(np.uint8)[ranks[o:o + segment]]
    return out


def main():
    inputs = {
        "tiny64": zipf_bytes(64, 1.0, 1),
        "tiny65": zipf_bytes(65, 1.0, 2),
        "tiny127": zipf_bytes(127, 1.0, 3),
        "small": zipf_bytes(4099, 1.0, 4),
        "flat": np.random.default_rng(5).integers(0, 256, 3001).astype(np.uint8),
        "skew": zipf_bytes(5000, 3.0, 6),
        "multi": zipf_bytes(262144 + 37, 1.0, 7, segment=65536),
    }
    runs = zipf_bytes(300000, 1.2, 8, segment=65536)
    runs[70000:230000] = 0x41  # a long single-symbol run in the middle -> memset blocks in block_/mt_
    runs[-20000:] = 0x7A       # and one reaching the (unaligned) end
    inputs["runs"] = runs[:299987]
    const = np.full(5000, 0x55, np.uint8)
    inputs["const"] = const

    out = {f"in/{k}": v for k, v in inputs.items()}
    cases = []
    for name in ("tiny64", "tiny65", "tiny127", "small", "flat", "skew"):
        for fam in (ck.RAW, ck.BLOCK, ck.MT):
            for states in (32, 64):
                if inputs[name].size < states:
                    continue
                for bits in range(10, 16):
                    cases.append((name, fam, states, bits))
    for fam, states, bits in ((ck.MT, 64, 15), (ck.MT, 32, 12), (ck.BLOCK, 32, 10), (ck.BLOCK, 64, 13), (ck.RAW, 64, 12)):
        cases.append(("multi", fam, states, bits))
    for fam, states, bits in ((ck.MT, 64, 11), (ck.MT, 32, 14), (ck.BLOCK, 32, 12), (ck.BLOCK, 64, 15)):
        cases.append(("runs", fam, states, bits))
    for fam, states, bits in ((ck.RAW, 32, 11), (ck.RAW, 64, 15), (ck.BLOCK, 32, 10), (ck.MT, 64, 13)):
        cases.append(("const", fam, states, bits))

    for name, fam, states, bits in cases:
        data = inputs[name]
        stream = ck.ref_encode(fam, states, bits, data)
        n, dec = ck.ref_decode(fam, states, bits, stream, data.size)
        if name == "const" and fam != ck.RAW:
            # quirk: an all-constant input encodes to a stream shorter than the decoders' own minimum-length
            # check (src/block_rANS32x32_16w_decode.cpp:21-22), so the reference rejects its own output with 0
            assert n == 0, (name, fam, states, bits, n)
        else:
            assert n == data.size and np.array_equal(dec[:n], data), (name, fam, states, bits)
        out[f"stream/{name}/{fam}/{states}/{bits}"] = stream
        out[f"ret/{name}/{fam}/{states}/{bits}"] = np.array([n], np.uint64)

    for name in ("small", "flat", "skew", "multi", "runs"):
        for bits in range(10, 16):
            cnt, cum = ck.ref_make_hist(inputs[name], bits)
            out[f"hist/{name}/{bits}"] = np.stack([cnt, cum])
    # normalize_hist with dataBytes != sum(hist), as the block_/mt_ encoders call it
    # (src/block_rANS32x32_16w_encode.cpp:91,201,331)
    h = np.bincount(inputs["multi"][:65536], minlength=256).astype(np.uint32)
    h2 = h.copy()
    extra = int((h2 == 0).sum())
    h2[h2 == 0] = 1
    for bits in range(10, 16):
        cnt, cum = ck.ref_normalize_hist(h2, 65536 + extra, bits)
        out[f"norm/safe/{bits}"] = np.stack([cnt, cum])
    out["norm/safe/hist"] = h2

    path = os.path.join(HERE, "golden.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {len(out)} arrays, {os.path.getsize(path) / 1e6:.2f} MB, {len(cases)} stream cases")


if __name__ == "__main__":
    main()
