"""Generates tests/golden/golden_rank4.npz from the UNMODIFIED reference (oracle/_ref/libhsrans_ref.so): the two
remaining 32-bit-state / 16-bit-word layouts of its registry,

    family 0, 16 states   rANS32x16_16w         (src/rANS32x16_16w.cpp)
    family 3, 32 states   rANS32x32_32blk_16w   (src/rans32x32_32blk_16w.cpp)

Same key layout as golden.npz (stream/<input>/<family>/<states>/<bits>, ret/..., in/<input>); inputs are re-generated
with the seeds of make_golden.py so both files agree on them. Run in the build container:
    python tests/golden/make_golden_rank4.py
Every stream is decoded by the reference's scalar AND fastest AVX2 decoder before it is stored (the AVX2 32blk
decoders work on two rows per step and crash below 64 symbols, so they are skipped there).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import checkers as ck  # noqa: E402
from make_golden import zipf_bytes  # noqa: E402


def main():
    inputs = {
        "tiny16": zipf_bytes(16, 1.0, 11),
        "tiny33": zipf_bytes(33, 1.0, 12),
        "tiny64": zipf_bytes(64, 1.0, 1),
        "tiny65": zipf_bytes(65, 1.0, 2),
        "tiny127": zipf_bytes(127, 1.0, 3),
        "small": zipf_bytes(4099, 1.0, 4),
        "flat": np.random.default_rng(5).integers(0, 256, 3001).astype(np.uint8),
        "skew": zipf_bytes(5000, 3.0, 6),
        "multi": zipf_bytes(262144 + 37, 1.0, 7, segment=65536),
        "const": np.full(5000, 0x55, np.uint8),
    }
    out = {f"in/{k}": v for k, v in inputs.items()}
    cases = []
    for name in ("tiny16", "tiny33", "tiny64", "tiny65", "tiny127", "small", "flat", "skew"):
        for fam, states in ((ck.RAW, 16), (ck.RAW32BLK, 32)):
            if inputs[name].size < states:
                continue
            for bits in range(10, 16):
                cases.append((name, fam, states, bits))
    cases += [("multi", ck.RAW, 16, 12), ("multi", ck.RAW, 16, 15), ("multi", ck.RAW32BLK, 32, 11), ("multi", ck.RAW32BLK, 32, 15),
              ("const", ck.RAW, 16, 11), ("const", ck.RAW32BLK, 32, 14)]
    for name, fam, states, bits in cases:
        data = inputs[name]
        stream = ck.ref_encode(fam, states, bits, data)
        n, dec = ck.ref_decode(fam, states, bits, stream, data.size)
        assert n == data.size and np.array_equal(dec[:n], data), (name, fam, states, bits)
        if not (fam == ck.RAW32BLK and data.size < 64) and not (name == "const" and bits == 12):
            n2, dec2 = ck.ref_decode(fam, states, bits, stream, data.size, ck.IMPL_AVX2)
            assert n2 == n and np.array_equal(dec2[:n], data), (name, fam, states, bits, "avx2")
        out[f"stream/{name}/{fam}/{states}/{bits}"] = stream
        out[f"ret/{name}/{fam}/{states}/{bits}"] = np.array([n], np.uint64)
    path = os.path.join(HERE, "golden_rank4.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {len(out)} arrays, {os.path.getsize(path) / 1e6:.2f} MB, {len(cases)} stream cases")


if __name__ == "__main__":
    main()
