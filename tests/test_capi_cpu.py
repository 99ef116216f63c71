"""CPU: the C-ABI library loads, exports what include/hsrans_b200.h declares, and its host logic is right.
No kernel is launched here (there is no GPU in the build container)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import checkers as ck
from conftest import ROOT, golden_stream_cases


def test_every_declared_symbol_is_exported(pkg):
    header = open(os.path.join(ROOT, "include", "hsrans_b200.h")).read()
    declared = set(re.findall(r"\b(hsr_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 25
    lib = C.CDLL(pkg.lib_path())
    missing = [name for name in sorted(declared) if not hasattr(lib, name)]
    assert not missing, missing
    from hypersonic_rans_b200 import capi
    assert declared == set(capi.SIGNATURES), declared ^ set(capi.SIGNATURES)


def test_capacity_matches_reference_formula(pkg):
    for states in (32, 64):
        for n in (0, 1, 4099, 100_000_000):
            for fam in (0, 1, 2):
                assert pkg.capacity(fam, states, n) == ck.oracle_capacity(fam, states, n)
                if ck.have_ref():
                    assert pkg.capacity(fam, states, n) == ck.ref().hsref_capacity(fam, states, n)


def test_mt_index_agrees_with_oracle_walk(pkg, golden):
    seen = 0
    for name, fam, states, bits, stream, ret, data in golden_stream_cases(golden):
        if fam != ck.MT or not ret:
            continue
        blocks = pkg.mt_index(states, stream)
        walk = ck.oracle_mt_walk(states, stream)
        assert walk is not None
        # fills may be split into several units; compare coverage and the coded blocks one by one
        coded = [b for b in blocks if b.kind == 0]
        coded_ref = [w for w in walk if (w[3] & 1) == 0]
        assert len(coded) == len(coded_ref)
        for b, w in zip(coded, coded_ref):
            assert b.inOffset == w[0] + 16 and b.outOffset == w[1]
            assert b.count - b.tail == w[2]
        pos = 0
        for b in blocks:
            assert b.outOffset == pos
            pos += b.count
        assert pos == data.size
        assert blocks[-1].inEnd <= stream.size
        seen += 1
    assert seen >= 70


def test_mt_index_multi_block_and_partition(pkg, golden):
    stream = golden["stream/multi/2/64/15"]
    blocks = pkg.mt_index(64, stream)
    assert len(blocks) >= 3
    n = golden["in/multi"].size
    assert blocks[-1].tail == n % 64
    for parts in (1, 2, 3, 8):
        first = pkg.mt_partition(blocks, parts)
        assert first[0] == 0 and first[-1] == len(blocks) and all(a <= b for a, b in zip(first, first[1:]))
    runs = golden["stream/runs/2/64/11"]
    rb = pkg.mt_index(64, runs)
    fills = [b for b in rb if b.kind == 1]
    assert fills and all(b.symbol == 0x41 for b in fills) and sum(b.count for b in fills) >= 100_000


def test_mt_index_rejects_malformed_chains(pkg, golden):
    stream = golden["stream/multi/2/32/12"].copy()
    with pytest.raises(pkg.HsrError):
        pkg.mt_index(32, stream[:200])
    bad = stream.copy()
    bad[24:32] = 0xFF  # skip offset far past the end
    with pytest.raises(pkg.HsrError):
        pkg.mt_index(32, bad)


def test_decode_argument_errors_return_zero_without_touching_the_gpu(pkg, golden):
    stream = golden["stream/small/0/32/11"]
    n = golden["in/small"].size
    assert pkg.decode(0, 32, 11, stream[:64], n)[0] == 0          # shorter than the fixed header
    assert "header" in pkg.last_error()
    assert pkg.decode(0, 32, 11, stream, n - 1)[0] == 0           # outCapacity too small
    assert pkg.decode(0, 48, 11, stream, n)[0] == 0               # unsupported state count
    assert pkg.decode(0, 32, 9, stream, n)[0] == 0                # unsupported bits
    assert pkg.decode(7, 32, 11, stream, n)[0] == 0               # unknown family
    short = stream.copy(); short[8:16] = np.frombuffer(np.uint64(stream.size + 9).tobytes(), np.uint8)
    assert pkg.decode(0, 32, 11, short, n)[0] == 0                # compressed length field > inLength


def test_no_cpu_fallback(pkg, golden):
    if pkg.device_count() > 0:
        pytest.skip("a GPU is present; the no-device behaviour is checked in the build container")
    stream = golden["stream/small/2/64/12"]
    n, _ = pkg.decode(2, 64, 12, stream, golden["in/small"].size)
    assert n == 0 and pkg.last_error() != ""
    with pytest.raises(pkg.HsrError):
        pkg.PreparedStream.upload(2, 64, 12, stream)


def test_synth_is_deterministic_and_zipf_shaped(pkg):
    a = pkg.synth_zipf(300_000, 1.0, seed=42, segment_bytes=65536)
    b = pkg.synth_zipf(300_000, 1.0, seed=42, segment_bytes=65536)
    assert np.array_equal(a, b)
    assert not np.array_equal(a, pkg.synth_zipf(300_000, 1.0, seed=43, segment_bytes=65536))
    for seg in range(4):
        p = np.bincount(a[seg * 65536:(seg + 1) * 65536], minlength=256) / 65536.0
        p = p[p > 0]
        assert 5.9 < -(p * np.log2(p)).sum() < 6.5  # H(Zipf s=1) = 6.22 bits/byte
    iid = pkg.synth_zipf(300_000, 1.0, seed=42, segment_bytes=0)
    top = np.bincount(iid, minlength=256).argmax()
    assert np.bincount(iid[:65536], minlength=256).argmax() == top == np.bincount(iid[-65536:], minlength=256).argmax()


def test_codec_registry_mirrors_reference_rows(pkg):
    names = {c.name for c in pkg.CODECS}
    assert "rANS32x64 16w 12 (raw)" in names and "rANS32x32 16w 10" in names and "rANS32x64 16w 15 mt" in names
    assert "rANS32x16 16w 13 (raw)" in names and "rANS32x32 32blk 16w 15 (raw)" in names  # main.cpp:216-228
    assert len(pkg.CODECS) == 48
    c = pkg.find_codec(pkg.HSR_MT, 64, 15)
    assert c.symbol == "mt_rANS32x64_16w_decode_15"
    assert pkg.find_codec(pkg.HSR_RAW32BLK, 32, 11).symbol == "rANS32x32_32blk_16w_decode_scalar_11"


def test_capacity_of_the_16_state_and_32blk_layouts(pkg):
    for n in (0, 1, 100, 12345, 100_000_000):
        assert pkg.capacity(pkg.HSR_RAW, 16, n) == ck.oracle_capacity(ck.RAW, 16, n) == n + 16 + 512 + 64 + 16
        assert pkg.capacity(pkg.HSR_RAW32BLK, 32, n) == ck.oracle_capacity(ck.RAW32BLK, 32, n) == n + 32 + 512 + 256 + 16
