"""GPU (-m gpu): rANS32x16_16w (family 0, 16 states) and rANS32x32_32blk_16w (family 3) through the C-ABI, bit-exact
against the reference's own streams (tests/golden/golden_rank4.npz, or fresh ones from oracle/_ref) and the oracle."""
import numpy as np
import pytest

import checkers as ck
from conftest import golden_stream_cases

pytestmark = pytest.mark.gpu

CODECS = ((ck.RAW, 16), (ck.RAW32BLK, 32))


@pytest.fixture(scope="module")
def gpu(pkg):
    if pkg.device_count() < 1:
        pytest.fail("no CUDA device: the -m gpu tests must run on the B200 box")
    return pkg


def _check(gpu, fam, states, bits, stream, data, label=""):
    n, out = gpu.decode(fam, states, bits, stream, data.size)
    assert n == data.size, (label, fam, states, bits, n, gpu.last_error())
    if not np.array_equal(out[:n], data):
        bad = np.nonzero(out[:n] != data)[0]
        pytest.fail(f"{label} fam {fam} N {states} bits {bits}: {bad.size} wrong bytes, first at {bad[:8]}")
    assert out[n:].size == 0 or np.all(out[n:] == 0xCC)


def test_rank4_golden_streams_bit_exact(gpu, golden_rank4):
    cases = golden_stream_cases(golden_rank4)
    assert len(cases) >= 90
    for name, fam, states, bits, stream, ret, data in cases:
        _check(gpu, fam, states, bits, stream, data, name)
        # the reference's error returns (src/rANS32x16_16w.cpp:165-181, src/rans32x32_32blk_16w.cpp:185-201)
        assert gpu.decode(fam, states, bits, stream, data.size - 1)[0] == 0
        assert gpu.decode(fam, states, bits, stream[:-1], data.size)[0] == 0


@pytest.mark.skipif(not ck.have_ref(), reason="stream production for these codecs needs oracle/_ref")
def test_rank4_fresh_streams_all_bits_ragged_lengths_and_entropies(gpu):
    lengths = [32, 33, 47, 64, 65, 127, 4097, 70_001, 1_000_031]
    seed = 500
    for fam, states in CODECS:
        for bits in range(10, 16):
            for s in (0.0, 1.0, 3.0):
                seed += 1
                n = lengths[seed % len(lengths)]
                data = gpu.synth_zipf(n, s, seed=seed, segment_bytes=65536 if seed % 2 else 0)
                try:
                    stream = ck.ref_encode(fam, states, bits, data)
                except ck.RefEncoderOverflow:
                    continue
                _check(gpu, fam, states, bits, stream, data, f"n={n} s={s}")
                on, oo = ck.oracle_decode(fam, states, bits, stream, n)
                assert on == n and np.array_equal(oo[:n], data)


def test_rank4_malformed_streams_return_zero(gpu, golden_rank4):
    n = golden_rank4["in/small"].size
    for fam, states, bits in ((ck.RAW, 16, 12), (ck.RAW32BLK, 32, 12)):
        stream = golden_rank4[f"stream/small/{fam}/{states}/{bits}"]
        bad = stream.copy(); bad[16 + 9] ^= 0x10  # histogram no longer sums to 2^12
        assert gpu.decode(fam, states, bits, bad, n)[0] == 0
        assert ck.oracle_decode(fam, states, bits, bad, n)[0] == 0
        assert gpu.decode(fam, states, bits, stream[:200], n)[0] == 0  # shorter than the fixed header
        _check(gpu, fam, states, bits, stream, golden_rank4["in/small"], "after errors")
    # 32blk: a sub-stream size that points past the end of the stream
    stream = golden_rank4["stream/small/3/32/12"]
    bad = stream.copy()
    sizes = 16 + 512 + 128
    bad[sizes:sizes + 4] = np.frombuffer(np.uint32(stream.size).tobytes(), np.uint8)
    assert gpu.decode(ck.RAW32BLK, 32, 12, bad, n)[0] == 0
    # random corruption of the word area never crashes and never writes past n
    rng = np.random.default_rng(9)
    for fam, states, bits in ((ck.RAW, 16, 15), (ck.RAW32BLK, 32, 15)):
        stream = golden_rank4[f"stream/multi/{fam}/{states}/{bits}"]
        data = golden_rank4["in/multi"]
        for _ in range(6):
            bad = stream.copy()
            pos = rng.integers(16 + 512 + 4 * 63, stream.size, 40)
            bad[pos] ^= rng.integers(1, 256, 40).astype(np.uint8)
            got, out = gpu.decode(fam, states, bits, bad, data.size)
            assert got in (0, data.size)
            assert np.all(out[data.size:] == 0xCC)


def test_rank4_batches_and_prepared_streams(gpu, golden_rank4):
    import torch
    for fam, states, bits in ((ck.RAW, 16, 13), (ck.RAW32BLK, 32, 13)):
        names = ["small", "tiny65", "skew", "flat", "tiny127", "tiny33"]
        streams = [golden_rank4[f"stream/{nm}/{fam}/{states}/{bits}"] for nm in names] * 3
        datas = [golden_rank4[f"in/{nm}"] for nm in names] * 3
        bad_index = 4
        streams[bad_index] = streams[bad_index].copy()
        streams[bad_index][16 + 11] ^= 0x40
        items, parts, pos_in, pos_out = [], [], 0, 0
        for s, d in zip(streams, datas):
            pad = (-pos_in) % 16
            parts.append(np.zeros(pad, np.uint8)); pos_in += pad
            items.append((pos_in, s.size, pos_out, d.size))
            parts.append(s); pos_in += s.size
            pos_out += d.size + 5
        in_base = np.concatenate(parts)
        out_base = np.full(pos_out + 64, 0xCC, np.uint8)
        ok, lengths = gpu.decode_batch(fam, states, bits, in_base, out_base, items)
        assert ok == len(items) - 1
        for k, ((io, il, oo, oc), d) in enumerate(zip(items, datas)):
            if k == bad_index:
                assert lengths[k] == 0 and np.all(out_base[oo:oo + oc] == 0xCC)
                continue
            assert lengths[k] == d.size and np.array_equal(out_base[oo:oo + d.size], d), (fam, k)
            assert np.all(out_base[oo + d.size: oo + d.size + 5] == 0xCC)
        # device-resident API
        stream, data = golden_rank4[f"stream/small/{fam}/{states}/{bits}"], golden_rank4["in/small"]
        ps = gpu.PreparedStream.upload(fam, states, bits, stream)
        out = torch.full((data.size + 64,), 0xCC, dtype=torch.uint8, device="cuda")
        ps.decode_async(out.data_ptr(), data.size, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        assert ps.status() == 0 and np.array_equal(out[:data.size].cpu().numpy(), data)
        assert bool((out[data.size:] == 0xCC).all())
        ps.free()
        d_in = torch.from_numpy(np.concatenate([stream, np.zeros(16, np.uint8)])).cuda()
        ps = gpu.PreparedStream.from_device(fam, states, bits, d_in.data_ptr(), stream.size)
        out.fill_(0xCC)
        ps.decode_async(out.data_ptr(), data.size, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        assert ps.status() == 0 and np.array_equal(out[:data.size].cpu().numpy(), data)
        ps.free()


@pytest.mark.skipif(not ck.have_ref(), reason="stream production for these codecs needs oracle/_ref")
def test_rank4_20mb_round_trip(gpu):
    n = 20_000_003
    data = gpu.synth_zipf(n, 1.0, seed=4242, segment_bytes=0)
    for fam, states, bits in ((ck.RAW, 16, 11), (ck.RAW32BLK, 32, 15)):
        stream = ck.ref_encode(fam, states, bits, data)
        _check(gpu, fam, states, bits, stream, data, "20 MB")


@pytest.mark.skipif(not ck.have_ref(), reason="stream production for these codecs needs oracle/_ref")
def test_32blk_streams_the_reference_encoder_corrupted_decode_like_the_reference_decoder(gpu):
    """The reference's 32blk encoder overflows a state's word region on incompressible input and then fails its own
    round trip (see tests/test_oracle_golden.py); the GPU must reproduce what the reference DEcoder makes of it."""
    seen = 0
    for seed, bits, n in ((0, 10, 50_000), (2, 10, 50_000), (1, 15, 50_000), (0, 12, 50_000)):
        data = np.random.default_rng(seed).integers(0, 256, n).astype(np.uint8)
        try:
            stream = ck.ref_encode(ck.RAW32BLK, 32, bits, data)
        except ck.RefEncoderOverflow:
            continue
        rn, ro = ck.ref_decode(ck.RAW32BLK, 32, bits, stream, n)
        got_n, got = gpu.decode(ck.RAW32BLK, 32, bits, stream, n)
        assert got_n == rn == n and np.array_equal(got[:n], ro[:n]), (seed, bits, n)
        seen += not np.array_equal(ro[:n], data)
    assert seen >= 1
