"""GPU (-m gpu): the device mt_ encoder (§8f rank 1). Parity bar for a producer whose block policy differs from the
reference's: every reference decoder, the oracle and the CUDA decoder must reproduce the input byte for byte from
its streams, and where the policies coincide (inputs of at most one block) the stream itself must be identical."""
import numpy as np
import pytest

import checkers as ck

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu(pkg):
    if pkg.device_count() < 1:
        pytest.fail("no CUDA device: the -m gpu tests must run on the B200 box")
    return pkg


def _roundtrip(gpu, states, bits, data, block_size=0):
    stream = gpu.encode_mt(states, bits, data, block_size)
    n = data.size
    got_n, got = gpu.decode(ck.MT, states, bits, stream, n)
    assert got_n == n and np.array_equal(got[:n], data), ("cuda decoder", states, bits, n, gpu.last_error())
    on, oo = ck.oracle_decode(ck.MT, states, bits, stream, n)
    assert on == n and np.array_equal(oo[:n], data), ("oracle", states, bits, n)
    if ck.have_ref():
        for impl in (ck.IMPL_SCALAR, ck.IMPL_POOL):
            rn, ro = ck.ref_decode(ck.MT, states, bits, stream, n, impl)
            assert rn == n and np.array_equal(ro[:n], data), ("reference decoder", impl, states, bits, n)
    return stream


def test_streams_decode_with_every_decoder(gpu):
    seed = 500
    for states in (32, 64):
        for bits in range(10, 16):
            for n in (64, 65, 127, 4099, 65536, 65537, 65536 + 63, 200_001, 1_000_031):
                seed += 1
                if (seed + bits) % 3 and n > 70_000:
                    continue  # keep the reference-decoder legs short
                data = gpu.synth_zipf(n, 1.0, seed=seed, segment_bytes=65536 if seed % 2 else 0)
                _roundtrip(gpu, states, bits, data)


def test_block_sizes_and_entropies(gpu):
    for s in (0.0, 0.5, 2.0, 3.0):
        data = gpu.synth_zipf(700_003, s, seed=9, segment_bytes=65536)
        for states, bits, bs in ((64, 15, 0), (64, 12, 32768), (32, 10, 131072), (32, 13, 64 * 1000)):
            stream = _roundtrip(gpu, states, bits, data, bs)
            blocks = gpu.mt_index(states, stream)
            want = (data.size + (bs or 65536) - 1) // (bs or 65536)
            assert len(blocks) in (want, want - 1)
    const = np.full(100_000, 0x33, np.uint8)
    _roundtrip(gpu, 64, 12, const)
    runs = gpu.synth_zipf(300_000, 1.2, seed=8, segment_bytes=65536)
    runs[70_000:230_000] = 0x41
    _roundtrip(gpu, 32, 15, runs)


@pytest.mark.skipif(not ck.have_ref(), reason="needs the reference encoder (oracle/_ref)")
def test_single_block_streams_are_byte_identical_to_the_reference(gpu):
    seed = 900
    for states in (32, 64):
        for bits in range(10, 16):
            for n in (64, 100, 4099, 65535, 65536):
                seed += 1
                data = gpu.synth_zipf(n, 1.0 if seed % 2 else 0.3, seed=seed, segment_bytes=0)
                mine = gpu.encode_mt(states, bits, data)
                ref = ck.ref_encode(ck.MT, states, bits, data)
                assert mine.size == ref.size and np.array_equal(mine, ref), (states, bits, n)


def test_block_histograms_are_the_reference_normalisation(gpu):
    data = gpu.synth_zipf(65536 * 5 + 1000, 1.0, seed=77, segment_bytes=65536)
    for states, bits in ((64, 15), (32, 11)):
        stream = gpu.encode_mt(states, bits, data)
        blocks = gpu.mt_index(states, stream)
        for k, b in enumerate(blocks):
            counts = stream[b.inOffset + 4 * states: b.inOffset + 4 * states + 512].view(np.uint16)
            want, _ = ck.oracle_make_hist(data[b.outOffset: b.outOffset + b.count], bits)
            assert np.array_equal(counts, want), (states, bits, k)


def test_device_pointer_encoder_and_capacity(gpu):
    import torch
    data = gpu.synth_zipf(3_000_000, 1.0, seed=4, segment_bytes=65536)
    d_in = torch.from_numpy(data).cuda()
    bound = gpu.encode_mt_bound(64, data.size)
    d_out = torch.empty(bound, dtype=torch.uint8, device="cuda")
    n = gpu.encode_mt_device(64, 15, d_in.data_ptr(), data.size, d_out.data_ptr(), bound, 0, torch.cuda.current_stream().cuda_stream)
    assert 0 < n <= bound
    stream = d_out[:n].cpu().numpy()
    got_n, got = gpu.decode(ck.MT, 64, 15, stream, data.size)
    assert got_n == data.size and np.array_equal(got[: data.size], data)
    assert gpu.encode_mt_device(64, 15, d_in.data_ptr(), data.size, d_out.data_ptr(), 1000, 0, 0) == 0   # capacity too small
    assert gpu.encode_mt_bound(64, 10) == 0                                                               # shorter than one row
    assert gpu.encode_mt_bound(64, 1000, 100) == 0                                                        # block size not a multiple of N


def test_full_size_100mb(gpu):
    n = 100_000_000
    data = gpu.synth_zipf(n, 1.0, seed=42, segment_bytes=0)   # stationary: the reference encoder would emit 5 blocks
    stream = gpu.encode_mt(64, 15, data)
    assert len(gpu.mt_index(64, stream)) == (n + 65535) // 65536
    got_n, got = gpu.decode(ck.MT, 64, 15, stream, n)
    assert got_n == n and np.array_equal(got[:n], data)
    if ck.have_ref():
        rn, ro = ck.ref_decode(ck.MT, 64, 15, stream, n, ck.IMPL_POOL)
        assert rn == n and np.array_equal(ro[:n], data)


def test_full_size_1gb_config4(gpu):
    """BASELINE config 4 size: 1,000,000,000 bytes through the device encoder and the CUDA decoder; the reference's
    thread-pool decoder must agree on the same stream."""
    n = 1_000_000_000
    data = gpu.synth_zipf(n, 1.0, seed=42, segment_bytes=65536)
    stream = gpu.encode_mt(64, 15, data)
    got_n, got = gpu.decode(ck.MT, 64, 15, stream, n)
    assert got_n == n and np.array_equal(got[:n], data)
    del got
    if ck.have_ref():
        rn, ro = ck.ref_decode(ck.MT, 64, 15, stream, n, ck.IMPL_POOL)
        assert rn == n and np.array_equal(ro[:n], data)


def test_encoder_output_is_pinned(gpu):
    """Multi-block device-encoded streams are deterministic: their SHA-256 is pinned in tests/golden/encoder_hashes.json
    (generated on a B200 with the same deterministic inputs; single-block streams are pinned against the reference
    encoder itself above)."""
    import hashlib, json, os
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "encoder_hashes.json")
    got = {}
    for states, bits, n, s, seg, bs in ((64, 15, 1_000_031, 1.0, 65536, 0), (32, 12, 700_001, 0.5, 0, 32768),
                                        (64, 10, 300_000, 2.0, 4096, 65536), (32, 14, 262_144 + 37, 1.3, 65536, 0)):
        data = gpu.synth_zipf(n, s, seed=1234, segment_bytes=seg)
        stream = gpu.encode_mt(states, bits, data, bs)
        got[f"{states}/{bits}/{n}/{s}/{seg}/{bs}"] = hashlib.sha256(stream.tobytes()).hexdigest()
    if os.environ.get("HSR_WRITE_ENCODER_HASHES") == "1":
        json.dump(got, open(path, "w"), indent=1, sort_keys=True)
    want = json.load(open(path))
    assert got == want
