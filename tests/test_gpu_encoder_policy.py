"""GPU (-m gpu): the device block-split policy (§8f rank 2, hsr_encode_mt_policy). Parity bar for a producer whose
splits need not equal the reference's (any split decodes): every reference decoder, the oracle and the
CUDA decoder reproduce the input byte for byte; on top of that the policy's effects are checked — stationary data is
merged into larger blocks, drifting data is not, single-symbol stretches become 8-byte run blocks."""
import numpy as np
import pytest

import checkers as ck

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu(pkg):
    if pkg.device_count() < 1:
        pytest.fail("no CUDA device: the -m gpu tests must run on the B200 box")
    return pkg


def _roundtrip(gpu, states, bits, data, max_block=0, ref_legs=True):
    stream = gpu.encode_mt_policy(states, bits, data, max_block)
    n = data.size
    got_n, got = gpu.decode(ck.MT, states, bits, stream, n)
    assert got_n == n and np.array_equal(got[:n], data), ("cuda decoder", states, bits, n, gpu.last_error())
    on, oo = ck.oracle_decode(ck.MT, states, bits, stream, n)
    assert on == n and np.array_equal(oo[:n], data), ("oracle", states, bits, n)
    if ck.have_ref() and ref_legs:
        for impl in (ck.IMPL_SCALAR, ck.IMPL_POOL):
            rn, ro = ck.ref_decode(ck.MT, states, bits, stream, n, impl)
            assert rn == n and np.array_equal(ro[:n], data), ("reference decoder", impl, states, bits, n)
    return stream


def _blocks(gpu, states, stream):
    return gpu.mt_index(states, stream)


def test_policy_streams_decode_with_every_decoder(gpu):
    seed = 900
    for states in (32, 64):
        for bits in range(10, 16):
            for n in (64, 65, 4099, 65536, 65537, 65536 + 63, 200_001, 1_000_031):
                seed += 1
                if (seed + bits) % 3 and n > 70_000:
                    continue
                data = gpu.synth_zipf(n, 1.0, seed=seed, segment_bytes=65536 if seed % 2 else 0)
                _roundtrip(gpu, states, bits, data, max_block=(seed % 3) * 131072)


def test_stationary_data_merges_and_drifting_data_does_not(gpu):
    n = 4 * 1024 * 1024 + 12345
    iid = gpu.synth_zipf(n, 1.0, seed=31, segment_bytes=0)          # one distribution for the whole buffer
    drift = gpu.synth_zipf(n, 1.0, seed=31, segment_bytes=65536)    # a new symbol permutation every 64 KiB
    for states, bits in ((64, 15), (32, 12)):
        fixed = gpu.encode_mt(states, bits, iid, 0)
        merged = _roundtrip(gpu, states, bits, iid, max_block=1 << 20)
        nb_fixed, nb_merged = len(_blocks(gpu, states, fixed)), len(_blocks(gpu, states, merged))
        assert nb_fixed == 65 and 4 <= nb_merged <= 8, (nb_fixed, nb_merged)     # 1 MiB chunks, everything inside merges
        assert merged.size < fixed.size                                           # ~60 headers saved
        sizes = sorted({int(b.count) for b in _blocks(gpu, states, merged)})
        assert sizes[-1] == 1 << 20
        kept = _roundtrip(gpu, states, bits, drift, max_block=1 << 20)
        assert len(_blocks(gpu, states, kept)) >= 60                              # nothing worth sharing a histogram


def test_single_symbol_stretches_become_run_blocks(gpu):
    data = gpu.synth_zipf(3_000_000, 1.2, seed=8, segment_bytes=65536)
    data[10 * 65536 + 5: 30 * 65536 + 77] = 0x41      # covers segments 11..29 entirely
    data[40 * 65536: 41 * 65536] = 0x00
    stream = _roundtrip(gpu, 64, 13, data, max_block=8 * 65536)
    blocks = _blocks(gpu, 64, stream)
    fills = [b for b in blocks if b.kind == 1]
    assert sum(int(b.count) for b in fills if b.symbol == 0x41) == 19 * 65536
    assert sum(int(b.count) for b in fills if b.symbol == 0x00) == 65536
    fixed = gpu.encode_mt(64, 13, data, 0)
    assert stream.size < fixed.size - 19 * 700        # 19 + 1 segments cost 8 bytes per run block instead of a header each
    # an all-constant input is one run block (three, in 1 MiB chunks) and decodes with the CUDA decoder and the oracle;
    # the reference rejects such tiny streams with its own minimum-length check (src/mt_rANS32x64_16w_decode.cpp:21-22)
    const = np.full(3 * 1024 * 1024, 0x7F, np.uint8)
    s2 = gpu.encode_mt_policy(32, 11, const, 1 << 20)
    assert s2.size == 16 + 3 * 8
    n2, out2 = ck.oracle_decode(ck.MT, 32, 11, s2, const.size)
    g2, gout2 = gpu.decode(ck.MT, 32, 11, s2, const.size)
    assert n2 == g2
    if g2:
        assert np.array_equal(gout2[:g2], const)


def test_policy_device_pointer_form_and_limits(gpu):
    import torch
    n = 5_000_000
    data = gpu.synth_zipf(n, 1.0, seed=77, segment_bytes=0)
    d_in = torch.from_numpy(data).cuda()
    bound = gpu.encode_mt_bound(64, n)
    d_out = torch.empty(bound, dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    comp = gpu.encode_mt_policy_device(64, 15, d_in.data_ptr(), n, d_out.data_ptr(), bound, 0, st)
    assert 0 < comp <= bound
    stream = d_out[:comp].cpu().numpy()
    got_n, got = gpu.decode(ck.MT, 64, 15, stream, n)
    assert got_n == n and np.array_equal(got[:n], data)
    assert max(int(b.count) for b in gpu.mt_index(64, stream)) <= 262144 + 64
    # bad arguments
    assert gpu.encode_mt_policy_device(64, 15, d_in.data_ptr(), n, d_out.data_ptr(), bound, 65536 + 64, st) == 0   # not a multiple of 64 KiB
    assert gpu.encode_mt_policy_device(64, 15, d_in.data_ptr(), n, d_out.data_ptr(), bound, 1 << 26, st) == 0      # above the reference's MaxBlockSize
    assert gpu.encode_mt_policy_device(48, 15, d_in.data_ptr(), n, d_out.data_ptr(), bound, 0, st) == 0
    assert gpu.encode_mt_policy_device(64, 15, d_in.data_ptr(), n, d_out.data_ptr(), 1000, 0, st) == 0             # output too small


def test_ragged_run_at_the_end_and_odd_lengths(gpu):
    for states, bits, n in ((64, 14, 323_457), (32, 10, 65536 * 3 + 1), (64, 15, 65536 + 65), (32, 12, 65536 * 2 + 31)):
        data = gpu.synth_zipf(n, 1.0, seed=n % 97, segment_bytes=65536)
        data[n // 3:] = 0x2A          # a run that reaches the (unaligned) end of the input
        stream = _roundtrip(gpu, states, bits, data, max_block=4 * 65536)
        fills = [b for b in _blocks(gpu, states, stream) if b.kind == 1]
        if n - n // 3 >= 2 * 65536:
            assert fills and sum(int(b.count) for b in fills) >= ((n - n // 3) // 65536 - 1) * 65536
        data2 = gpu.synth_zipf(n, 1.0, seed=5, segment_bytes=0)
        data2[: n // 2] = 0x00        # and one at the beginning
        _roundtrip(gpu, states, bits, data2, max_block=8 * 65536)
