"""GPU (-m gpu): the CUDA path, called through the C-ABI, against the oracle / the reference — bit-exact.

Nothing here reads /root/reference: streams come from the committed golden fixtures or from the reference build
that travels in oracle/_ref/ (falling back to the oracle's raw encoder twin when that is absent)."""
import numpy as np
import pytest

import checkers as ck
from conftest import golden_stream_cases

pytestmark = pytest.mark.gpu


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
    assert out[n:].size == 0 or np.all(out[n:] == 0xCC)  # never writes past the decoded length


@pytest.mark.parametrize("table", [0, 1, 3])
def test_golden_streams_bit_exact(gpu, golden, table):
    gpu.set_option("table", table)
    try:
        cases = golden_stream_cases(golden)
        for name, fam, states, bits, stream, ret, data in cases:
            if ret == 0:
                assert gpu.decode(fam, states, bits, stream, data.size)[0] == 0, (name, fam, states, bits)
                continue
            _check(gpu, fam, states, bits, stream, data, name)
    finally:
        gpu.set_option("table", 0)


def _zipf(pkg, n, s, seed, seg):
    return pkg.synth_zipf(n, s, seed=seed, segment_bytes=seg)


@pytest.mark.parametrize("fam", [ck.RAW, ck.BLOCK, ck.MT])
def test_fresh_streams_all_bits_and_ragged_lengths(gpu, fam):
    if fam != ck.RAW and not ck.have_ref():
        pytest.skip("block_/mt_ stream production needs oracle/_ref")
    lengths = [64, 65, 95, 127, 128, 4097, 70_001, 1_000_031]
    seed = 100
    for states in (32, 64):
        for bits in range(10, 16):
            n = lengths[(bits + states) % len(lengths)]
            for n in {n, lengths[(bits * 3 + states // 32) % len(lengths)]}:
                seed += 1
                data = _zipf(gpu, n, 1.0, seed, 65536 if seed % 2 else 0)
                stream = ck.encode(fam, states, bits, data)
                _check(gpu, fam, states, bits, stream, data, f"n={n}")
                on, oo = ck.oracle_decode(fam, states, bits, stream, n)
                assert on == n and np.array_equal(oo[:n], data)


@pytest.mark.parametrize("s", [0.0, 0.5, 1.5, 3.0])
def test_entropy_sweep(gpu, s):
    """BASELINE config 5: near-uniform to highly skewed sources, raw vs block_ (per-block normalised) histograms."""
    n = 600_011
    data = _zipf(gpu, n, s, 77, 65536)
    for fam in (ck.RAW, ck.BLOCK, ck.MT):
        if fam != ck.RAW and not ck.have_ref():
            continue
        for states, bits in ((32, 10), (64, 12), (32, 13), (64, 15)):
            stream = ck.encode(fam, states, bits, data)
            _check(gpu, fam, states, bits, stream, data, f"s={s}")


def test_single_symbol_runs_and_constant_input(gpu, golden):
    for key in ("stream/runs/2/64/11", "stream/runs/2/32/14", "stream/runs/1/32/12", "stream/runs/1/64/15"):
        _, name, fam, states, bits = key.split("/")
        _check(gpu, int(fam), int(states), int(bits), golden[key], golden["in/runs"], key)
    const = golden["in/const"]
    _check(gpu, ck.RAW, 32, 11, golden["stream/const/0/32/11"], const, "const raw 11")
    _check(gpu, ck.RAW, 64, 15, golden["stream/const/0/64/15"], const, "const raw 15")
    # freq == 2^12 does not fit the reference's packed 12-bit field (src/hist.cpp:304); ours must still decode it
    data = np.full(9000, 0x21, np.uint8)
    stream = ck.oracle_encode_raw(64, 12, data)
    _check(gpu, ck.RAW, 64, 12, stream, data, "const raw 12")


def test_malformed_streams_return_zero(gpu, golden):
    n = golden["in/multi"].size
    mt = golden["stream/multi/2/64/15"]
    bad = mt.copy(); bad[16 + 16 + 256 + 9] ^= 0x20          # first block's histogram no longer sums to 2^15
    assert gpu.decode(ck.MT, 64, 15, bad, n)[0] == 0
    assert "status" in gpu.last_error()
    blk = golden["stream/multi/1/32/10"]
    bad = blk.copy(); bad[16 + 128 + 8 + 5] ^= 0x10
    assert gpu.decode(ck.BLOCK, 32, 10, bad, n)[0] == 0
    raw = golden["stream/multi/0/64/12"]
    bad = raw.copy(); bad[16 + 3] ^= 0x01
    assert gpu.decode(ck.RAW, 64, 12, bad, n)[0] == 0
    assert gpu.decode(ck.RAW, 64, 12, raw, n - 1)[0] == 0     # outCapacity too small
    # and a good stream still decodes afterwards
    _check(gpu, ck.RAW, 64, 12, raw, golden["in/multi"], "after errors")


def test_prepared_stream_device_api_and_shards(gpu, golden):
    import torch
    data = golden["in/multi"]
    stream = golden["stream/multi/2/64/15"]
    n = data.size
    out = torch.full((n + 64,), 0xCC, dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    ps = gpu.PreparedStream.upload(ck.MT, 64, 15, stream)
    assert ps.decoded_length == n and ps.units >= 4 and ps.shard_out_bytes == n
    assert ps.decode_async(out.data_ptr(), n, st) >= 1
    torch.cuda.synchronize()
    assert ps.status() == 0
    assert np.array_equal(out[:n].cpu().numpy(), data) and bool((out[n:] == 0xCC).all())
    ps.free()
    # three shards decoded into the same device buffer tile it exactly
    out.fill_(0xCC)
    total = 0
    for r in range(3):
        sh = gpu.PreparedStream.upload(ck.MT, 64, 15, stream, shard=r, shards=3)
        total += sh.shard_out_bytes
        sh.decode_async(out.data_ptr(), n, st)
        torch.cuda.synchronize()
        assert sh.status() == 0
        # shard-local output: only this rank's bytes, based at 0
        loc = torch.full((sh.shard_out_bytes + 16,), 0xCC, dtype=torch.uint8, device="cuda")
        sh.decode_async(loc.data_ptr(), sh.shard_out_bytes, st, shard_local=True)
        torch.cuda.synchronize()
        lo = sh.shard_out_offset
        assert np.array_equal(loc[: sh.shard_out_bytes].cpu().numpy(), data[lo: lo + sh.shard_out_bytes])
        sh.free()
    assert total == n and np.array_equal(out[:n].cpu().numpy(), data)
    # stream that only exists in device memory: the chain is walked by a kernel
    dev_in = torch.from_numpy(stream.copy()).cuda()
    ds = gpu.PreparedStream.from_device(ck.MT, 64, 15, dev_in.data_ptr(), stream.size)
    host_idx = gpu.mt_index(64, stream)
    dev_idx = ds.index()
    assert len(dev_idx) == len(host_idx)
    for a, b in zip(dev_idx, host_idx):
        assert (a.inOffset, a.inEnd, a.outOffset, a.count, a.kind, a.symbol, a.tail) == \
               (b.inOffset, b.inEnd, b.outOffset, b.count, b.kind, b.symbol, b.tail)
    out.fill_(0xCC)
    ds.decode_async(out.data_ptr(), n, st)
    torch.cuda.synchronize()
    assert ds.status() == 0 and np.array_equal(out[:n].cpu().numpy(), data)
    ds.free()
    for fam, states, bits in ((ck.RAW, 64, 12), (ck.BLOCK, 32, 10)):
        s2 = golden[f"stream/multi/{fam}/{states}/{bits}"]
        p2 = gpu.PreparedStream.upload(fam, states, bits, s2)
        out.fill_(0xCC)
        p2.decode_async(out.data_ptr(), n, st)
        torch.cuda.synchronize()
        assert p2.status() == 0 and np.array_equal(out[:n].cpu().numpy(), data)
        p2.free()


@pytest.mark.parametrize("fam", [ck.RAW, ck.BLOCK, ck.MT])
def test_batch_of_independent_streams(gpu, golden, fam):
    """One launch over many streams: the only way the single-recurrence codecs fill a GPU."""
    states, bits = {ck.RAW: (64, 12), ck.BLOCK: (32, 10), ck.MT: (64, 15)}[fam]
    names = ["multi", "small", "tiny65", "skew", "flat", "tiny127"]
    if fam == ck.MT:
        names = ["multi", "small", "tiny65"]
    streams, datas = [], []
    for k in range(24):
        name = names[k % len(names)]
        key = f"stream/{name}/{fam}/{states}/{bits}"
        if key not in golden:
            continue
        streams.append(golden[key]); datas.append(golden[f"in/{name}"])
    bad_index = 3
    streams[bad_index] = streams[bad_index].copy()
    off = {ck.RAW: 16 + 9, ck.BLOCK: 16 + 4 * states + 8 + 9, ck.MT: 16 + 16 + 4 * states + 9}[fam]
    streams[bad_index][off] ^= 0x40   # break one histogram: that stream must fail alone
    items, in_parts, pos_in, pos_out = [], [], 0, 0
    for s, d in zip(streams, datas):
        pad = (-pos_in) % 16
        in_parts.append(np.zeros(pad, np.uint8)); pos_in += pad
        items.append((pos_in, s.size, pos_out, d.size))
        in_parts.append(s); pos_in += s.size
        pos_out += d.size + 7  # leave small gaps: nothing may be written there
    in_base = np.concatenate(in_parts)
    out_base = np.full(pos_out + 64, 0xCC, np.uint8)
    ok, lengths = gpu.decode_batch(fam, states, bits, in_base, out_base, items)
    assert ok == len(items) - 1
    for k, ((io, il, oo, oc), d) in enumerate(zip(items, datas)):
        if k == bad_index:
            assert lengths[k] == 0 and np.all(out_base[oo:oo + oc] == 0xCC)
            continue
        assert lengths[k] == d.size and np.array_equal(out_base[oo:oo + d.size], d), k
        assert np.all(out_base[oo + d.size: oo + d.size + 7] == 0xCC)


def test_multi_device_entry_point_on_one_gpu(gpu, golden):
    data = golden["in/multi"]
    n, out = gpu.decode_mt_multi(64, 15, golden["stream/multi/2/64/15"], data.size, devices=[0, 0])
    assert n == data.size and np.array_equal(out[:n], data)


def test_pinned_host_buffers(gpu, golden):
    data = golden["in/multi"]
    stream = golden["stream/multi/2/32/12"]
    hin, hout = gpu.host_alloc(stream.size), gpu.host_alloc(data.size)
    hin.array[:] = stream
    hout.array[:] = 0xCC
    n, out = gpu.decode(ck.MT, 32, 12, hin.array, data.size, out=hout.array)
    assert n == data.size and np.array_equal(out, data)
    hin.free(); hout.free()


def test_device_histogram_bit_exact(gpu, golden):
    for key in golden:
        if not key.startswith("hist/"):
            continue
        _, name, bits = key.split("/")
        cnt, cum = gpu.make_hist(golden[f"in/{name}"], int(bits))
        assert np.array_equal(cnt, golden[key][0]) and np.array_equal(cum, golden[key][1]), key
    data = _zipf(gpu, 3_000_001, 1.0, 5, 65536)
    for bits in (10, 12, 15):
        want = ck.oracle_make_hist(data, bits)
        got = gpu.make_hist(data, bits)
        assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])


def test_device_segment_histograms(gpu):
    import torch
    data = _zipf(gpu, 1_000_000, 1.2, 9, 65536)
    seg = 65536
    nseg = (data.size + seg - 1) // seg
    d = torch.from_numpy(data).cuda()
    counts = torch.zeros((nseg, 256), dtype=torch.int16, device="cuda")
    for bits in (10, 15):
        assert gpu.make_hist_segments_device(d.data_ptr(), data.size, seg, bits, counts.data_ptr(),
                                             torch.cuda.current_stream().cuda_stream) > 0
        torch.cuda.synchronize()
        got = counts.cpu().numpy().view(np.uint16)
        for k in range(nseg):
            want, _ = ck.oracle_make_hist(data[k * seg:(k + 1) * seg], bits)
            assert np.array_equal(got[k], want), (bits, k)


def test_observe_hist_device_counts(gpu):
    import torch
    data = _zipf(gpu, 5_000_003, 1.0, 11, 0)
    d = torch.from_numpy(data).cuda()
    hist = torch.zeros(256, dtype=torch.int32, device="cuda")
    # deliberately misaligned start
    assert gpu.observe_hist_device(d.data_ptr() + 3, data.size - 3, hist.data_ptr(), torch.cuda.current_stream().cuda_stream) > 0
    torch.cuda.synchronize()
    assert np.array_equal(hist.cpu().numpy().astype(np.int64), np.bincount(data[3:], minlength=256))


def test_histogram_counter_columns_edges(gpu):
    """The lane-private u16 counter halves are flushed every 64,512 bytes per thread: ranges long enough for several
    epochs, a constant input (one bin takes every increment of every thread), odd segment sizes with unaligned starts,
    and inputs shorter than one 16-byte vector must all count exactly (src/hist.cpp:8-14) and normalise like the
    reference (:16-215)."""
    import torch
    st = torch.cuda.current_stream().cuda_stream
    # 1. one 9.4 MB range = 2.3 epochs of a 64-thread CTA; 20,000,003 bytes -> 3 ranges, the last one short and odd
    data = _zipf(gpu, 20_000_003, 1.0, 21, 0)
    d = torch.from_numpy(data).cuda()
    seg = 9_437_187
    nseg = (data.size + seg - 1) // seg
    counts = torch.zeros((nseg, 256), dtype=torch.int16, device="cuda")
    assert gpu.make_hist_segments_device(d.data_ptr(), data.size, seg, 15, counts.data_ptr(), st) > 0
    torch.cuda.synchronize()
    got = counts.cpu().numpy().view(np.uint16)
    for k in range(nseg):
        want, _ = ck.oracle_make_hist(data[k * seg:(k + 1) * seg], 15)
        assert np.array_equal(got[k], want), k
    # 2. constant bytes: every thread hammers one bin; 6,000,000 bytes on one CTA range is 93,750 per thread (two epochs)
    const = np.full(6_000_000, 0xA7, np.uint8)
    dc = torch.from_numpy(const).cuda()
    hist = torch.zeros(256, dtype=torch.int32, device="cuda")
    assert gpu.observe_hist_device(dc.data_ptr(), const.size, hist.data_ptr(), st) > 0
    torch.cuda.synchronize()
    assert int(hist[0xA7]) == const.size and int(hist.sum()) == const.size
    c1 = torch.zeros((1, 256), dtype=torch.int16, device="cuda")
    assert gpu.make_hist_segments_device(dc.data_ptr(), const.size, const.size, 12, c1.data_ptr(), st) > 0
    torch.cuda.synchronize()
    want, _ = ck.oracle_make_hist(const, 12)
    assert np.array_equal(c1.cpu().numpy().view(np.uint16)[0], want)
    # 3. odd segment sizes from an unaligned base, and inputs below one vector
    small = _zipf(gpu, 70_001, 0.7, 22, 0)
    ds = torch.from_numpy(small).cuda()
    for seg, off in ((1, 0), (7, 1), (15, 3), (16, 5), (17, 2), (333, 7), (4099, 9)):
        n = min(small.size - off, seg * 40 + 5)
        nseg = (n + seg - 1) // seg
        cs = torch.zeros((nseg, 256), dtype=torch.int16, device="cuda")
        assert gpu.make_hist_segments_device(ds.data_ptr() + off, n, seg, 11, cs.data_ptr(), st) > 0
        torch.cuda.synchronize()
        got = cs.cpu().numpy().view(np.uint16)
        for k in range(nseg):
            want, _ = ck.oracle_make_hist(small[off + k * seg: off + min(n, (k + 1) * seg)], 11)
            assert np.array_equal(got[k], want), (seg, off, k)
    for n in (1, 2, 15, 16, 17, 31, 33):
        hist.zero_()
        assert gpu.observe_hist_device(ds.data_ptr() + 1, n, hist.data_ptr(), st) > 0
        torch.cuda.synchronize()
        assert np.array_equal(hist.cpu().numpy().astype(np.int64), np.bincount(small[1:1 + n], minlength=256)), n


def test_full_size_100mb_round_trip(gpu):
    """BASELINE.json sizes: 100,000,000-byte Zipf stream; decode(encode(x)) == x for configs 1-3 and the mt_ codec."""
    n = 100_000_000
    data = _zipf(gpu, n, 1.0, 42, 65536)
    cases = [(ck.RAW, 64, 12), (ck.RAW, 64, 15)]  # 15 bits, one unit: the wide one-lookup table (160 KB of shared memory)
    if ck.have_ref():
        cases += [(ck.MT, 64, 15), (ck.BLOCK, 32, 10), (ck.RAW, 32, 11), (ck.BLOCK, 64, 14)]
    for fam, states, bits in cases:
        stream = ck.encode(fam, states, bits, data)
        n_out, out = gpu.decode(fam, states, bits, stream, n)
        assert n_out == n, gpu.last_error()
        assert np.array_equal(out[:n], data), (fam, states, bits)


def test_corrupted_streams_never_crash_or_overrun(gpu, golden):
    """Random byte flips anywhere in a stream: the call must return 0 or n, never write past n, and leave the
    device healthy (the reference has undefined behaviour here; ours is bounded by the host-validated index and by
    the clamped word ring)."""
    rng = np.random.default_rng(12345)
    cases = [("multi", ck.MT, 64, 15), ("multi", ck.MT, 32, 12), ("multi", ck.BLOCK, 32, 10), ("multi", ck.BLOCK, 64, 13),
             ("multi", ck.RAW, 64, 12), ("runs", ck.MT, 64, 11), ("runs", ck.BLOCK, 64, 15)]
    for name, fam, states, bits in cases:
        good = golden[f"stream/{name}/{fam}/{states}/{bits}"]
        data = golden[f"in/{name}"]
        n = data.size
        for trial in range(12):
            bad = good.copy()
            flips = int(rng.integers(1, 6))
            # bias the flips towards the headers, where the framing lives
            for _ in range(flips):
                pos = int(rng.integers(0, min(bad.size, 2048))) if rng.random() < 0.6 else int(rng.integers(0, bad.size))
                bad[pos] ^= np.uint8(1 << int(rng.integers(0, 8)))
            out = np.full(n + 256, 0xCC, np.uint8)
            got, _ = gpu.decode(fam, states, bits, bad, n, out=out)
            assert got <= n, (name, fam, states, bits, trial, got)  # a flipped length field may legitimately shorten it
            assert np.all(out[n:] == 0xCC), "wrote past the decoded length"
        _check(gpu, fam, states, bits, good, data, "after fuzz")


def _index_tuple(b):
    return (b.inOffset, b.inEnd, b.outOffset, b.count, b.kind, b.symbol, b.tail)


def test_parallel_device_index_equals_host_walk(gpu, golden):
    """hsr_index.cu: K warps find headers by signature and walk their segments; the result must be the serial walk's."""
    import torch
    streams = [(golden["stream/multi/2/64/15"], 64, 15), (golden["stream/multi/2/32/12"], 32, 12),
               (golden["stream/runs/2/64/11"], 64, 11), (golden["stream/runs/2/32/14"], 32, 14),
               (golden["stream/small/2/64/10"], 64, 10), (golden["stream/tiny65/2/32/13"], 32, 13)]
    big = gpu.synth_zipf(40_000_003, 1.0, seed=3, segment_bytes=65536)
    big[5_000_000:9_000_000] = 0x20   # a long single-symbol run in the middle of a many-segment stream
    if ck.have_ref():
        streams.append((ck.ref_encode(ck.MT, 64, 15, big), 64, 15))
        streams.append((ck.ref_encode(ck.MT, 32, 10, big[:20_000_001]), 32, 10))
    streams.append((gpu.encode_mt(64, 13, big), 64, 13))
    streams.append((gpu.encode_mt(32, 15, big, 32768), 32, 15))
    for stream, states, bits in streams:
        host = [_index_tuple(b) for b in gpu.mt_index(states, stream)]
        dev_in = torch.from_numpy(stream.copy()).cuda()
        for mode in (2, 1):  # parallel only, then the serial walk
            gpu.set_option("index", mode)
            try:
                if mode == 2 and stream.size < 16 + 2 * (16 + 4 * states + 512):
                    continue  # too short for the segment scheme: auto mode falls back
                ds = gpu.PreparedStream.from_device(ck.MT, states, bits, dev_in.data_ptr(), stream.size)
                assert [_index_tuple(b) for b in ds.index()] == host, (states, bits, mode, stream.size)
                ds.free()
            finally:
                gpu.set_option("index", 0)
        # auto mode must always work and decode
        n = int(np.frombuffer(stream[:8].tobytes(), np.uint64)[0])
        ds = gpu.PreparedStream.from_device(ck.MT, states, bits, dev_in.data_ptr(), stream.size)
        out = torch.empty(n + 16, dtype=torch.uint8, device="cuda")
        ds.decode_async(out.data_ptr(), n, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        assert ds.status() == 0
        ds.free()
    # corrupted chains: the parallel path must decline or agree, never invent an index
    rng = np.random.default_rng(5)
    base, states, bits = streams[0]
    for trial in range(20):
        bad = base.copy()
        bad[int(rng.integers(16, bad.size))] ^= np.uint8(1 << int(rng.integers(0, 8)))
        dev_in = torch.from_numpy(bad).cuda()
        try:
            host = [_index_tuple(b) for b in gpu.mt_index(states, bad)]
        except gpu.HsrError:
            host = None
        try:
            ds = gpu.PreparedStream.from_device(ck.MT, states, bits, dev_in.data_ptr(), bad.size)
            got = [_index_tuple(b) for b in ds.index()]
            ds.free()
        except gpu.HsrError:
            got = None
        assert got == host, trial


def test_concurrent_host_threads_share_one_device(gpu, golden):
    """The reference decoders are re-entrant; ours serialise per device behind a mutex."""
    import threading
    jobs = [("multi", ck.MT, 64, 15), ("multi", ck.RAW, 64, 12), ("multi", ck.BLOCK, 32, 10), ("runs", ck.MT, 32, 14),
            ("small", ck.MT, 64, 12), ("multi", ck.MT, 32, 12)]
    errors = []

    def work(name, fam, states, bits):
        stream, data = golden[f"stream/{name}/{fam}/{states}/{bits}"], golden[f"in/{name}"]
        for _ in range(8):
            n, out = gpu.decode(fam, states, bits, stream, data.size)
            if n != data.size or not np.array_equal(out[:n], data):
                errors.append((name, fam, states, bits, n))

    threads = [threading.Thread(target=work, args=j) for j in jobs]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
