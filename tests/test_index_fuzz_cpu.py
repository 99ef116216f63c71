"""CPU: the host header-chain walk must survive arbitrary corruption (no crash, no out-of-range record)."""
import numpy as np

import checkers as ck


def test_mt_index_on_corrupted_chains(pkg, golden):
    rng = np.random.default_rng(7)
    for key, states in (("stream/multi/2/64/15", 64), ("stream/multi/2/32/12", 32), ("stream/runs/2/64/11", 64)):
        good = golden[key]
        n = int(np.frombuffer(good[:8].tobytes(), np.uint64)[0])
        for trial in range(300):
            bad = good.copy()
            for _ in range(int(rng.integers(1, 5))):
                pos = int(rng.integers(0, bad.size))
                bad[pos] = np.uint8(rng.integers(0, 256))
            if trial % 3 == 0:
                bad = bad[: int(rng.integers(1, bad.size))]
            try:
                blocks = pkg.mt_index(states, bad)
            except pkg.HsrError:
                continue
            hdr_n = int(np.frombuffer(bad[:8].tobytes(), np.uint64)[0]) if bad.size >= 8 else 0
            pos = 0
            for b in blocks:
                assert b.outOffset == pos and b.inOffset < bad.size and b.inEnd <= bad.size
                pos += b.count
            assert pos == hdr_n
            # the oracle walk agrees on whether the chain is well formed
            assert ck.oracle_mt_walk(states, bad) is not None
