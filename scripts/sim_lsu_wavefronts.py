"""sim_lsu_wavefronts.py — shared-memory wavefronts per half-row of the 15-bit decode loop, simulated on the CPU.

The units kernel is bound by the SM's shared-memory data pipe (profiles/r1/ncu_v14_summary.txt: 87.7 % busy, 56 % of its
wavefronts are bank-conflict replays). This script replays the two table lookups of `symbol_step_rank` (group word at
bank (slot >> 4) & 31, entry word at bank rank & 31) for the bench's own bytes (Zipf(1) pw64k, 64 KiB blocks, 15 bits,
lane -> byte position as in idx2idx_lane) and counts wavefronts = max over banks of DISTINCT words per warp request. It
reproduces the ncu counters of the shipped kernel (3.46 / 2.67 measured, 3.48 / 2.67 simulated; pinned by
tests/test_sim_lsu_cpu.py), so the table-layout ideas of VERDICT r1 task 2 can be priced without GPU time:
  ent_perm    entry placement idx = (rank * m) & 255, m chosen per block by minimising the sum of squared bank masses
  ent_sorted  upper bound for ANY static placement: entries sorted by frequency, dealt round-robin over the banks
  *_topK      the K most frequent symbols served from warp-uniform registers, their lanes predicated off both lookups
Results: profiles/r2/sim_lsu_wavefronts.txt. Development tool; reads tests/checkers.py only for the synthetic bytes.

    python scripts/sim_lsu_wavefronts.py [blocks]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)

BITS = 15
T = 1 << BITS
PERM_MULTIPLIERS = (1, 3, 5, 7, 9, 11, 13, 15, 17, 19, 21, 23, 25, 27, 29, 31, 33, 37, 41, 45, 51, 57, 63, 73, 85, 97, 113, 127)


def idx2idx(lane):
    return (lane & 3) | ((lane & 4) << 2) | ((lane & 24) >> 1)


LANE_POS = np.array([idx2idx(l) for l in range(32)])


def wavefronts(words, active=None):
    """words: (rows, 32) word indices of one warp request per row -> wavefronts per row = the busiest bank's number of
    DISTINCT words (same word = broadcast). `active` masks lanes that are predicated off."""
    out = np.zeros(words.shape[0], int)
    for r in range(words.shape[0]):
        w = words[r] if active is None else words[r][active[r]]
        if w.size:
            out[r] = np.bincount(np.unique(w) & 31, minlength=32).max()
    return out


def normalised_hist(block):
    """every symbol present (IsSafeHist, src/mt_rANS32x64_16w_encode.cpp:193-203), counts scaled to 2^15"""
    c = np.bincount(block, minlength=256).astype(np.int64) + 1
    f = np.maximum(1, (c * T) // c.sum())
    f[np.argmax(f)] += T - f.sum()
    return f


def simulate(blocks=16, seed=1):
    import checkers as ck
    rng = np.random.default_rng(seed)
    data = ck.synth_zipf(blocks * 65536, 1.0, seed=42, segment_bytes=65536)
    acc = {}

    def add(key, value):
        acc.setdefault(key, []).append(value)

    for b in range(blocks):
        blk = data[b * 65536:(b + 1) * 65536]
        f = normalised_hist(blk)
        cum = np.concatenate([[0], np.cumsum(f)[:-1]])
        p = f / T
        order = np.argsort(-f)
        place = np.empty(256, int)
        place[order] = np.arange(256)
        best_m = min(PERM_MULTIPLIERS, key=lambda m: (np.bincount(((np.arange(256) * m) & 255) & 31, weights=p, minlength=32) ** 2).sum())
        rows = blk.reshape(-1, 64)
        for half in range(2):
            sym = rows[:, half * 32 + LANE_POS].astype(np.int64)                     # (1024, 32) in lane order
            slot = cum[sym] + (rng.random(sym.shape) * f[sym]).astype(np.int64)      # uniform inside the symbol's range
            add("grp", wavefronts(slot >> 4).mean())
            add("ent", wavefronts(sym).mean())
            add("ent_perm", wavefronts((sym * best_m) & 255).mean())
            add("ent_sorted", wavefronts(place[sym]).mean())
            for k in (1, 2, 4):
                act = ~np.isin(sym, order[:k])
                add(f"grp_top{k}", wavefronts(slot >> 4, act).mean())
                add(f"ent_top{k}", wavefronts(sym, act).mean())
    return {k: float(np.mean(v)) for k, v in acc.items()}


if __name__ == "__main__":
    for key, value in simulate(int(sys.argv[1]) if len(sys.argv) > 1 else 16).items():
        print(key, round(value, 3))
