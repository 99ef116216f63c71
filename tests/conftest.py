import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def pkg():
    import __graft_entry__ as entry
    if not os.path.exists(os.path.join(entry.PKG_DIR, "libhsrans_b200.so")):
        entry.build()
    return entry.load_package()


@pytest.fixture(scope="session")
def golden():
    path = os.path.join(ROOT, "tests", "golden", "golden.npz")
    z = np.load(path)
    return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def golden_rank4():
    """rANS32x16_16w and rANS32x32_32blk_16w vectors (tests/golden/make_golden_rank4.py)"""
    z = np.load(os.path.join(ROOT, "tests", "golden", "golden_rank4.npz"))
    return {k: z[k] for k in z.files}


def golden_stream_cases(golden):
    """[(name, family, states, bits, stream, expected_return, input)]"""
    cases = []
    for key in sorted(golden):
        if not key.startswith("stream/"):
            continue
        _, name, fam, states, bits = key.split("/")
        ret = int(golden[f"ret/{name}/{fam}/{states}/{bits}"][0])
        cases.append((name, int(fam), int(states), int(bits), golden[key], ret, golden[f"in/{name}"]))
    return cases
