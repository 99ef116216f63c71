"""The reference's OWN console harness (src/main.cpp) with the "dec B200" rows of INTEGRATION.md §1.

`make -C oracle harness` (run by __graft_entry__.build() where /root/reference exists) generates a copy of main.cpp
under the git-ignored oracle/_ref/harness/ with oracle/patch_reference_harness.py, and links it against the unmodified
reference objects and libhsrans_b200.so into oracle/_ref/hsrans_b200. That binary travels to the GPU box; nothing here
reads /root/reference at run time."""
import os
import subprocess

import numpy as np
import pytest

import checkers as ck

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HARNESS = os.path.join(ROOT, "oracle", "_ref", "hsrans_b200")


def _input(tmp_path, n):
    path = os.path.join(str(tmp_path), "zipf.bin")
    ck.synth_zipf(n, 1.0, seed=8, segment_bytes=65536).tofile(path)
    return path


def _rows(stdout):
    return [l for l in stdout.replace("\r", "\n").split("\n") if "dec B200" in l and "clk/byte" in l]


def test_patch_recipe_extends_every_row():
    """The committed recipe, applied to a miniature main.cpp of the same shape (no reference text needed)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("patch_reference_harness", os.path.join(ROOT, "oracle", "patch_reference_harness.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    rows = []
    for name, (fn, pool) in mod.NAMES.items():
        for bits in range(10, 16):
            rows.append(f'  {{ "{name}", {bits}, {{{{ "enc", e_{bits}, true }}, {{}}}}, {{{{ "dec", d_{bits}, true }}, {{}}}}}},')
    mini = "\n".join(['#include "mt_rANS32x64_16w.h"', "constexpr size_t MaxDecoderCount = 32; // c", "static const codec_info_t _Codecs[] =", "{"] + rows + ["};", ""])
    out = mod.patch(mini)
    assert out.count('"dec B200"') == 48 and out.count('"dec B200 (pool signature)"') == 12
    assert "MaxDecoderCount = 34;" in out and '#include "hsrans_b200_codecs.hpp"' in out
    assert "decode_with_thread_pool_wrapper<cuda_mt_rANS32x64_16w_decode_mt_15>" in out
    assert "cuda_block_rANS32x32_16w_decode_10, true }, {}}}," in out


def test_harness_without_a_gpu_fails_loudly(tmp_path, pkg):
    """No CPU fallback: without a CUDA device the first B200 row returns 0 and the reference's own Validate fails it."""
    if not os.path.exists(HARNESS):
        pytest.skip("oracle/_ref/hsrans_b200 not built (needs /root/reference at build time)")
    if pkg.device_count() > 0:
        pytest.skip("a CUDA device is present: covered by the gpu test")
    proc = subprocess.run([HARNESS, _input(tmp_path, 300_000), "--test"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert proc.returncode == 1 and "Failed to validate" in proc.stdout
    assert "dec B200" in proc.stdout and "decompressed to 0 bytes" in proc.stdout


@pytest.mark.gpu
def test_reference_harness_validates_every_b200_row(tmp_path, pkg):
    """hsrans <8 MB zipf> --test: every codec, every decoder variant, every B200 row through the reference's own
    0xCC-poison / size / memcmp protocol (src/main.cpp:860-897); exit status 0 = nothing differed anywhere."""
    assert pkg.device_count() >= 1
    if not os.path.exists(HARNESS):
        pytest.fail("oracle/_ref/hsrans_b200 is missing: build it in the container (`make -C oracle harness`), it travels with gpurun")
    proc = subprocess.run([HARNESS, _input(tmp_path, 8_000_000), "--test"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=1500)
    rows = _rows(proc.stdout)
    tail = "\n".join(proc.stdout.replace("\r", "\n").split("\n")[-12:])
    assert proc.returncode == 0, tail
    assert "Failed to validate" not in proc.stdout
    assert len(rows) == 60, (len(rows), tail)   # 48 codec rows + 12 pool-signature twins
    print("\n" + "\n".join(rows[:3]) + f"\n... {len(rows)} 'dec B200' rows validated by the reference harness")
