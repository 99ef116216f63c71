"""The C++ wrappers carry the reference's decodeFunc type and names: build the harness (CPU), run it (GPU)."""
import os
import subprocess

import pytest

from conftest import ROOT

SRC = os.path.join(ROOT, "tests", "cpp", "drop_in_harness.cpp")


def _build(tmp_path, pkg):
    exe = str(tmp_path / "drop_in_harness")
    libdir = os.path.dirname(pkg.lib_path())
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-Werror", SRC, "-o", exe, f"-L{libdir}", "-lhsrans_b200",
                           f"-Wl,-rpath,{libdir}"])
    return exe


def test_wrappers_compile_link_and_list_every_reference_row(tmp_path, pkg):
    exe = _build(tmp_path, pkg)
    names = subprocess.check_output([exe, "--list"], text=True).split()
    assert len(names) == 48
    for fam in ("rANS32x32_16w", "rANS32x64_16w", "block_rANS32x32_16w", "block_rANS32x64_16w", "mt_rANS32x32_16w",
                "mt_rANS32x64_16w", "rANS32x16_16w", "rANS32x32_32blk_16w"):
        for bits in range(10, 16):
            assert f"cuda_{fam}_decode_{bits}" in names


@pytest.mark.gpu
def test_harness_protocol_validates_on_gpu(tmp_path, pkg, golden, golden_rank4):
    exe = _build(tmp_path, pkg)
    for name, key, inp in [("cuda_rANS32x16_16w_decode_12", "multi/0/16/12", "multi"),
                           ("cuda_rANS32x32_32blk_16w_decode_15", "multi/3/32/15", "multi")]:
        s, e = tmp_path / "stream.bin", tmp_path / "expected.bin"
        golden_rank4[f"stream/{key}"].tofile(s)
        golden_rank4[f"in/{inp}"].tofile(e)
        res = subprocess.run([exe, name, str(s), str(e)], capture_output=True, text=True)
        assert res.returncode == 0, (name, res.stdout, res.stderr)
    cases = [("cuda_mt_rANS32x64_16w_decode_15", "multi/2/64/15", "multi"), ("cuda_block_rANS32x32_16w_decode_10", "multi/1/32/10", "multi"),
             ("cuda_rANS32x64_16w_decode_12", "multi/0/64/12", "multi"), ("cuda_rANS32x32_16w_decode_11", "small/0/32/11", "small"),
             ("cuda_mt_rANS32x32_16w_decode_14", "runs/2/32/14", "runs")]
    for name, key, inp in cases:
        s, e = tmp_path / "stream.bin", tmp_path / "expected.bin"
        golden[f"stream/{key}"].tofile(s)
        golden[f"in/{inp}"].tofile(e)
        res = subprocess.run([exe, name, str(s), str(e)], capture_output=True, text=True)
        assert res.returncode == 0, (name, res.stdout, res.stderr)
    # wrong codec for the stream -> the harness reports a validation failure, like the reference's --test
    golden["stream/multi/2/64/15"].tofile(tmp_path / "stream.bin")
    golden["in/multi"].tofile(tmp_path / "expected.bin")
    res = subprocess.run([exe, "cuda_mt_rANS32x64_16w_decode_12", str(tmp_path / "stream.bin"), str(tmp_path / "expected.bin")],
                         capture_output=True, text=True)
    assert res.returncode == 1
