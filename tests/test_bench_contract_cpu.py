"""CPU: the bench.py output contract that does not need a GPU — the reference arm prints exactly ONE JSON line on
stdout with the keys the driver reads, and never maps the product library."""
import json
import os
import subprocess
import sys

import pytest

import checkers as ck

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    if not ck.have_ref():
        pytest.skip("needs oracle/_ref")
    proc = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--size", "8000000", "--steps", "2", "--warmup", "1"],
                          cwd=ROOT, capture_output=True, text=True, timeout=300)
    assert proc.returncode == 0, proc.stderr[-2000:]
    lines = [l for l in proc.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, proc.stdout
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert key in d, key
    assert d["impl"] == "reference" and d["metric"] == "decoded_GBps" and d["unit"] == "GB/s" and d["higher_is_better"] is True
    assert d["steps"] == 2 and d["warmup"] == 1 and d["value"] > 0 and d["gpu_launches"] == 0
    assert d["config"]["workload"].startswith("mt_rANS32x64_16w 15-bit")
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "libhsrans_b200.so" not in d["loaded_libraries"] and "libhsrans_ref.so" in d["loaded_libraries"]


def test_ours_arm_refuses_to_run_without_a_gpu():
    """No CPU fallback anywhere: without a CUDA device the product arm exits non-zero and prints no JSON line."""
    import __graft_entry__ as entry
    if entry.load_package().device_count() > 0:
        pytest.skip("a GPU is present")
    proc = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--size", "1000000", "--steps", "1", "--warmup", "1"],
                          cwd=ROOT, capture_output=True, text=True, timeout=300)
    assert proc.returncode != 0
    assert not [l for l in proc.stdout.splitlines() if l.startswith("{")]
