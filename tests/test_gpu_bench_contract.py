"""The driver's bench contract: one JSON line with the agreed keys, from both arms."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, out.stdout[-2000:]
    return json.loads(lines[0])


def test_native_arm_line():
    d = _run("--steps", "3", "--warmup", "3", "--batch", "4", "--cpu-baseline-frames", "1", "--no-extra")
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["metric"] == "frames_per_sec_1080p" and d["unit"] == "frames/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["steps"] == 3 and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert "workload" in d["config"] and "frontalface_alt" in d["config"]["workload"]
    assert d["value"] > 0 and d["gpu_launches"] > 0
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] == 4 * 1920 * 1080 and e["d2h_bytes_per_step"] > 0
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor") and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-3
    assert set(("sm_mhz", "sm_max_mhz", "reasons")) <= set(d["clocks"])
    c = d["cpu_baseline"]
    assert c["kind"] in ("port", "reference") and c["cores"] >= 1 and c["value"] > 0 and c["sample"]
    # the reference's own CPU detector (CLOD-CPU, clod.cpp:1339-1500) beside it
    assert c["clod_cpu"]["value"] > 0 and c["clod_cpu"]["cores"] >= 1 and c["clod_cpu"]["windows_per_frame"] > 0


def test_reference_arm_line():
    d = _run("--impl", "reference", "--steps", "1", "--warmup", "1", "--ref-frames", "1")
    assert d["impl"] == "reference" and d["metric"] == "frames_per_sec_1080p" and d["value"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["cores"] >= 1
    assert d["gpu_launches"] == 0 and d["cpu_baseline"]["clod_cpu"]["value"] > 0
