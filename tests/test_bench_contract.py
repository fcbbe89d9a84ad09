"""bench.py's reference arm runs without a GPU (it times the oracle port on the host cores): check the JSON line it prints
carries the keys the driver reads, and that the product arm refuses to run without a CUDA device instead of falling back."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], cwd=ROOT, capture_output=True, text=True, timeout=600)


def test_reference_arm_json_line():
    p = _run("--impl", "reference", "--steps", "1", "--warmup", "1", "--ref-step-seconds", "0.5")
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "reads/s" and d["unit"] == "reads/s"
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 1 and d["higher_is_better"] is True
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["vs_baseline"] is None and d["dtype"] == "f64"
    assert d["config"]["workload"].startswith("C2")
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "reads" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_product_arm_needs_a_gpu():
    import ctypes as C
    sys.path.insert(0, ROOT)
    from clique_b200 import _lib as L
    if L.load_library().clq_device_count() > 0:
        import pytest
        pytest.skip("a GPU is present")
    p = _run("--steps", "1", "--warmup", "1", "--reads", "64", "--no-cpu-baseline", "--no-live-peak")
    assert p.returncode != 0, "bench.py must not produce a number without the CUDA path"
    assert not any(l.startswith("{") and '"value"' in l for l in p.stdout.splitlines())
