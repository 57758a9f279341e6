"""bench.py's reference arm runs on host cores only (it is the one bench leg a CPU box can execute): check that it follows the
JSON-line contract of the driver -- one line, the BASELINE.json metric, the keys and types the base contract names."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1
    j = json.loads(lines[0])
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert j["impl"] == "reference" and j["unit"] == "pairs/s" and j["higher_is_better"] is True and j["scaling"] == "weak"
    assert "pairs/sec" in j["metric"] and "36 regions" in j["metric"]
    assert str(base.get("metric", "")).split()[0].lower() in j["metric"].lower() or "pairs" in j["metric"]
    assert j["n_gpus"] == 1 and j["steps"] == 1 and j["value"] > 0 and j["ms_per_step"] > 0
    assert j["data"] == "synthetic" and j["dtype"] == "f32" and j["vs_baseline"] is None
    assert "workload" in j["config"] and "model" not in j["config"]
    cb = j["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] == j["value"] and cb["sample"]
    e2e = j["e2e"]
    assert e2e["value"] == j["value"] and e2e["unit"] == j["unit"]
    assert e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0


def test_gpu_arm_fails_loudly_without_a_gpu():
    """No CUDA device: the product arm must refuse (non-zero exit, no JSON line) instead of falling back to anything."""
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a GPU is present")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1", "--no-cpu-baseline"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode != 0
    assert "no CPU fallback" in (r.stderr + r.stdout)
    assert not [l for l in r.stdout.splitlines() if l.startswith("{")]
