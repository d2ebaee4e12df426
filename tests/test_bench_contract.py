"""The parts of bench.py's contract that can be checked without a GPU: the reference arm runs the compiled
reference (oracle/_ref) on the host cores and prints one JSON line with the driver's keys; ranks other than 0 stay
silent and exit 0."""
import json
import os
import subprocess
import sys

import pytest

from oracle import ref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra_env=None, *flags):
    env = dict(os.environ)
    env.update(extra_env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--n-side", "20", "--render-res", "64",
                           "--steps", "2", "--warmup", "1", *flags], capture_output=True, text=True, env=env, timeout=300)


@pytest.mark.skipif(not ref.available(), reason="compiled reference not built")
def test_reference_arm_prints_the_contract_line():
    p = _run()
    assert p.returncode == 0, p.stderr
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "particles/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("particles/sec for sim['rho']") and d["steps"] == 2 and d["warmup"] == 1
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["dtype"] == "f64" and d["data"] == "synthetic"
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"] and "sample" in d["cpu_baseline"]
    assert d["e2e"] == {"value": d["value"], "unit": "particles/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]
    # the arm runs the workload it names (the same particle set as the GPU arm), and the render half of the metric
    assert d["config"]["particles_per_gpu"] == 20 ** 3 and "8000 particles" in d["cpu_baseline"]["sample"]
    img = d["render"]["image"]
    assert img["mpix_per_s"] > 0 and img["cpu_baseline"]["unit"] == "Mpix/s" and img["resolution"] == 64


def test_reference_arm_is_silent_on_other_ranks():
    p = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}, "--gpus", "2")
    assert p.returncode == 0 and p.stdout.strip() == ""


def test_ncu_constants_are_quoted_only_for_the_sources_they_were_captured_on(tmp_path, monkeypatch):
    """roofline.traffic & co. come from a committed ncu capture; bench.py must refuse to print them once the kernel
    sources have changed (a stale constant is worse than a null)."""
    sys.path.insert(0, ROOT)
    import bench
    const = json.load(open(os.path.join(ROOT, "profiles", "r2_ncu_constants.json")))
    key = "k_knn:f64:256:dram_bytes"
    assert key in const and const[key] > 0
    if const["source_hashes"]["k_knn"] == bench.source_hash("k_knn"):
        assert bench.load_ncu(key) == const[key]
    else:
        assert bench.load_ncu(key) is None
    monkeypatch.setattr(bench, "source_hash", lambda kernel="k_knn": "0" * 16)
    assert bench.load_ncu(key) is None
