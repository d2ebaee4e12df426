"""The evidence under profiles/ is machine-readable: the launch-list summariser reproduces the committed shares, the bench
records carry the contract keys, and the ncu constants bench.py may quote are keyed by source hashes."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_launch_list_summary_reproduces_the_committed_shares():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_summary.py"), "--launches",
                          os.path.join(ROOT, "profiles", "r2b_launches_bench.csv")], capture_output=True, text=True, check=True).stdout
    rows = [l.split() for l in out.strip().splitlines()]
    top = rows[0]
    assert top[-1] == "pnb::k_knn" and 75.0 < float(top[2]) < 90.0          # the search dominates the step
    shares = sum(float(r[2]) for r in rows[:-1])
    assert abs(shares - 100.0) < 0.5
    committed = open(os.path.join(ROOT, "profiles", "r2b_launches_summary.txt")).read()
    assert " ".join(top[-2:]) in " ".join(committed.split())


def test_bench_records_carry_the_contract_keys():
    for name in ("r2b_bench_n1.json", "r2b_bench_n1_f32.json", "r2b_bench_n8.json"):
        with open(os.path.join(ROOT, "profiles", name)) as f:
            d = json.loads(f.read().strip().splitlines()[-1])
        for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "dtype",
                    "data", "config", "e2e", "gpu_launches", "clocks", "roofline"):
            assert key in d, (name, key)
        assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["gpu_launches"] > 0
        assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}


def test_ncu_constants_are_keyed_by_source_hashes():
    sys.path.insert(0, ROOT)
    import bench
    with open(os.path.join(ROOT, "profiles", "r2_ncu_constants.json")) as f:
        c = json.load(f)
    assert set(c["source_hashes"]) == set(bench.KERNEL_SOURCES)
    for kernel in bench.KERNEL_SOURCES:
        v = bench.load_ncu(f"{kernel}:f64:256:dram_bytes")
        if c["source_hashes"][kernel] == bench.source_hash(kernel):
            assert v and v > 1e8                    # quoted only for the very sources the capture was taken on
        else:
            assert v is None
