"""The JSON line bench.py prints is a contract with the driver: the committed lines of the last GPU runs (profiles/) must
carry every key it reads, with consistent arithmetic.  CPU-only: nothing is executed on a GPU here."""
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _line(name):
    path = os.path.join(ROOT, "profiles", name)
    lines = [l for l in open(path).read().splitlines() if l.startswith("{")]
    assert lines, path
    return json.loads(lines[-1])


@pytest.mark.parametrize("name,n", [("r02_bench_n1.json", 1), ("r02_bench_n2.json", 2), ("r02_bench_n8.json", 8)])
def test_bench_line_carries_the_contract(name, n):
    d = _line(name)
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert k in d, k
    assert d["n_gpus"] == n and d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["dtype"] == "f32" and d["data"] == "synthetic" and "workload" in d["config"] and "model" not in d["config"]
    assert d["gpu_launches"] > 0 and d["warmup"] >= 3
    # value = whole-job vertex-frames per second of exactly `steps` steps
    per_step = 1_000_000 * d["config"]["slots_per_step_per_gpu"] * n
    assert d["value"] == pytest.approx(per_step / (d["ms_per_step"] * 1e-3), rel=1e-6)
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and r["frac"] == pytest.approx(r["achieved"] / r["peak"], rel=1e-9)
    assert r["traffic"] and r["achieved"] == pytest.approx(r["traffic"] / (r["avg_launch_ms"] * 1e-3) / 1e9, rel=1e-6)
    assert 0.0 < r["frac"] < 1.0, "the physical fraction cannot exceed the measured peak"
    e = d["e2e"]
    assert e["unit"] == d["unit"] and e["d2h_bytes_per_step"] == 1_000_000 * 24 * d["config"]["slots_per_step_per_gpu"]
    assert e["value"] < d["value"] and e["h2d_bytes_per_step"] > 0
    c = d["clocks"]
    assert not set(c["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    if n == 1:
        b = d["cpu_baseline"]
        assert b["kind"] in ("reference", "port") and b["cores"] >= 1 and b["value"] > 0 and b["sample"]
    also = d["also"]
    for k in ("C4", "C5", "C1", "C2_128", "C2_256", "C2_512", "C3_random"):
        assert k in also and also[k]["value"] > 0, k
    assert also["C4"]["scaling"] == "strong" and also["C5"]["scaling"] == "strong"
    if n > 1:
        g = also["C5"]["gather"]
        assert g["comm_nranks_seen"] == n and g["gbs_into_root"] > 0
        assert also["C5"]["p2p_fused"]["bit_identical_to_local_bake"] is True


def test_reference_arm_line():
    d = _line("r02_bench_reference_n1.json")
    g = _line("r02_bench_n1.json")
    assert d["impl"] == "reference" and d["metric"] == g["metric"] and d["unit"] == g["unit"] and d["config"] == g["config"]
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["kind"] in ("reference", "port")
