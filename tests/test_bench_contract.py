"""CPU tier: the committed bench lines carry every key the bench contract names (the GPU box writes them; this
checks the evidence under profiles/ and the reference-arm line against the same list, so a key cannot go missing
unnoticed)."""
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LINE_KEYS = ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks")


def _line(name):
    with open(os.path.join(ROOT, "profiles", name)) as fh:
        return json.loads(fh.read().strip().splitlines()[-1])


@pytest.mark.parametrize("name", ["r02_bench_default_ieee123_v3.json", "r02_bench_weak_n8_final.json"])
def test_gpu_arm_line(name):
    j = _line(name)
    missing = [k for k in LINE_KEYS if k not in j and not (k == "cpu_baseline" and j["n_gpus"] > 1)]
    assert not missing, missing
    with open(os.path.join(ROOT, "BASELINE.json")) as fh:
        base = json.load(fh)
    assert base["metric"].startswith(j["metric"]) and j["higher_is_better"] is True and j["vs_baseline"] is None
    assert j["dtype"] == "f64" and j["data"] == "synthetic" and "workload" in j["config"]
    r = j["roofline"]
    assert r["bound"] in ("hbm", "tensor") and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert r["traffic"] and r["traffic"] > 0
    e = j["e2e"]
    assert all(k in e for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"))
    assert 0 < e["value"] <= j["value"] * 1.001 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert j["gpu_launches"] == j["steps"]                      # one kernel launch per step in the timed region
    c = j["clocks"]
    assert not set(c["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    assert c["sm_mhz"] >= 0.9 * c["sm_max_mhz"]
    if j["n_gpus"] == 1:
        b = j["cpu_baseline"]
        assert all(k in b for k in ("value", "unit", "cores", "kind", "sample")) and b["kind"] in ("port", "reference")
    # every other BASELINE configuration rides along
    for k in ("ieee13", "ieee34", "synthetic1000", "synthetic1000_rollout", "ieee123_mesh"):
        assert j["configs"][k]["value"] > 0, k


def test_reference_arm_line():
    j = _line("r02_bench_reference_arm.json")
    assert j["impl"] == "reference" and j["gpu_launches"] == 0
    assert j["cpu_baseline"]["kind"] in ("port", "reference") and j["cpu_baseline"]["value"] == j["value"]
    assert j["e2e"] == {"value": j["value"], "unit": j["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
