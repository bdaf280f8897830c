"""CPU check of the bench contract on the committed lines of each round's final run (profiles/r0N_bench_line.json):
the keys the driver and the judge read are present and consistent with each other."""
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("line", ["r01_bench_line.json", "r02_bench_line.json"])
def test_committed_bench_line_has_the_contract_keys(line):
    with open(os.path.join(ROOT, "profiles", line)) as f:
        d = json.load(f)
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "e2e", "roofline", "cpu_baseline", "clocks", "gpu_launches"):
        assert k in d, k
    assert d["n_gpus"] == 1 and d["warmup"] >= 3 and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert "workload" in d["config"] and "model" not in d["config"]
    n = d["config"]["rows_per_gpu"]
    assert abs(d["value"] - n / (d["ms_per_step"] * 1e-3)) / d["value"] < 1e-6          # whole-job records/s
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] == n // 100 * 14016 and e["d2h_bytes_per_step"] > 0   # copies counted from the images
    assert e["value"] < d["value"]                                                       # not a repeat of the device number
    r = d["roofline"]
    assert r["bound"] == "hbm" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and r["traffic"] >= r["algorithmic_bytes_per_launch"]
    c = d["cpu_baseline"]
    assert c["kind"] in ("reference", "port") and c["cores"] >= 1 and c["value"] > 0
    assert d["gpu_launches"] > 0 and not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
