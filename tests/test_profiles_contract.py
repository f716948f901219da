"""The committed bench lines under profiles/ carry what the bench contract asks for (keys, units, internal consistency)."""
import glob
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LINES = sorted(glob.glob(os.path.join(ROOT, "profiles", "r2_bench_*.json")))
REQUIRED = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
            "data", "config", "e2e"]


def _line(path):
    return json.loads(open(path).read().strip().splitlines()[-1])


@pytest.mark.parametrize("path", LINES, ids=[os.path.basename(p) for p in LINES])
def test_bench_line_has_the_contract_keys(path):
    d = _line(path)
    for k in REQUIRED:
        assert k in d, k
    assert d["metric"] == "aligned_gbp_per_s" and d["unit"] == "Gbp/s" and d["higher_is_better"] is True
    assert "workload" in d["config"] and d["warmup"] >= 1 and d["vs_baseline"] is None
    assert set(d["e2e"]) >= {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"}
    if d.get("impl") == "reference":
        assert d["cpu_baseline"]["kind"] == "port" and d["e2e"]["h2d_bytes_per_step"] == 0
        return
    assert d["scaling"] == "weak" and d["data"] == "synthetic" and d["warmup"] >= 3
    assert d["gpu_launches"] > 0 and "roofline" in d and "clocks" in d
    r = d["roofline"]
    assert set(r) >= {"bound", "achieved", "peak", "unit", "frac", "traffic"}
    if r["achieved"] is not None:
        assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-6
        # the dominant kernel's time fits inside the step
        assert r["kernel_ms"] <= 1.5 * d["ms_per_step"] * max(1, d.get("config", {}).get("reads_per_gpu_per_step", 1) and 1)
    cl = d["clocks"]
    assert not set(cl["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    assert cl["sm_mhz"] is None or cl["sm_mhz"] >= 0.9 * cl["sm_max_mhz"]


def test_headline_lines_are_present_and_consistent():
    names = {os.path.basename(p) for p in LINES}
    for want in ("r2_bench_default_10k.json", "r2_bench_reference_arm.json", "r2_bench_n4_torchrun.json", "r2_bench_n8_torchrun.json",
                 "r2_bench_cfg2_n1.json", "r2_bench_cfg2_n8_torchrun.json", "r2_bench_cfg3_n4_torchrun.json", "r2_bench_cfg4_n1.json"):
        assert want in names, want
    one = _line(os.path.join(ROOT, "profiles", "r2_bench_default_10k.json"))
    eight = _line(os.path.join(ROOT, "profiles", "r2_bench_n8_torchrun.json"))
    assert one["n_gpus"] == 1 and eight["n_gpus"] == 8
    # value = whole-job throughput: reads x bases per step / time
    for d in (one, eight):
        bases = d["config"]["reads_per_gpu_per_step"] * d["config"]["read_len"] * d["n_gpus"]
        assert abs(d["value"] - bases / (d["ms_per_step"] / 1e3) / 1e9) / d["value"] < 0.02
    assert one["cpu_baseline"]["cores"] >= 1 and one["cpu_baseline"]["value"] < one["e2e"]["value"]
    assert 0.5 < eight["value"] / (8 * one["value"]) <= 1.05
