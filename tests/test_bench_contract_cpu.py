"""The committed bench lines under profiles/ carry every key of the bench contract (SURVEY 8d / task brief):
a change of bench.py that drops one shows up here without a GPU."""
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"}


def _line(name):
    with open(os.path.join(ROOT, "profiles", name)) as f:
        return json.loads(f.read().strip().splitlines()[-1])


def test_single_gpu_line_has_the_contract_keys():
    d = _line("r01_g_bench_full_chain.json")
    assert BASE_KEYS <= set(d)
    assert d["metric"] == "pixel_traces_per_s" and d["unit"] == "traces/s" and d["n_gpus"] == 1
    assert d["warmup"] >= 3 and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert "workload" in d["config"] and "model" not in d["config"]
    r = d["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r) and r["bound"] == "hbm"
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    # the algorithmic bytes never exceed the measured DRAM traffic of the same kernel
    assert r["traffic"] is None or r["traffic"] >= 0.99 * r["algorithmic_bytes_per_launch"]
    c = d["cpu_baseline"]
    assert {"value", "unit", "cores", "kind", "sample"} <= set(c) and c["kind"] in ("port", "reference")
    e = d["e2e"]
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(e)
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] < d["value"]
    assert d["gpu_launches"] > 0
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    # the whole step is the sum of its kernels (events inside the library) to within launch gaps
    st = d["stage_breakdown"]
    parts = sum(v["ms"] for k, v in st.items() if isinstance(v, dict) and "ms" in v)
    assert 0.9 * d["ms_per_step"] <= parts <= 1.02 * d["ms_per_step"]


def test_reference_arm_line():
    d = _line("r01_g_bench_reference_arm.json")
    assert d["impl"] == "reference" and d["metric"] == "pixel_traces_per_s"
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["cpu_baseline"]["value"] == d["value"]


@pytest.mark.parametrize("n", [2, 4, 8])
def test_scaling_lines(n):
    d = _line(f"r01_h_scale_{n}gpu.json")
    one = _line("r01_g_bench_full_chain.json")
    assert d["n_gpus"] == n and d["scaling"] == "strong" and d["cpu_baseline"] is None
    assert d["config"]["workload"] == one["config"]["workload"]
    assert one["value"] < d["value"] < n * 1.05 * one["value"]
    assert d["rank0_phases_ms"] and d["gpu_launches"] > 0
