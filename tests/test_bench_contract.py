"""The bench line the driver parses: every key of the contract is present in the last committed B200 run
(profiles/*_bench_e2e_b1.json is bench.py's stdout, verbatim) and in the reference arm's line.  CPU-only."""
import glob
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _latest(pattern):
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", pattern)))
    assert files, pattern
    return json.load(open(files[-1]))


def test_bench_line_contract():
    d = _latest("r*_bench_e2e_b1.json")
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["n_gpus"] == 1 and d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert "workload" in d["config"] and "model" not in d["config"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in d["roofline"], k
    assert abs(d["roofline"]["frac"] - d["roofline"]["achieved"] / d["roofline"]["peak"]) < 1e-9
    for k in ("value", "unit", "cores", "kind", "sample"):
        assert k in d["cpu_baseline"], k
    assert d["cpu_baseline"]["kind"] in ("port", "reference")
    for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
        assert k in d["e2e"], k
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0   # host buffers inside the timed region
    assert d["e2e"]["value"] != d["value"]
    assert d["gpu_launches"] > 0
    for k in ("sm_mhz", "sm_max_mhz", "reasons"):
        assert k in d["clocks"], k
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    # value = units / time
    assert abs(d["value"] - d["config"]["utterances_per_gpu"] * 1e3 / d["ms_per_step"]) / d["value"] < 1e-6


def test_reference_arm_contract():
    r = _latest("r*_bench_reference_arm.json")
    assert r["impl"] == "reference" and r["gpu_launches"] == 0
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "config", "cpu_baseline", "e2e"):
        assert k in r, k
    assert r["e2e"]["value"] == r["value"] and r["e2e"]["h2d_bytes_per_step"] == 0 and r["e2e"]["d2h_bytes_per_step"] == 0
    assert r["cpu_baseline"]["value"] == r["value"] and r["cpu_baseline"]["cores"] >= 1
    d = _latest("r*_bench_e2e_b1.json")
    assert r["metric"] == d["metric"] and r["unit"] == d["unit"] and r["config"]["workload"] == d["config"]["workload"]
