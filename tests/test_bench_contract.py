"""The bench line's contract (driver-facing keys, roofline and cpu_baseline objects, e2e with its byte counts), checked on
the committed output of the last GPU round -- the same code paths build the line on every run -- and the workload
definitions both arms share.  CPU only: nothing here touches a GPU."""
import glob
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _latest(pattern):
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", pattern)))   # round tags sort by name: r19 < r2j < r3a < r4a < r5e
    if not files:
        pytest.skip("no committed bench output")
    lines = [l for l in open(files[-1]).read().splitlines() if l.startswith("{")]
    return json.loads(lines[-1])


def test_gpu_arm_line_has_the_contract_keys():
    d = _latest("r*_bench.json")
    baseline = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert d["metric"] == baseline["metric"] and d["unit"] == "MS/s"
    for key in ("value", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
                "data", "config", "e2e", "gpu_launches", "roofline", "cpu_baseline", "clocks"):
        assert key in d, key
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["warmup"] >= 3 and d["gpu_launches"] > 0 and d["data"] == "synthetic"
    assert "workload" in d["config"] and "configs[4]" in d["config"]["workload"] and "model" not in d["config"]
    # the step moves what the config says it moves
    assert abs(d["value"] - d["config"]["input_complex_samples_per_step_per_gpu"] * d["n_gpus"] / d["ms_per_step"] / 1e3) < 1e-6 * d["value"]
    e = d["e2e"]
    assert e["unit"] == d["unit"] and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert 0 < e["value"] < d["value"]            # host buffers can only cost time
    assert e["h2d_bytes_per_step"] == 2 * d["config"]["input_complex_samples_per_step_per_gpu"]   # 8-bit I/Q
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor", "fp32") and r["unit"] in ("GB/s", "TFLOP/s")
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and 0 < r["frac"] < 1
    assert r["traffic"] is None or r["traffic"] > 0
    for k in r["kernels"]:
        assert 0 < k["frac"] < 1 and k["ms"] > 0, k
    c = d["cpu_baseline"]
    assert c["kind"] in ("reference", "port") and c["cores"] >= 1 and c["value"] > 0 and c["unit"] == d["unit"] and c["sample"]
    assert d["clocks"]["sm_mhz"] > 0 and not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}


def test_reference_arm_line_matches_the_gpu_arm():
    g, r = _latest("r*_bench.json"), _latest("r*_bench_reference.json")
    assert r["impl"] == "reference"
    for key in ("metric", "unit", "higher_is_better", "config"):
        assert r[key] == g[key], key
    assert r["e2e"]["value"] == r["value"] and r["e2e"]["h2d_bytes_per_step"] == 0 and r["e2e"]["d2h_bytes_per_step"] == 0
    assert r["cpu_baseline"]["value"] == r["value"] and r["cpu_baseline"]["kind"] in ("reference", "port")


def test_workload_definitions():
    import bench
    w = bench.WORKLOADS["c4fm_20m"]
    assert w["fs"] == 20e6 and w["demod"] == "c4fm"
    cfg = bench.workload_config("c4fm_20m", 8)
    assert cfg["tuners_per_gpu"] == 8 and cfg["channels_per_tuner"] == 800 and "configs[4]" in cfg["workload"]
    # 72 filter bank + FFT + 2 channel samples x (4 x 72 taps + 40)
    assert 800 < bench.chain_flop_per_input_sample(800, 72) < 850
