"""CPU suite, part 3: the bench line of record kept under profiles/ carries every key the driver's
contract names (a guard against silently dropping one while editing bench.py), and bench.py's
algorithmic-work helpers agree with SURVEY.md §8(d)'s figures."""
import importlib.util
import json
import os

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def load_bench():
    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_recorded_bench_line_has_the_contract_keys():
    d = json.load(open(os.path.join(ROOT, "profiles", "bench_r2_final3.json")))
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
              "scaling", "vs_baseline", "dtype", "data", "config", "roofline", "cpu_baseline", "e2e",
              "gpu_launches", "clocks"):
        assert k in d, k
    assert d["config"]["workload"].startswith("configs[1]") and "model" not in d["config"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in d["roofline"], k
    assert abs(d["roofline"]["frac"] - d["roofline"]["achieved"] / d["roofline"]["peak"]) < 1e-9
    for k in ("value", "unit", "cores", "kind", "sample"):
        assert k in d["cpu_baseline"], k
    for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
        assert k in d["e2e"], k
    assert d["e2e"]["h2d_bytes_per_step"] == 8 * 128 * (1 << 16)
    assert d["gpu_launches"] > 0 and d["vs_baseline"] is None and d["dtype"] == "u64"
    assert d["self_checks"]["failed"] == []
    assert d["e2e_pageable"]["inputs_only"]["ms_per_step"] < d["e2e_pageable"]["inputs_only"]["ms_per_step_driver_staging"]
    assert d["step_standin"]["device_quotient_ms"] > 0
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    ref = json.load(open(os.path.join(ROOT, "profiles", "bench_r2_reference_arm3.json")))
    assert ref["impl"] == "reference" and ref["metric"] == d["metric"] and ref["unit"] == d["unit"]
    assert ref["config"]["workload"] == d["config"]["workload"]
    assert ref["e2e"]["h2d_bytes_per_step"] == 0 and ref["e2e"]["d2h_bytes_per_step"] == 0


def test_algorithmic_work_matches_the_survey():
    b = load_bench()
    n = 1 << 16
    assert b.algorithmic_bytes(128, n, 3, 4) == 704642560       # µ
    assert b.permutations(128, n, 3, 4) == 8912880
    assert sum(b.permutations(c, n, 3, 4) for c in (135, 20, 16)) == 13107152  # S, per step
    assert b.IMAD_PER_PERMUTATION == 6700
