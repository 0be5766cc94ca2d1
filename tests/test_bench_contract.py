"""CPU: the JSON lines bench.py printed on the B200 (committed under profiles/) carry every key of the measurement
contract, and bench.py's stdout discipline (exactly one line on the real stdout) works."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PROFILES = os.path.join(ROOT, "profiles")


def _load(name):
    lines = [l for l in open(os.path.join(PROFILES, name)).read().splitlines() if l.strip()]
    assert len(lines) == 1, f"{name}: expected exactly one JSON line"
    return json.loads(lines[0])


def test_fnx_line_has_the_contract_keys():
    d = _load("r1_bench_fnx_n1.json")
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
              "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["unit"] == "iters/s" and d["higher_is_better"] is True and d["n_gpus"] == 1 and d["warmup"] >= 3 and d["vs_baseline"] is None
    assert d["dtype"] == "f32" and d["data"] == "synthetic" and "workload" in d["config"] and "model" not in d["config"]
    assert abs(d["value"] - d["config"]["frames_in_flight"] * d["steps"] / (d["ms_per_step"] * d["steps"] / 1e3)) < 1e-2 * d["value"]
    e = d["e2e"]
    assert e["unit"] == "iters/s" and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] <= d["value"] * 1.05
    # 16 frames x 5 views x 3 channels x 512 x 512 fp32 uploaded per step
    assert e["h2d_bytes_per_step"] == d["config"]["frames_in_flight"] * 5 * 3 * 512 * 512 * 4
    assert d["gpu_launches"] > 0
    r = d["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in r, k
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-4
    assert r["traffic"] is None or r["traffic"] > 0
    c = d["cpu_baseline"]
    assert c["kind"] in ("port", "reference") and c["cores"] >= 1 and c["value"] > 0 and c["unit"] == "iters/s" and c["sample"]
    k = d["clocks"]
    assert k["sm_mhz"] and k["sm_max_mhz"] and not set(k["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    s = d["static_tile_cache"]
    assert s["on"] is True and 0 < s["value_with_cache_off"] < d["value"]


def test_reference_line_has_the_contract_keys():
    d = _load("r1_bench_reference_n1.json")
    assert d["impl"] == "reference" and d["unit"] == "iters/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "iters/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    mine = _load("r1_bench_fnx_n1.json")
    assert d["metric"] == mine["metric"] and d["config"]["workload"] == mine["config"]["workload"]


@pytest.mark.parametrize("n", [1, 2, 4, 8])
def test_scaling_lines(n):
    d = _load(f"r1_bench_fnx_n{n}.scaling.json")
    assert d["n_gpus"] == n and d["scaling"] == "strong" and d["value"] > 0 and d["config"]["frames_in_flight"] == 16


def test_stdout_carries_only_the_emitted_line():
    """claim_stdout() sends everything printed afterwards to stderr; emit() writes to the real stdout."""
    code = ("import sys; sys.path.insert(0, %r); import bench; bench.claim_stdout(); print('library banner'); "
            "import os; os.system('echo child process banner'); bench.emit('{\"ok\": 1}')" % ROOT)
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert p.returncode == 0, p.stderr
    assert p.stdout == '{"ok": 1}\n'
    assert "library banner" in p.stderr and "child process banner" in p.stderr


# ---------------------------------------------------------------------------------------------------------------------
# round 2 lines
# ---------------------------------------------------------------------------------------------------------------------
def test_r2_fnx_line():
    d = _load("r2_bench_fnx_n1.json")
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
              "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline", "leg", "value_lanes1", "latency_one_frame_ms",
              "launches_per_iteration", "dropin_unchanged_python", "static_tile_cache"):
        assert k in d, k
    assert d["n_gpus"] == 1 and d["dtype"] == "f32" and d["vs_baseline"] is None and "model" not in d["config"]
    # every timed leg: >= 5 intervals, >= 2 s (the short legs >= 1 s), median reported
    assert d["leg"]["intervals"] >= 5 and d["leg"]["total_s"] >= 2.0 and d["e2e"]["leg"]["intervals"] >= 5 and d["e2e"]["leg"]["total_s"] >= 2.0
    assert abs(d["ms_per_step"] * d["steps"] - d["leg"]["median_ms"]) < 1e-2 * d["leg"]["median_ms"]
    assert abs(d["value"] - d["config"]["frames_in_flight"] * 1e3 / d["ms_per_step"]) < 1e-3 * d["value"]
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] == d["config"]["frames_in_flight"] * 5 * 3 * 512 * 512 * 4 and e["d2h_bytes_per_step"] == 4
    assert 0 < e["value"] <= d["value"] * 1.02 and e["with_gt_cache"]["value"] >= e["value"] and e["with_gt_cache"]["h2d_bytes_per_step"] == 0
    assert 0 < d["value_lanes1"] <= d["value"] * 1.02 and 0.3 < d["latency_one_frame_ms"] < 2.0
    r = d["roofline"]
    assert r["bound"] == "hbm" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-4 and r["traffic"] > 0
    assert r["issue_rate"]["capture"].startswith("profiles/r2_ncu_smoke_") and 0 < r["issue_rate"]["frac"] < 1
    assert os.path.exists(os.path.join(ROOT, r["issue_rate"]["capture"].split(" ")[0]))
    x = d["dropin_unchanged_python"]
    assert x["loop_body"].startswith("entries_fluid_nexus/train_physical_particle.py:") and x["with_accelerators"]["value"] > x["value"] > 0
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}


def test_r2_reference_line_runs_the_stock_loop():
    d = _load("r2_bench_reference_n1.json")
    mine = _load("r2_bench_fnx_n1.json")
    assert d["impl"] == "reference" and d["metric"] == mine["metric"] and d["config"]["workload"] == mine["config"]["workload"]
    assert d["config"]["glue"].startswith("stock: loop body of entries_fluid_nexus/train_physical_particle.py:")
    assert d["steps"] == d["config"]["steps_requested"] == mine["steps"] and d["config"]["steps_note"] is None
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1 and d["reference_gpu_part_ms_per_iteration"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": "iters/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


@pytest.mark.parametrize("wl", ["scalar", "c2", "ball"])
def test_r2_other_workloads_have_both_arms(wl):
    d, r = _load(f"r2_bench_{wl}_n1.json"), _load(f"r2_bench_reference_{wl}.json")
    assert d["config"]["workload"].startswith(wl + ":") and r["config"]["workload"] == d["config"]["workload"] and r["impl"] == "reference"
    assert d["roofline"]["traffic"] > 0 and d["roofline"]["issue_rate"]["capture"].startswith(f"profiles/r2_ncu_{wl}_")
    assert d["value"] > 100 * r["value"] > 0


@pytest.mark.parametrize("n", [2, 4, 8])
def test_r2_scaling_lines_carry_the_sharding_parity_check(n):
    d = _load(f"r2_bench_fnx_n{n}.json")
    assert d["n_gpus"] == n and d["scaling"] == "strong" and d["config"]["frames_in_flight"] == 16
    sp = d["sharding_parity"]
    assert sp["ok"] is True and sp["max_abs_param_delta"] <= sp["tolerance"] and sp["loss_rel_delta_last_step"] < 1e-4
    assert any(o != [0] for o in sp["owners"])            # frames that other ranks own were re-run on rank 0 and compared
    assert "no gradient exchange" in d["config"]["parallelism"]
    one = _load("r2_bench_fnx_n1.json")
    assert d["value"] > 0.8 * n * one["value"]            # strong-scaling efficiency of the device-resident step >= 0.8
    # per-rank sections shrink like 1 / N (the round-1 bug made them GROW with N)
    assert d["sections_ms_per_step"]["blend_bwd"] < 1.5 * one["sections_ms_per_step"]["blend_bwd"] / n


@pytest.mark.parametrize("n", [2, 4])
def test_r2_view_sharded_lines(n):
    d = _load(f"r2_bench_fnx_n{n}_views.json")
    assert d["n_gpus"] == n and d["config"]["frames_in_flight"] == 1 and f"split over {n} ranks" in d["config"]["parallelism"]
    sp = d["sharding_parity"]
    assert sp["ok"] is True and len(sp["owners"][0]) > 1 and sp["max_abs_param_delta"] <= sp["tolerance"]   # frame 0 lives on several ranks
    one = _load("r2_bench_fnx_n1_oneframe.json")
    assert d["value"] > one["value"]                       # splitting the views of one frame is faster than one GPU, if modestly
