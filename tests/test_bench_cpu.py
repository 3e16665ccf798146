"""bench.py on a box without a GPU: the reference arm (the reference's CPU path on the host cores), the
configuration table, and the no-CPU-fallback rule of the product arm."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BENCH = os.path.join(ROOT, "bench.py")


def run_bench(*args, env=None):
    e = dict(os.environ)
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
        e.pop(k, None)
    e.update(env or {})
    return subprocess.run([sys.executable, BENCH, *args], capture_output=True, text=True, env=e, timeout=600)


def test_config_table_matches_baseline_json():
    sys.path.insert(0, ROOT)
    import bench
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert bench.METRIC.split()[0] in base["metric"]
    # (cols, rows, slice length) of BASELINE.json configs[1..4]
    want = {"cfg2": (240, 180, 0.030), "cfg3": (346, 260, 0.050), "cfg4": (640, 480, 0.020), "cfg5": (1280, 720, 0.010)}
    for name, (cols, rows, slice_s) in want.items():
        bench.select_config(name)
        assert (bench.SENSOR_COLS, bench.SENSOR_ROWS, bench.SLICE_S) == (cols, rows, slice_s)
        assert "%dx%d" % (cols, rows) in bench.WORKLOAD
        assert bench.MAX_ITER == -1 and bench.SCALE == 3                      # GD to convergence, reference default scale
        assert bench.RATE_EPS * bench.SLICE_S * bench.SLICES_PER_STEP * 8 > 50e6   # tens of MB of events per step
    bench.select_config("cfg2", slices=7)
    assert bench.SLICES_PER_STEP == 7 and "7 slices per step" in bench.WORKLOAD
    bench.select_config("cfg2")


def test_reference_arm_prints_one_json_line():
    r = run_bench("--impl", "reference", "--steps", "1", "--warmup", "0")
    assert r.returncode == 0, r.stderr[-500:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "Mevents/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 1 and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"].startswith("DAVIS-240C")


def test_reference_arm_other_ranks_do_no_work():
    r = run_bench("--impl", "reference", "--steps", "1", "--warmup", "0", "--gpus", "2",
                  env={"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_one_core_child_mode():
    r = run_bench("--config", "cfg3", "--cpu-one-core", "1", env={"BF_ORACLE_THREADS": "1"})
    assert r.returncode == 0, r.stderr[-500:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d["events"] > 50000 and d["seconds"] > 0 and d["cores"] == 1


def test_product_arm_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = run_bench("--steps", "1")
    assert r.returncode != 0
    assert "no CPU fallback" in (r.stderr + r.stdout)
