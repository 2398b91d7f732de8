"""bench.py's one-JSON-line contract: the reference arm (CPU, runs here) and the keys of the product arm (GPU)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config"}


def run_bench(*args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, env=e, timeout=900)
    assert p.returncode == 0, p.stderr[-2000:]
    return p.stdout


def test_reference_arm_prints_one_json_line():
    out = run_bench("--impl", "reference", "--steps", "1", "--warmup", "3", "--ref-n", "32")
    lines = [l for l in out.splitlines() if l.strip()]
    assert len(lines) == 1, out
    d = json.loads(lines[0])
    assert BASE_KEYS <= set(d)
    assert d["impl"] == "reference" and d["dtype"] == "f64" and d["higher_is_better"] is True
    assert d["metric"] == "dynamics element-subcycle updates/s (FP64)" and d["unit"] == "element-subcycles/s"
    assert d["config"]["workload"] == "mevp_rect2048x2048_dg2cg2_nsteps100"  # the product arm's workload, sampled
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["value"] > 0 and d["cpu_baseline"]["cores"] >= 1


def test_reference_arm_other_ranks_exit_without_work():
    out = run_bench("--impl", "reference", "--gpus", "2", "--steps", "1", "--ref-n", "32", env={"RANK": "1", "WORLD_SIZE": "2"})
    assert out.strip() == ""


def test_reference_arm_follows_rheology_and_mesh():
    d = json.loads(run_bench("--impl", "reference", "--steps", "1", "--ref-n", "32", "--rheology", "bbm", "--mesh", "distorted"))
    assert d["config"]["workload"] == "bbm_para_distorted2048x2048_dg2cg2_nsteps100"


@pytest.mark.gpu
def test_product_arm_json_line_small_grid():
    out = run_bench("--n", "256", "--steps", "2", "--warmup", "3", "--cpu-n", "32", "--e2e-steps", "1")
    lines = [l for l in out.splitlines() if l.strip()]
    assert len(lines) == 1, out
    d = json.loads(lines[0])
    assert BASE_KEYS | {"roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"} <= set(d)
    assert d["gpu_launches"] > 0 and d["value"] > 0 and d["e2e"]["value"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] == 7 * 256 * 256 * 8 and d["e2e"]["d2h_bytes_per_step"] == 6 * 256 * 256 * 8
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    assert d["cpu_baseline"]["kind"] in ("reference", "port")


def test_committed_ncu_traffic_covers_the_bench_workloads():
    """bench.py's roofline.traffic comes from profiles/ncu_traffic.json, keyed by the workload name of the default runs."""
    import argparse
    import importlib.util

    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    table = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
    for rheo in ("mevp", "bbm"):
        name = bench.workload_name(argparse.Namespace(rheology=rheo, n=2048))
        assert name in table and table[name]["dram_bytes_per_launch"] > 1e9, name
        src = table[name]["source"].split(" ")[0]
        assert os.path.exists(os.path.join(ROOT, src)), src
