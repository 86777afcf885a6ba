"""The bench.py JSON contract, checked on the committed lines under profiles/ (CPU) and on a live reference-arm run."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e"}


def check_line(d, reference=False):
    assert BASE_KEYS <= set(d), BASE_KEYS - set(d)
    assert d["metric"] == "particle-steps/sec" and d["unit"] == "particle-steps/s" and d["higher_is_better"] is True
    assert d["vs_baseline"] is None and d["data"] == "synthetic" and d["dtype"] in ("f64", "f32")
    assert "workload" in d["config"] and "model" not in d["config"]
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(d["e2e"])
    assert d["value"] > 0 and d["ms_per_step"] > 0
    if reference:
        assert d["impl"] == "reference" and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
        assert {"value", "unit", "cores", "kind", "sample"} <= set(d["cpu_baseline"]) and d["cpu_baseline"]["kind"] == "port"
    else:
        assert d["gpu_launches"] > 0 and {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
        r = d["roofline"]
        assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r) and r["bound"] == "hbm" and r["unit"] == "GB/s"
        assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9


@pytest.mark.parametrize("name", ["r1g_bench_1gpu.json", "r1g_bench_2gpu_weak.json", "r1g_bench_8gpu_weak.json"])
def test_committed_bench_lines_follow_the_contract(name):
    d = json.load(open(os.path.join(ROOT, "profiles", name)))
    check_line(d)
    if d["n_gpus"] == 1:
        assert {"value", "unit", "cores", "kind", "sample"} <= set(d["cpu_baseline"])
        assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0


def test_reference_arm_runs_on_the_host_and_prints_one_json_line():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
                        "--ref-particles", "200000"], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1                                   # exactly one JSON line on stdout
    check_line(json.loads(lines[0]), reference=True)
