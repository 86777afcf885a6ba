"""The C ABI from plain C: examples/pf_lgssm.c (gcc, no Python/torch in the process) runs BASELINE config 1 through
include/genb200.h and checks the particle-filter log-ML estimate against the exact Kalman value it computes itself."""
import json
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build_driver(tmp_path, name="pf_lgssm"):
    import __graft_entry__ as entry
    entry.build()
    exe = str(tmp_path / name)
    libdir = os.path.join(ROOT, "gen.jl_b200")
    subprocess.check_call(["gcc", "-O2", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", name + ".c"), "-o", exe, "-L" + libdir, "-lgenb200",
                           "-Wl,-rpath," + libdir, "-lm"])
    return exe


def test_c_driver_compiles_and_fails_loudly_without_gpu(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu test")
    exe = build_driver(tmp_path)
    p = subprocess.run([exe, "1000", "10", "2"], capture_output=True, text=True)
    assert p.returncode == 2 and "no CPU fallback" in p.stderr        # GB_ENOGPU, never a silent host path
    exe = build_driver(tmp_path, "pf_multigpu")
    p = subprocess.run([exe, "1000", "3", "2"], capture_output=True, text=True)
    assert p.returncode == 2 and "no CPU fallback" in p.stderr


@pytest.mark.gpu
def test_c_driver_matches_kalman_on_gpu(tmp_path):
    exe = build_driver(tmp_path)
    p = subprocess.run([exe, "1000", "100", "32"], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout + p.stderr
    out = json.loads(p.stdout.strip().splitlines()[-1])
    assert out["ok"] and out["resamples"] > 0 and abs(out["pf_mean_log_ml"] - out["kalman_log_ml"]) < 0.3


@pytest.mark.gpu
@pytest.mark.parametrize("shards", [2, 5])
def test_c_driver_multigpu_identical_to_one_shard(tmp_path, shards):
    """examples/pf_multigpu.c: the gb_spf_* entries from plain C.  G shards (on as many GPUs as the box has; on one GPU
    they share it) through the synchronous calls and through gb_spf_run == one shard, bit for bit."""
    exe = build_driver(tmp_path, "pf_multigpu")
    p = subprocess.run([exe, "150001", "12", str(shards)], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout + p.stderr
    out = json.loads(p.stdout.strip().splitlines()[-1])
    assert out["ok"] and out["traces_identical"] and out["resamples"][0] > 0
