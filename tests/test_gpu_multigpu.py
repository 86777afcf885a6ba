"""gb_spf_*: one filter sharded over several GPUs from one process (csrc/spf.inc) against the single-GPU filter and the
oracle.  On a one-GPU box the shards share the device (devices = [0, 0, ...]): same kernels, same peer-memory exchange
(mailboxes, peer stores, flags), only the links are local.  With >= 2 GPUs the same tests also run with one shard per
device, and tests/run_sharded_multiproc.py checks the one-process-per-GPU flavour (CUDA IPC) under torch.distributed.run."""
import math
import os
import subprocess
import sys

import numpy as np
import pytest

import fixtures as F

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def device_lists(g):
    n = g.device_count()
    out = [[0], [0, 0], [0, 0, 0], [0] * 8]
    if n >= 2:
        out += [list(range(min(n, 8))), [0, 1, 1, 0]]
    return out


def mg(g):
    from genb200 import multigpu
    return multigpu


@pytest.mark.parametrize("dtype", ["f64", "f32"])
def test_spf_sync_api_equals_single_gpu(g, dtype):
    M = mg(g)
    model, ys = g.bearings_model(), F.bearings_series(10)
    n, seed = 70003, 41
    ref = g.initialize_particle_filter(model, (0,), ys[0], n, dtype=dtype, seed=seed)
    dids = []
    for t in range(1, len(ys)):
        dids.append(g.maybe_resample_(ref, ess_threshold=(n + 1 if t % 2 else n / 2), resampler="systematic"))
        g.particle_filter_step_(ref, (t,), (g.UnknownChange,), ys[t], return_incr=False)
    lml_ref, cols_ref, par_ref, lw_ref = g.log_ml_estimate(ref), [ref.get_state(f) for f in range(4)], ref.parents, ref.log_weights
    for devs in device_lists(g):
        st = M.initialize_particle_filter(model, (0,), ys[0], n, dtype=dtype, seed=seed, devices=devs)
        assert st.num_shards == len(devs)
        for t in range(1, len(ys)):
            assert M.maybe_resample_(st, ess_threshold=(n + 1 if t % 2 else n / 2)) == dids[t - 1], (devs, t)
            M.particle_filter_step_(st, (t,), (g.UnknownChange,), ys[t])
            if t == 4:       # statistics are cached: estimate twice, then resample, costs one pass
                assert M.log_ml_estimate(st) == M.log_ml_estimate(st)
        for f in range(4):
            assert np.array_equal(st.get_state(f), cols_ref[f]), (devs, f)
        assert np.array_equal(st.parents, par_ref), devs
        assert np.array_equal(st.log_weights, lw_ref), devs
        lml = M.log_ml_estimate(st)
        assert abs(lml - lml_ref) <= 1e-12 * abs(lml_ref), (devs, lml, lml_ref)
        assert abs(st.ess() - ref.ess()) <= 1e-9 * ref.ess()
        del st


@pytest.mark.parametrize("name", ["bearings", "sv", "hmm"])
def test_spf_run_equals_single_gpu_run(g, name):
    M = mg(g)
    if name == "bearings":
        model, ys = g.bearings_model(), F.bearings_series(12)
    elif name == "sv":
        model, ys = g.sv_model(**F.SV), F.sv_series(12)
    else:
        model, ys = g.hmm_model(F.HMM_PRIOR, F.HMM_TRANSITION, F.HMM_EMISSION), np.array([1, 1, 2, 3, 3, 1, 2, 2, 1, 3, 1, 2], dtype=float)
    n, seed = 90001, 8
    for thr in (n / 2, n + 1):
        ref = g.initialize_particle_filter(model, (0,), ys[0], n, seed=seed)
        nres_ref = g.particle_filter_run_(ref, ys[1:], ess_threshold=thr)
        for devs in device_lists(g)[1:]:
            st = M.initialize_particle_filter(model, (0,), ys[0], n, seed=seed, devices=devs)
            assert M.particle_filter_run_(st, ys[1:6], ess_threshold=thr) + M.particle_filter_run_(st, ys[6:], ess_threshold=thr) == nres_ref
            for f in range(model.state_dim):
                assert np.array_equal(st.get_state(f), ref.get_state(f)), (devs, f)
            assert np.array_equal(st.parents, ref.parents) and np.array_equal(st.log_weights, ref.log_weights)
            assert abs(M.log_ml_estimate(st) - g.log_ml_estimate(ref)) <= 1e-12 * abs(g.log_ml_estimate(ref))
            assert np.array_equal(st.log_incremental_weights, ref.log_incremental_weights)
            del st


def test_spf_matches_oracle(g, O):
    """The sharded filter against the sequential oracle directly (exported Philox noise), not only against one GPU."""
    M = mg(g)
    model, ys = g.bearings_model(), F.bearings_series(8)
    n, seed = 40009, 3
    st = M.initialize_particle_filter(model, (0,), ys[0], n, seed=seed, devices=[0, 0, 0])
    ref = O.OraclePF.init(O.MODEL_BEARINGS, F.BEARINGS_PARAMS, n, [ys[0]], normals=g.philox_normals(seed, 0, 0, n, 4))
    for t in range(1, len(ys)):
        thr = n + 1 if t % 2 == 0 else n / 2
        did = M.maybe_resample_(st, ess_threshold=thr)
        assert did == ref.maybe_resample(ess_threshold=thr, kind=O.RESAMPLE_SYSTEMATIC, u0=g.philox_resample_u0(seed, t - 1))
        M.particle_filter_step_(st, (t,), (g.UnknownChange,), ys[t])
        ref.step([ys[t]], normals=g.philox_normals(seed, t, 0, n, 2))
    for f in range(4):
        assert np.array_equal(st.get_state(f), ref.state(f))
    assert np.array_equal(st.parents, ref.parents)
    assert np.allclose(st.log_weights, ref.log_weights, rtol=1e-10, atol=1e-10)
    assert abs(M.log_ml_estimate(st) - ref.log_ml_estimate()) <= 1e-10 * abs(ref.log_ml_estimate())


def test_spf_importance_sampling_blr(g, O):
    """cfg 4 sharded: only the log-sum-exp crosses the shards (importance.jl:20-36)."""
    M = mg(g)
    X, y, sigma = F.blr_problem(D=64, n=64)
    model = g.blr_model(X, sigma)
    n = 30011
    tr1, lnw1, lml1 = g.importance_sampling(model, (), y, n, seed=5)
    for devs in device_lists(g)[1:]:
        st, lnw, lml = M.importance_sampling(model, (), y, n, seed=5, devices=devs)
        assert abs(lml - lml1) <= 1e-12 * abs(lml1), devs
        assert np.allclose(lnw, lnw1, rtol=0, atol=1e-9)
        assert abs(g.logsumexp(lnw)) < 1e-9                       # importance_sampling.jl:18-34
        assert np.array_equal(st.get_state(63), tr1.column(63))


@pytest.mark.parametrize("shape", ["normal", "onehot"])
def test_spf_raw_resample_equals_oracle(g, O, shape):
    """cfg 5 sharded: global-prefix systematic resampling of a raw weight array over K shards == the sequential oracle,
    as long as no shard has to take in more rows than its extension rows hold (one-hot weights: loud failure)."""
    M = mg(g)
    n = 100003
    lw = F.resample_logw(n, shape)
    exp = O.resample(O.RESAMPLE_SYSTEMATIC, lw, u0=999)
    for devs in device_lists(g):
        if shape == "onehot" and len(devs) > 1:
            with pytest.raises(g.GenB200Error):
                M.resample(lw, u0=999, devices=devs)
            continue
        assert np.array_equal(M.resample(lw, u0=999, devices=devs), exp), devs


def test_spf_argument_errors(g):
    M = mg(g)
    with pytest.raises(g.GenB200Error):
        M.initialize_particle_filter(g.lgssm_model(), (0,), 0.1, 100, devices=[0, 99])
    with pytest.raises(g.GenB200Error):
        M.initialize_particle_filter(g.lgssm_model(), (0,), 0.1, 3, devices=[0] * 4)      # fewer particles than shards
    st = M.initialize_particle_filter(g.hmm_model(F.HMM_PRIOR, F.HMM_TRANSITION, F.HMM_EMISSION), (0,), 1.0, 5000, devices=[0, 0])
    with pytest.raises(g.GenB200Error):                              # checked before the first launch: the handle stays usable
        M.particle_filter_run_(st, [1.0, 2.0, 7.0])
    assert M.particle_filter_run_(st, [1.0, 2.0, 3.0]) >= 0


def test_multiprocess_ipc_sharded_filter(g):
    """One process per GPU (torch.distributed.run, CUDA IPC peer pointers, no collective in the step): bit-identical to the
    single-GPU filter.  Needs >= 2 GPUs; the same script is what bench.py's sharded_parity object runs at --gpus N."""
    n = g.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus 2)")
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(min(n, 4)),
                        "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "run_sharded_multiproc.py")],
                       capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    assert "MISMATCH" not in p.stdout and p.stdout.count("OK") >= 4
