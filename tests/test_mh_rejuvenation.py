"""MH rejuvenation moves on every particle's latest latent choices.

Reference semantics: src/inference/mh.jl:16-27 (selection / regenerate), :37-58 (custom proposal), used on particle
filter states as in docs/src/tutorials/smc.md:459-463, :518-520.  CPU tests pin the oracle's acceptance ratios on
independently computed target densities and on the exact posterior of the reference's HMM fixture; GPU tests compare
the CUDA kernel with the oracle (accept decisions and states bit-exact, ratios 1e-10)."""
import math

import numpy as np
import pytest

import fixtures as F

RTOL64 = 1e-10


def npdf(x, mu, sd):
    z = (x - mu) / sd
    return -(z * z + math.log(2 * math.pi)) / 2 - np.log(sd)


# ---------------------------------------------------------------- CPU: the oracle's moves
def test_oracle_drift_ratio_is_target_density_ratio(O):
    """Gaussian drift is symmetric, so alpha = log pi(x') - log pi(x), pi = p(x | parent) p(y | x)."""
    n = 500
    rng = np.random.default_rng(3)
    ys = F.lgssm_series(3)
    p = F.lgssm_params()
    pf = O.OraclePF.init(O.MODEL_LGSSM, p, n, [ys[0]], normals=rng.standard_normal((1, n)))
    pf.step([ys[1]], normals=rng.standard_normal((1, n)))
    x_old = pf.state(0)
    eps, ua = rng.standard_normal((1, n)), rng.random((1, n))
    # the parents x0: the same generate sweep replayed from the same seed
    rng2 = np.random.default_rng(3)
    parent = O.OraclePF.init(O.MODEL_LGSSM, p, n, [ys[0]], normals=rng2.standard_normal((1, n))).state(0)
    nacc, alpha, acc = pf.rejuvenate([ys[1]], move=1, drift_std=0.3, normals=eps, uniforms=ua)
    x_prop = x_old + 0.3 * eps[0]
    logpi = lambda x: npdf(x, p[2] * parent, p[3]) + npdf(ys[1], x, p[4])
    assert np.allclose(alpha, logpi(x_prop) - logpi(x_old), rtol=1e-9, atol=1e-9)
    assert np.array_equal(acc, np.log(ua[0]) < alpha) and nacc == acc.sum()
    assert np.array_equal(pf.state(0), np.where(acc, x_prop, x_old))
    # weights are untouched by mh (the reference never writes state.log_weights there)
    assert 0 < nacc < n


def test_oracle_resimulation_ratio_is_likelihood_ratio(O):
    n = 400
    rng = np.random.default_rng(4)
    zs = F.bearings_series(3)
    p = F.BEARINGS_PARAMS
    pf = O.OraclePF.init(O.MODEL_BEARINGS, p, n, [zs[0]], normals=rng.standard_normal((4, n)))
    lw0 = pf.log_weights
    old = np.stack([pf.state(f) for f in range(4)])
    eps, ua = rng.standard_normal((4, n)), rng.random((1, n))
    nacc, alpha, acc = pf.rejuvenate([zs[0]], move=0, normals=eps, uniforms=ua)
    new = np.stack([p[2 * k] + p[2 * k + 1] * eps[k] for k in range(4)])
    lik = lambda s: npdf(zs[0], np.arctan2(s[1], s[0]), p[9])
    assert np.allclose(alpha, lik(new) - lik(old), rtol=1e-9, atol=1e-9)
    assert np.array_equal(np.stack([pf.state(f) for f in range(4)]), np.where(acc, new, old))
    assert np.array_equal(pf.log_weights, lw0)
    pf.maybe_resample(ess_threshold=n + 1, u0=5)
    with pytest.raises(ValueError):                      # parents are gone after a resample
        pf.rejuvenate([zs[0]], move=0, normals=eps, uniforms=ua)


def test_oracle_hmm_resimulation_converges_to_exact_posterior(O):
    """The reference's HMM fixture (test/inference/particle_filter.jl:52-78): repeated mh(trace, select(:z)) at t = 0
    leaves every particle distributed as p(z0 | x0) = prior .* emission[x0, :] normalised."""
    n = 40000
    pf = O.OraclePF.init(O.MODEL_HMM, F.hmm_params(), n, [F.HMM_OBS[0]], seed=11)
    for k in range(12):
        pf.rejuvenate([F.HMM_OBS[0]], move=0, move_index=k)
    post = F.hmm_init_proposal(F.HMM_OBS[0])
    freq = np.bincount(pf.state(0).astype(int), minlength=4)[1:] / n
    assert np.abs(freq - post).max() < 5 * math.sqrt(0.25 / n)


# ---------------------------------------------------------------- GPU: CUDA kernel vs oracle
def zoo(g, O):
    return {
        "lgssm": (g.lgssm_model(**F.LGSSM), O.MODEL_LGSSM, F.lgssm_params(), F.lgssm_series(8)),
        "sv": (g.sv_model(**F.SV), O.MODEL_SV, F.sv_params(), F.sv_series(8)),
        "bearings": (g.bearings_model(), O.MODEL_BEARINGS, F.BEARINGS_PARAMS, F.bearings_series(8)),
        "hmm": (g.hmm_model(F.HMM_PRIOR, F.HMM_TRANSITION, F.HMM_EMISSION), O.MODEL_HMM, F.hmm_params(),
                np.array([1, 1, 2, 3, 3, 1, 2, 2], dtype=float)),
    }


def mh_rows(model, move, init, name):
    if move == "drift":
        return ({"bearings": 4 if init else 2}.get(name, 1), 1)
    return (model.num_normals(init), model.num_uniforms(init) + 1)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["lgssm", "sv", "bearings", "hmm"])
@pytest.mark.parametrize("move", ["resimulate", "drift"])
@pytest.mark.parametrize("lazy", [True, False])
def test_rejuvenation_predrawn_matches_oracle(g, O, name, move, lazy):
    if name == "hmm" and move == "drift":
        pytest.skip("drift needs continuous latents")
    model, okind, params, ys = zoo(g, O)[name]
    n = 2777
    rng = np.random.default_rng(hash((name, move)) % 2 ** 32)
    drift = {"lgssm": 0.3, "sv": 0.2, "bearings": 1e-3}.get(name, 0.0)
    omove = 0 if move == "resimulate" else 1
    st = g.create_particle_filter(model, n)
    st.set_lazy_gather(lazy)

    def noise(init):
        nn, nu = model.num_normals(init), model.num_uniforms(init)
        return (rng.standard_normal((max(nn, 1), n)) if nn else None), (rng.random((max(nu, 1), n)) if nu else None)

    def sweep(t):
        nn, nu = mh_rows(model, move, t == 0, name)
        e, u = rng.standard_normal((max(nn, 1), n)), rng.random((nu, n))
        st.set_noise(e if nn else None, u)
        nacc, alpha = g.rejuvenate_(st, ys[t], move=move, drift_std=drift, return_alpha=True)
        nacc_ref, alpha_ref, acc_ref = ref.rejuvenate([ys[t]], move=omove, drift_std=drift, normals=e if nn else None, uniforms=u)
        fin = np.isfinite(alpha_ref)
        assert np.array_equal(np.isfinite(alpha), fin)
        assert np.allclose(alpha[fin], alpha_ref[fin], rtol=RTOL64, atol=RTOL64 * 10)
        assert nacc == nacc_ref
        for f in range(model.state_dim):
            assert np.array_equal(st.get_state(f), ref.state(f)), (t, f)
        return nacc

    e, u = noise(True)
    st.set_noise(e, u)
    g.generate_(st, ys[0])
    ref = O.OraclePF.init(okind, params, n, [ys[0]], normals=e, uniforms=u)
    total = sweep(0) + sweep(0)
    for t in range(1, len(ys)):
        u0 = int(rng.integers(0, 2 ** 32))
        thr = n + 0.5 if t % 2 else n / 8
        st.set_resample_noise(u0=u0)
        assert g.maybe_resample_(st, ess_threshold=thr, resampler="systematic") == ref.maybe_resample(ess_threshold=thr, u0=u0)
        e, u = noise(False)
        st.set_noise(e, u)
        g.particle_filter_step_(st, (t,), (g.UnknownChange,), ys[t])
        ref.step([ys[t]], normals=e, uniforms=u)
        lw = st.log_weights
        total += sweep(t)
        if t % 3 == 0:
            total += sweep(t)
        assert np.array_equal(st.log_weights, lw)                    # mh never touches the weights
    assert 0 < total
    assert np.array_equal(st.parents, ref.parents)
    assert abs(g.log_ml_estimate(st) - ref.log_ml_estimate()) <= RTOL64 * abs(ref.log_ml_estimate())


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["lgssm", "bearings", "hmm"])
def test_rejuvenation_philox_mode_matches_oracle(g, O, name):
    """Production mode: both sides draw the move's noise from Philox block 0x800000 + 16 * move."""
    model, okind, params, ys = zoo(g, O)[name]
    n = 5000
    st = g.initialize_particle_filter(model, (0,), ys[0], n, seed=77)
    ref = O.OraclePF.init(okind, params, n, [ys[0]], seed=77)
    moves = ["resimulate"] if name == "hmm" else ["resimulate", "drift"]
    drift = 0.3 if name == "lgssm" else 1e-3
    for t in range(0, 4):
        if t > 0:
            assert g.maybe_resample_(st, ess_threshold=n + 0.5, resampler="systematic") == ref.maybe_resample(ess_threshold=n + 0.5)
            g.particle_filter_step_(st, (t,), (g.UnknownChange,), ys[t])
            ref.step([ys[t]])
        for k, mv in enumerate(moves * 2):
            nacc = g.rejuvenate_(st, ys[t], move=mv, drift_std=drift)
            nacc_ref, _, _ = ref.rejuvenate([ys[t]], move=0 if mv == "resimulate" else 1, drift_std=drift, move_index=k)
            # engine normals come from the table-driven log/sincos (<= 2 ulp from libm): states agree to rounding and
            # an accept decision can differ only for a draw within ~1e-13 of the ratio
            assert abs(nacc - nacc_ref) <= 1
            for f in range(model.state_dim):
                a, b = st.get_state(f), ref.state(f)
                assert np.mean(np.abs(a - b) <= 1e-12 * np.maximum(1.0, np.abs(b))) > 0.999
    assert abs(g.log_ml_estimate(st) - ref.log_ml_estimate()) < 1e-6


@pytest.mark.gpu
def test_rejuvenation_hmm_converges_to_exact_posterior_and_state_errors(g):
    n = 40000
    model = g.hmm_model(F.HMM_PRIOR, F.HMM_TRANSITION, F.HMM_EMISSION)
    st = g.initialize_particle_filter(model, (0,), F.HMM_OBS[0], n, seed=5)
    for _ in range(12):
        g.rejuvenate_(st, F.HMM_OBS[0])
    freq = np.bincount(st.get_state(0).astype(int), minlength=4)[1:] / n
    assert np.abs(freq - F.hmm_init_proposal(F.HMM_OBS[0])).max() < 5 * math.sqrt(0.25 / n)
    with pytest.raises(g.GenB200Error):                              # drift needs continuous latent choices
        g.rejuvenate_(st, F.HMM_OBS[0], move="drift", drift_std=0.1)
    assert g.maybe_resample_(st, ess_threshold=n + 0.5)
    with pytest.raises(g.GenB200Error, match="maybe_resample"):       # parents gone once a resample was decided
        g.rejuvenate_(st, F.HMM_OBS[0])
    g.particle_filter_step_(st, (1,), (g.UnknownChange,), F.HMM_OBS[1])
    assert g.rejuvenate_(st, F.HMM_OBS[1]) > 0


@pytest.mark.gpu
def test_rejuvenation_fp32_and_history(g, O):
    """fp32 filter: ratios to 1e-4 of the fp64 oracle; with history on, the latest snapshot follows the moved traces."""
    n = 4096
    rng = np.random.default_rng(9)
    ys = F.lgssm_series(4)
    model = g.lgssm_model(**F.LGSSM)
    st = g.create_particle_filter(model, n, dtype="f32")
    st.enable_history()
    e0 = rng.standard_normal((1, n)).astype(np.float32).astype(np.float64)
    st.set_noise(e0)
    g.generate_(st, ys[0])
    ref = O.OraclePF.init(O.MODEL_LGSSM, F.lgssm_params(), n, [ys[0]], normals=e0)
    e1 = rng.standard_normal((1, n)).astype(np.float32).astype(np.float64)
    st.set_noise(e1)
    g.particle_filter_step_(st, (1,), (g.UnknownChange,), ys[1])
    ref.step([ys[1]], normals=e1)
    em, um = rng.standard_normal((1, n)).astype(np.float32).astype(np.float64), rng.random((1, n))
    st.set_noise(em, um)
    nacc, alpha = g.rejuvenate_(st, ys[1], move="drift", drift_std=0.25, return_alpha=True)
    nacc_ref, alpha_ref, _ = ref.rejuvenate([ys[1]], move=1, drift_std=0.25, normals=em, uniforms=um)
    assert np.allclose(alpha, alpha_ref, rtol=1e-4, atol=2e-4)
    assert abs(nacc - nacc_ref) <= 0.002 * n + 2
    traj = st.trajectory(0)
    assert np.array_equal(traj[-1], st.get_state(0))


@pytest.mark.gpu
@pytest.mark.parametrize("name,move,n", [("bearings", "drift", 30011), ("lgssm", "resimulate", 30011), ("hmm", "resimulate", 2500),
                                         ("sv", "drift", 1000)])
def test_fused_run_with_rejuvenation_equals_api_loop(g, O, name, move, n):
    """gb_pf_run_mh: the tutorial's particle_filter_rejuv loop in one call == maybe_resample_ / step_ / rejuvenate_ x 2."""
    model, _, _, ys = zoo(g, O)[name]
    drift = {"bearings": 1e-3, "sv": 0.2}.get(name, 0.0)
    a = g.initialize_particle_filter(model, (0,), ys[0], n, seed=6)
    b = g.initialize_particle_filter(model, (0,), ys[0], n, seed=6)
    nres = nacc = 0
    for t in range(1, len(ys)):
        nres += g.maybe_resample_(a, ess_threshold=0.9 * n, resampler="systematic")
        g.particle_filter_step_(a, (t,), (g.UnknownChange,), ys[t], return_incr=False)
        for _ in range(2):
            nacc += g.rejuvenate_(a, ys[t], move=move, drift_std=drift)
    nres_b, nacc_b = g.particle_filter_run_(b, ys[1:], ess_threshold=0.9 * n, rejuvenation=(move, drift, 2))
    assert (nres, nacc) == (nres_b, nacc_b) and nacc > 0 and nres > 0
    for f in range(model.state_dim):
        assert np.array_equal(a.get_state(f), b.get_state(f))
    assert np.array_equal(a.parents, b.parents) and np.array_equal(a.log_weights, b.log_weights)
    assert a.log_ml_est == b.log_ml_est


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["bearings", "hmm"])
def test_spec_model_resimulation_equals_builtin_kind(g, O, name, monkeypatch):
    """GB_MODEL_SPEC programs get the resimulation move through their NVRTC-specialised Transition; the score of the current
    latents is carried per particle instead of recomputed.  Same Philox streams => same moves as the hand-written kind
    (the ratio differs by the rounding of a constant-folded log, so an accept can flip only on a measure-zero tie)."""
    from test_gpu_parity import bearings_spec, hmm_spec
    monkeypatch.setenv("GENB200_SPEC_JIT", "1")
    builtin, _, _, ys = zoo(g, O)[name]
    spec = bearings_spec(g) if name == "bearings" else hmm_spec(g)
    n = 20011
    a = g.initialize_particle_filter(builtin, (0,), ys[0], n, seed=3)
    b = g.initialize_particle_filter(spec, (0,), ys[0], n, seed=3)
    total = 0
    for t in range(0, 5):
        if t > 0:
            thr = n + 0.5 if t % 2 else n / 3
            assert g.maybe_resample_(a, ess_threshold=thr, resampler="systematic") == g.maybe_resample_(b, ess_threshold=thr, resampler="systematic")
            g.particle_filter_step_(a, (t,), (g.UnknownChange,), ys[t], return_incr=False)
            g.particle_filter_step_(b, (t,), (g.UnknownChange,), ys[t], return_incr=False)
        for _ in range(3):
            na, alpha_a = g.rejuvenate_(a, ys[t], return_alpha=True)
            nb, alpha_b = g.rejuvenate_(b, ys[t], return_alpha=True)
            fin = np.isfinite(alpha_a)
            assert np.array_equal(fin, np.isfinite(alpha_b)) and np.allclose(alpha_a[fin], alpha_b[fin], rtol=1e-10, atol=1e-9)
            assert abs(na - nb) <= 1
            total += nb
            for f in range(builtin.state_dim):
                assert np.mean(a.get_state(f) == b.get_state(f)) > 0.9999
    assert total > 0 and b.spec_backend == "nvrtc"
    assert np.array_equal(a.parents, b.parents)
    with pytest.raises(g.GenB200Error, match="resimulation move only"):
        g.rejuvenate_(b, ys[4], move="drift", drift_std=0.1)
    monkeypatch.setenv("GENB200_SPEC_JIT", "0")                       # interpreter backend: no specialised Transition
    c = g.initialize_particle_filter(spec, (0,), ys[0], 1000, seed=3)
    with pytest.raises(g.GenB200Error, match="NVRTC"):
        g.rejuvenate_(c, ys[0])


@pytest.mark.gpu
def test_spec_model_with_inline_proposal_has_no_resimulation_move(g):
    p = g.SpecProgram()
    x = p.sample_normal(p.const(0.0), p.const(1.0))
    p.observe_normal(p.obs(0), x, p.const(0.5))
    p.add_weight(p.neg(p.logpdf_normal(x, p.const(0.0), p.const(1.0))))
    p.store(0, x)
    st = g.initialize_particle_filter(g.spec_model(1, [], 1, p), (0,), 0.3, 1000)
    with pytest.raises(g.GenB200Error, match="custom proposal"):
        g.rejuvenate_(st, 0.3)


# ---------------------------------------------------------------- sliding-window rejuvenation (smc.md:492-541)
def _bearings_log_joint(p, traj, zs_window):
    """log p(latent choices of steps 1..L, observations | step 0) for a bearings trajectory [L+1][4][n] (numpy, independent
    of the oracle's C code): velocities ~ normal(previous velocity, sqrt(var)); positions accumulate; z ~ normal(atan2(y, x), noise)."""
    lj = 0.0
    sdv = math.sqrt(p[8])
    for j in range(1, len(traj)):
        lj = lj + npdf(traj[j][2], traj[j - 1][2], sdv) + npdf(traj[j][3], traj[j - 1][3], sdv)
        lj = lj + npdf(zs_window[j - 1], np.arctan2(traj[j][1], traj[j][0]), p[9])
    return lj


def test_oracle_window_move_is_joint_density_ratio(O):
    """Gaussian drift is symmetric, so the acceptance ratio of the block move on the velocities of steps a..b is the ratio of
    the log joint of the whole replayed trajectory a-1..T, positions recomputed from the perturbed velocities."""
    n, T, W = 300, 6, 5
    rng = np.random.default_rng(8)
    zs = F.bearings_series(T + 1)
    p = F.BEARINGS_PARAMS
    pf = O.OraclePF.init(O.MODEL_BEARINGS, p, n, [zs[0]], normals=rng.standard_normal((4, n)))
    pf.enable_window(W)
    snaps = [np.stack([pf.state(f) for f in range(4)])]
    for t in range(1, T + 1):
        pf.step([zs[t]], normals=rng.standard_normal((2, n)))
        snaps.append(np.stack([pf.state(f) for f in range(4)]))         # no resampling: lineage = identity
    a, b, sd = 3, 5, 1e-3
    nb = b - a + 1
    eps, ua = rng.standard_normal((2 * nb, n)), rng.random((1, n))
    nacc, alpha, acc = pf.rejuvenate_window(a, b, zs[a:T + 1], sd, normals=eps, uniforms=ua)
    old = [snaps[t] for t in range(a - 1, T + 1)]
    new = [old[0]]
    for j in range(1, len(old)):
        vx, vy = old[j][2].copy(), old[j][3].copy()
        if j <= nb:
            vx = vx + sd * eps[2 * (j - 1)]
            vy = vy + sd * eps[2 * (j - 1) + 1]
        new.append(np.stack([new[j - 1][0] + vx, new[j - 1][1] + vy, vx, vy]))
    ref = _bearings_log_joint(p, new, zs[a:T + 1]) - _bearings_log_joint(p, old, zs[a:T + 1])
    assert np.allclose(alpha, ref, rtol=1e-8, atol=1e-8)
    assert np.array_equal(acc, np.log(ua[0]) < alpha) and nacc == acc.sum() and 0 < nacc < n
    assert np.allclose(pf.state(0), np.where(acc, new[-1][0], old[-1][0]), rtol=0, atol=1e-15)
    with pytest.raises(ValueError):
        pf.rejuvenate_window(1, 2, zs[1:T + 1], sd)                     # step 0 lies outside a 5-deep window at T = 6


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["bearings", "lgssm", "sv"])
@pytest.mark.parametrize("dtype", ["f64", "f32"])
def test_window_rejuvenation_matches_oracle(g, O, name, dtype):
    """The tutorial's particle_filter_rejuv loop (smc.md:513-531): block drift moves on the last five steps before every
    resample.  Pre-drawn noise on both sides: accept decisions, traces (current AND inside the window, checked through the
    later moves that read them) and ancestors bit-exact, log acceptance ratios to 1e-10; then the Philox-driven move."""
    from test_gpu_parity import model_zoo, _series
    model, okind_m, params, _, _, _ = model_zoo(g, O)[name]
    n, T, W, sd = 6007, 12, 5, (1e-3 if name == "bearings" else 0.2)
    ys = _series(name, T + 1)
    rng = np.random.default_rng(5)
    nn0, nn1 = model.num_normals(True), model.num_normals(False)
    nd = 2 if name == "bearings" else 1
    f32 = dtype == "f32"
    rnd = (lambda *s: rng.standard_normal(s).astype(np.float32).astype(np.float64)) if f32 else (lambda *s: rng.standard_normal(s))
    st = g.create_particle_filter(model, n, dtype=dtype, seed=3)
    st.enable_window(W)
    eps = rnd(nn0, n)
    st.set_noise(eps)
    g.generate_(st, ys[0])
    O.lib().go_set_storage_f32(1 if f32 else 0)
    try:
        ref = O.OraclePF.init(okind_m, params, n, [ys[0]], seed=3, normals=eps)
    finally:
        O.lib().go_set_storage_f32(0)
    ref.enable_window(W)
    tol = 1e-5 if f32 else RTOL64
    total = 0
    for t in range(1, T + 1):
        eps = rnd(nn1, n)
        st.set_noise(eps)
        g.particle_filter_step_(st, (t,), (g.UnknownChange,), ys[t], return_incr=False)
        ref.step([ys[t]], normals=eps)
        a, b = max(1, t - 4), t                                          # the last five steps (smc.md:524)
        nb = b - a + 1
        e2 = rnd(nb * nd, n)
        ua = rng.random((1, n)).astype(np.float32).astype(np.float64) if f32 else rng.random((1, n))
        st.set_noise(e2, ua)
        nacc, alpha = g.rejuvenate_window_(st, a, b, ys[a:t + 1], sd, return_alpha=True)
        nacc_ref, alpha_ref, acc_ref = ref.rejuvenate_window(a, b, ys[a:t + 1], sd, normals=e2, uniforms=ua)
        assert nacc == nacc_ref, (t, nacc, nacc_ref)
        assert np.allclose(alpha, alpha_ref, rtol=tol, atol=tol * 10), (t, np.abs(alpha - alpha_ref).max())
        for f in range(model.state_dim):
            assert np.array_equal(st.get_state(f), ref.state(f)), (t, f)
        total += nacc
        if t == T and not f32:
            # production noise (Philox): same counts / traces as the oracle driven by the same counter-based stream
            nacc = g.rejuvenate_window_(st, T - 2, T, ys[T - 2:T + 1], sd)
            nacc_ref, _, _ = ref.rejuvenate_window(T - 2, T, ys[T - 2:T + 1], sd, move_index=1)
            assert nacc == nacc_ref and nacc > 0
            # (the engine's Box-Muller uses its own table-driven log / sincos: same Philox integers, normals equal to ~1e-16)
            assert np.allclose(st.get_state(0), ref.state(0), rtol=0, atol=1e-12)
        u0 = 777 + t
        st.set_resample_noise(u0=u0)
        did = g.maybe_resample_(st, ess_threshold=n / 2 if t % 2 else n, resampler="systematic")
        assert did == ref.maybe_resample(ess_threshold=n / 2 if t % 2 else n, kind=O.RESAMPLE_SYSTEMATIC, u0=u0)
        if did:
            assert np.array_equal(st.parents, ref.parents)
    assert 0 < total < T * n
    # reaching back beyond the window is refused, not approximated
    with pytest.raises(g.GenB200Error):
        g.rejuvenate_window_(st, T - 6, T, ys[T - 6:T + 1], sd)
