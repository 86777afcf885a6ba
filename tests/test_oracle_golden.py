"""Pins the CPU oracle (oracle/gen_oracle.c) against every known-answer test the reference holds
for this path.  Gen.jl cannot run here (no Julia), so these KATs -- taken from
the reference's own test files, cited per test -- are what anchors the oracle."""
import math

import numpy as np

import fixtures as F


def test_normal_logpdf_overflow_kat(O):
    # test/modeling_library/distributions.jl:109: logpdf(normal, 1e13, 0, 1) == -5e25 exactly
    assert O.lib().go_logpdf_normal(1e13, 0.0, 1.0) == -5e25
    # closed form (normal.jl:56-59)
    assert O.lib().go_logpdf_normal(0.3, -0.2, 1.7) == -(((0.3 + 0.2) / 1.7) ** 2 + math.log(2 * math.pi)) / 2 - math.log(1.7)


def test_out_of_support_kats(O):
    # test/modeling_library/distributions.jl:50: logpdf(categorical, -1, [0.2, 0.3, 0.5]) == -Inf
    assert O.logpdf_categorical(-1, [0.2, 0.3, 0.5]) == -math.inf
    assert O.logpdf_categorical(4, [0.2, 0.3, 0.5]) == -math.inf
    assert O.logpdf_categorical(2, [0.2, 0.3, 0.5]) == math.log(0.3)
    # :73: logpdf(gamma, -1, 1, 1) == -Inf
    assert O.lib().go_logpdf_gamma(-1.0, 1.0, 1.0) == -math.inf
    # gamma.jl:10-16 closed form against scipy
    from scipy import stats
    for x, k, th in [(0.7, 2.5, 1.3), (3.0, 0.5, 2.0), (1e-3, 1.0, 1.0)]:
        assert abs(O.lib().go_logpdf_gamma(x, k, th) - stats.gamma.logpdf(x, k, scale=th)) < 1e-12
    # bernoulli.jl:10-12
    assert O.lib().go_logpdf_bernoulli(1, 0.3) == math.log(0.3)
    assert O.lib().go_logpdf_bernoulli(0, 0.3) == math.log(1.0 - 0.3)


def test_mvnormal_closed_form(O):
    # mvnormal.jl:12-16 (no value KAT in the reference: pinned against scipy's closed form)
    from scipy import stats
    rng = np.random.default_rng(0)
    A = rng.standard_normal((5, 5))
    cov = A @ A.T + 5 * np.eye(5)
    mu, x = rng.standard_normal(5), rng.standard_normal(5)
    assert abs(O.logpdf_mvnormal(x, mu, cov) - stats.multivariate_normal.logpdf(x, mu, cov)) < 1e-12


def test_hmm_forward_hand_example(O):
    # test/inference/particle_filter.jl:29-48
    prior = np.array([0.4, 0.6])
    em = np.array([[0.1, 0.9], [0.7, 0.3]]).T
    tr = np.array([[0.5, 0.5], [0.2, 0.8]]).T
    obs = [2, 1]
    exp = 0.0
    for z1 in (1, 2):
        for z2 in (1, 2):
            exp += prior[z1 - 1] * tr[z2 - 1, z1 - 1] * em[obs[0] - 1, z1 - 1] * em[obs[1] - 1, z2 - 1]
    assert math.isclose(O.hmm_forward(prior, em, tr, obs), exp, rel_tol=1e-14)


def test_hmm_forward_fixture_value(O):
    # test/inference/particle_filter.jl:52-81 fixture; SURVEY.md section 4: log ML = -4.87645083351704
    ml = O.hmm_forward(F.HMM_PRIOR, F.HMM_EMISSION, F.HMM_TRANSITION, F.HMM_OBS)
    assert abs(ml - 0.007624025) < 1e-9
    assert abs(math.log(ml) - F.HMM_EXPECTED_LOG_ML) < 1e-12


def _run_hmm_pf(O, custom, n, seed, resampler):
    obs = F.HMM_OBS
    params = F.hmm_params()
    if custom:
        pf = O.OraclePF.init(O.MODEL_HMM, params, n, [obs[0]], seed=seed, proposal_kind=O.PROPOSAL_CATEGORICAL_TABLE,
                             pargs=F.hmm_init_proposal(obs[0]))
    else:
        pf = O.OraclePF.init(O.MODEL_HMM, params, n, [obs[0]], seed=seed)
    for T in range(2, len(obs) + 1):
        pf.maybe_resample(ess_threshold=n, kind=resampler)   # ess_threshold = N exercises resampling (:100)
        if custom:
            incr = pf.step([obs[T - 1]], proposal_kind=O.PROPOSAL_CATEGORICAL_TABLE,
                           pargs=F.hmm_step_proposal(obs[T - 1]).ravel(order="F"))
        else:
            incr = pf.step([obs[T - 1]])
        assert len(incr) == n                                  # :137
    return pf.log_ml_estimate()


def test_pf_hmm_custom_proposal_reference_tolerance(O):
    # test/inference/particle_filter.jl:96-144: N = 20000, atol = 0.02
    for resampler in (O.RESAMPLE_MULTINOMIAL, O.RESAMPLE_SYSTEMATIC, O.RESAMPLE_RESIDUAL):
        lml = _run_hmm_pf(O, True, 20000, 1, resampler)
        assert abs(lml - F.HMM_EXPECTED_LOG_ML) < 0.02, (resampler, lml)


def test_pf_hmm_default_proposal_reference_tolerance(O):
    # test/inference/particle_filter.jl:146-170
    for resampler in (O.RESAMPLE_MULTINOMIAL, O.RESAMPLE_SYSTEMATIC, O.RESAMPLE_RESIDUAL):
        lml = _run_hmm_pf(O, False, 20000, 1, resampler)
        assert abs(lml - F.HMM_EXPECTED_LOG_ML) < 0.02, (resampler, lml)


def test_importance_sampling_invariants(O):
    # test/inference/importance_sampling.jl:18-34: N = 4, logsumexp(log_norm_weights) ~ 0 @ 1e-14
    pf, lnw, lml = O.importance_sampling(O.MODEL_LGSSM, F.lgssm_params(), 4, [0.3], seed=3)
    assert len(lnw) == 4 and not math.isnan(lml)
    assert abs(O.logsumexp(lnw)) < 1e-14
    pf, lnw, lml = O.importance_sampling(O.MODEL_LGSSM, F.lgssm_params(), 4, [0.3], seed=3,
                                         proposal_kind=O.PROPOSAL_AFFINE_NORMAL, pargs=[0.1, 0.0, 0.8])
    assert abs(O.logsumexp(lnw)) < 1e-14 and not math.isnan(lml)


def test_custom_proposal_weight_algebra(O):
    # test/inference/trace_translators.jl:58-69: log_weight == update_weight - proposal score, exactly,
    # for a normal-normal model extended by one step
    p = F.lgssm_params()
    n = 16
    rng = np.random.default_rng(5)
    eps0, eps1 = rng.standard_normal((1, n)), rng.standard_normal((1, n))
    pf = O.OraclePF.init(O.MODEL_LGSSM, p, n, [0.2], normals=eps0)
    x0 = pf.state(0)
    pa = [0.05, 0.7, 0.6]
    incr = pf.step([-0.4], proposal_kind=O.PROPOSAL_AFFINE_NORMAL, pargs=pa, normals=eps1)
    x1 = pf.state(0)
    L = O.lib()
    for i in range(n):
        qmu = pa[0] + pa[1] * x0[i]
        assert x1[i] == qmu + pa[2] * eps1[0, i]                       # normal.jl:128, two roundings
        update_w = 0.0
        update_w += L.go_logpdf_normal(x1[i], p[2] * x0[i], p[3])
        update_w += L.go_logpdf_normal(-0.4, x1[i], p[4])
        assert incr[i] == update_w - L.go_logpdf_normal(x1[i], qmu, pa[2])


def test_generate_weight_is_sum_of_constrained_logpdfs(O):
    # test/modeling_library/unfold.jl:41-70, map.jl:30-39, static_ir.jl:89-106: only constrained choices
    n = 8
    rng = np.random.default_rng(2)
    eps = rng.standard_normal((4, n))
    pf = O.OraclePF.init(O.MODEL_BEARINGS, F.BEARINGS_PARAMS, n, [1.55], normals=eps)
    L = O.lib()
    lw = pf.log_weights
    for i in range(n):
        x, y = pf.state(0)[i], pf.state(1)[i]
        assert x == 0.01 + 0.01 * eps[0, i] and y == 0.95 + 0.01 * eps[1, i]
        assert lw[i] == L.go_logpdf_normal(1.55, math.atan2(y, x), 0.005)


def test_logsumexp_edge_cases(O):
    # inference.jl:3-11
    assert O.logsumexp([-math.inf, -math.inf]) == -math.inf
    assert O.lib().go_logsumexp2(-math.inf, -math.inf) == -math.inf
    a = np.array([1.0, 2.0, 3.0])
    assert abs(O.logsumexp(a) - math.log(np.exp(a).sum())) < 1e-15
    assert abs(O.lib().go_logsumexp2(1.0, 2.0) - math.log(math.e + math.e ** 2)) < 1e-15
    # particle_filter.jl:3-12
    lse, lnw = O.normalize_weights(a)
    w = np.exp(lnw)
    assert abs(O.effective_sample_size(lnw) - 1.0 / (w ** 2).sum()) < 1e-12
    # all -Inf: normalised weights are NaN, `ess < thr` is false => no resample
    pf = O.OraclePF.init(O.MODEL_LGSSM, F.lgssm_params(), 4, [0.0], seed=1)
    pf.log_weights = [-math.inf] * 4
    assert pf.maybe_resample(ess_threshold=4) is False
    assert pf.log_ml_estimate() == -math.inf


def test_pf_state_copy_semantics(O):
    # test/inference/particle_filter.jl:172-198
    pf = O.OraclePF.init(O.MODEL_LGSSM, F.lgssm_params(), 32, [0.1], seed=9)
    c = pf.copy()
    assert np.array_equal(c.log_weights, pf.log_weights) and np.array_equal(c.parents, pf.parents)
    pf.step([0.2])
    assert not np.array_equal(c.log_weights, pf.log_weights)
    assert np.array_equal(c.parents, np.arange(1, 33))


def test_kalman_oracle_and_pf_agree(O):
    # cfg 1: PF log-ML vs exact Kalman value (north_star "Kalman-exact")
    ys = F.lgssm_series(31)
    kf = O.kalman_logml(0.0, 1.0, 0.9, 0.5, 1.0, ys)
    # brute-force check of the Kalman recursion on T = 2 by quadrature
    from scipy import integrate, stats
    y2 = ys[:2]
    f = lambda x1, x0: (stats.norm.pdf(x0, 0, 1) * stats.norm.pdf(y2[0], x0, 1.0) * stats.norm.pdf(x1, 0.9 * x0, 0.5)
                        * stats.norm.pdf(y2[1], x1, 1.0))
    val, _ = integrate.dblquad(f, -8, 8, -8, 8, epsabs=1e-12)
    assert abs(math.log(val) - O.kalman_logml(0.0, 1.0, 0.9, 0.5, 1.0, y2)) < 1e-8
    ests = []
    for seed in range(8):
        pf = O.OraclePF.init(O.MODEL_LGSSM, F.lgssm_params(), 2000, [ys[0]], seed=seed)
        for t in range(1, len(ys)):
            pf.maybe_resample(kind=O.RESAMPLE_SYSTEMATIC)
            pf.step([ys[t]])
        ests.append(pf.log_ml_estimate())
    se = np.std(ests) / math.sqrt(len(ests))
    assert abs(np.mean(ests) - kf) < 4 * se + 0.05, (np.mean(ests), kf, se)


def test_blr_evidence_oracle(O):
    # cfg 4 exact evidence: log N(y; 0, sigma^2 I + X X') against scipy
    from scipy import stats
    X, y, sigma = F.blr_problem(D=6, n=5)
    ref = stats.multivariate_normal.logpdf(y, np.zeros(5), sigma ** 2 * np.eye(5) + X @ X.T)
    assert abs(O.blr_log_evidence(X, y, sigma) - ref) < 1e-10


def test_resamplers_against_python_bigint_definition(O):
    # SURVEY.md Appendix B, restated independently with Python integers
    rng = np.random.default_rng(11)
    for n in (1, 2, 7, 100, 1000):
        lw = 3.0 * rng.standard_normal(n)
        q, Q = O.quantized_weights(lw)
        q = [int(v) for v in q]
        assert sum(q) == Q
        cdf = np.cumsum(np.array(q, dtype=object))
        u0 = int(rng.integers(0, 2 ** 32))
        anc = O.resample(O.RESAMPLE_SYSTEMATIC, lw, u0=u0)
        exp = []
        for k in range(n):
            tau = ((k << 32) + u0) * Q // (n << 32)
            exp.append(next(j for j in range(n) if cdf[j] > tau))
        assert list(anc) == exp
        us = rng.integers(0, 2 ** 64, size=n, dtype=np.uint64)
        anc = O.resample(O.RESAMPLE_MULTINOMIAL, lw, u64s=us)
        exp = [next(j for j in range(n) if cdf[j] > (int(u) * Q >> 64)) for u in us]
        assert list(anc) == exp
        anc = O.resample(O.RESAMPLE_RESIDUAL, lw, u0=u0)
        cnt = [n * qj // Q for qj in q]
        res = [n * qj - c * Q for qj, c in zip(q, cnt)]
        R = n - sum(cnt)
        rc = np.cumsum(np.array(res, dtype=object))
        for k in range(R):
            tau = ((k << 32) + u0) * Q >> 32
            cnt[next(j for j in range(n) if rc[j] > tau)] += 1
        assert list(anc) == [j for j in range(n) for _ in range(cnt[j])]
        # sorted multinomial: thresholds T_k Q / T from integer spacings, exact rational comparison
        e = [int(v) for v in rng.integers(1, 2 ** 36, size=n + 1, dtype=np.uint64)]
        anc = O.resample(O.RESAMPLE_MULTINOMIAL_SORTED, lw, u64s=np.array(e, dtype=np.uint64))
        T, run, exp = sum(e), 0, []
        for k in range(n):
            run += e[k]
            exp.append(next(j for j in range(n) if cdf[j] * T > run * Q))
        assert list(anc) == exp and exp == sorted(exp)


def test_sorted_multinomial_is_multinomial_in_distribution(O):
    # exponential spacings give the order statistics of iid uniforms: with exponential integer spacings the offspring counts
    # are multinomial(N, w) -- mean N w_j and variance N w_j (1 - w_j) over repetitions (systematic would have variance < 1/4)
    rng = np.random.default_rng(8)
    n, reps = 200, 400
    lw = 1.5 * rng.standard_normal(n)
    w = np.exp(lw - O.logsumexp(lw))
    counts = np.zeros((reps, n))
    for r in range(reps):
        e = np.floor(rng.exponential(size=n + 1) * 2.0 ** 30).astype(np.uint64) + 1
        counts[r] = np.bincount(O.resample(O.RESAMPLE_MULTINOMIAL_SORTED, lw, u64s=e), minlength=n)
    big = w > 0.01
    assert np.allclose(counts.mean(axis=0)[big], n * w[big], rtol=0.08)
    assert np.allclose(counts.var(axis=0)[big], (n * w * (1 - w))[big], rtol=0.3)


def test_resampler_statistics(O):
    # offspring counts are unbiased: E[count_j] = N w_j (systematic: |count - N w| < 1)
    rng = np.random.default_rng(4)
    n = 2000
    lw = 2.0 * rng.standard_normal(n)
    w = np.exp(lw - O.logsumexp(lw))
    anc = O.resample(O.RESAMPLE_SYSTEMATIC, lw, u0=123456789)
    cnt = np.bincount(anc, minlength=n)
    assert np.all(np.abs(cnt - n * w) < 1.0 + 1e-6)
    assert np.all(np.diff(anc) >= 0)


def test_guide_table_bracket_contains_every_multinomial_draw():
    """The argument behind the guided multinomial kernel (csrc/kernels_resample.cuh), checked on the CPU with exact integers:
    with guide[j] = smallest i with cdf[i] > floor(j Q / N), a draw tau in [0, Q) whose cell estimate j is anywhere in
    {J-2, J-1, J} (J = floor(tau N / Q); the device gets j from a double product minus one) has its ancestor -- the
    smallest i with cdf[i] > tau -- inside [guide[j], guide[min(j + 3, N - 1)] or N - 1]."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=200, deadline=None)
    @given(st.integers(1, 300), st.integers(0, 2 ** 31), st.sampled_from(["uniform", "sparse", "tiny", "onehot"]))
    def check(n, seed, shape):
        rng = np.random.default_rng(seed)
        if shape == "uniform":
            q = rng.integers(0, 2 ** 40, size=n)
        elif shape == "sparse":
            q = rng.integers(0, 2 ** 40, size=n) * (rng.random(n) < 0.1)
        elif shape == "tiny":
            q = rng.integers(0, 4, size=n)
            q[rng.integers(0, n)] += 2 ** 45
        else:
            q = np.zeros(n, dtype=np.int64)
            q[rng.integers(0, n)] = 2 ** 50
        q = [int(v) for v in q]
        Q = sum(q)
        if Q == 0:
            return
        cdf, c = [], 0
        for v in q:
            c += v
            cdf.append(c)
        cdf_np = np.array(cdf, dtype=object)

        def first_above(t):                       # smallest i with cdf[i] > t
            lo, hi = 0, n - 1
            while lo < hi:
                mid = (lo + hi) // 2
                if cdf[mid] > t:
                    hi = mid
                else:
                    lo = mid + 1
            return lo
        guide = [first_above((j * Q) // n) for j in range(n)]
        taus = [int(t) for t in rng.integers(0, min(Q, 2 ** 62), size=40)] + [0, Q - 1]
        for tau in taus:
            ans = first_above(tau)
            J = (tau * n) // Q
            for j in (J - 2, J - 1, J):
                j = min(max(j, 0), n - 1)
                a = guide[j]
                b = guide[j + 3] if j + 3 < n else n - 1
                assert a <= ans <= b, (n, tau, j, a, ans, b)
    check()
