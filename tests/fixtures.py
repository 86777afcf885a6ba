"""Shared synthetic workloads for tests and bench.py.  Data generation uses
numpy's Philox bit generator with the fixed seeds of BASELINE.md so every run sees the same series."""
import math

import numpy as np

# ---- the reference's own PF test fixture (test/inference/particle_filter.jl:52-81) ----
HMM_PRIOR = np.array([0.2, 0.3, 0.5])
HMM_EMISSION = np.array([[0.1, 0.2, 0.7], [0.2, 0.7, 0.1], [0.7, 0.2, 0.1]]).T      # [x, z] = p(x | z)
HMM_TRANSITION = np.array([[0.4, 0.4, 0.2], [0.2, 0.3, 0.5], [0.9, 0.05, 0.05]]).T  # [z, prev_z]
HMM_OBS = [1, 1, 2, 3]
HMM_EXPECTED_LOG_ML = -4.87645083351704   # log(hmm_forward_alg(...)) of that fixture


def hmm_params():
    K, M = 3, 3
    return np.concatenate([[K, M], HMM_PRIOR, HMM_TRANSITION.ravel(order="F"), HMM_EMISSION.ravel(order="F")])


def hmm_init_proposal(x):
    """init_proposal of the reference test (:104-107): prior .* emission[x, :] normalised."""
    d = HMM_PRIOR * HMM_EMISSION[x - 1, :]
    return d / d.sum()


def hmm_step_proposal(x):
    """step_proposal (:117-127): column prev_z = transition[:, prev_z] .* emission[x, :] normalised."""
    tab = HMM_TRANSITION * HMM_EMISSION[x - 1, :][:, None]
    return tab / tab.sum(axis=0, keepdims=True)


# ---- cfg 1: 1-D linear-Gaussian SSM ----
LGSSM = dict(m0=0.0, s0=1.0, a=0.9, q=0.5, r=1.0)


def lgssm_params():
    return [LGSSM["m0"], LGSSM["s0"], LGSSM["a"], LGSSM["q"], LGSSM["r"]]


def lgssm_series(T1=101, seed=20240):
    rng = np.random.Generator(np.random.Philox(seed))
    x = LGSSM["m0"] + LGSSM["s0"] * rng.standard_normal()
    ys = []
    for t in range(T1):
        if t > 0:
            x = LGSSM["a"] * x + LGSSM["q"] * rng.standard_normal()
        ys.append(x + LGSSM["r"] * rng.standard_normal())
    return np.array(ys)


# ---- cfg 2: stochastic volatility ----
SV = dict(mu=-1.0, phi=0.95, sigma=0.25)


def sv_params():
    return [SV["mu"], SV["phi"], SV["sigma"], SV["sigma"] / math.sqrt(1.0 - SV["phi"] ** 2)]


def sv_series(T1=1001, seed=20241):
    rng = np.random.Generator(np.random.Philox(seed))
    p = sv_params()
    h = p[0] + p[3] * rng.standard_normal()
    ys = []
    for t in range(T1):
        if t > 0:
            h = p[0] + p[1] * (h - p[0]) + p[2] * rng.standard_normal()
        ys.append(math.exp(h / 2.0) * rng.standard_normal())
    return np.array(ys)


# ---- cfg 3: bearings-only tracking (docs/src/tutorials/smc.md:607-661) ----
BEARINGS_PARAMS = [0.01, 0.01, 0.95, 0.01, 0.002, 0.01, -0.013, 0.01, 1e-6, 0.005]


def bearings_series(T1=51, seed=20242):
    rng = np.random.Generator(np.random.Philox(seed))
    p = BEARINGS_PARAMS
    x = p[0] + p[1] * rng.standard_normal()
    y = p[2] + p[3] * rng.standard_normal()
    vx = p[4] + p[5] * rng.standard_normal()
    vy = p[6] + p[7] * rng.standard_normal()
    zs = [math.atan2(y, x) + p[9] * rng.standard_normal()]
    for t in range(1, T1):
        vx = vx + math.sqrt(p[8]) * rng.standard_normal()
        vy = vy + math.sqrt(p[8]) * rng.standard_normal()
        x, y = x + vx, y + vy
        zs.append(math.atan2(y, x) + p[9] * rng.standard_normal())
    return np.array(zs)


# ---- cfg 4: Bayesian linear regression ----
def blr_problem(D=64, n=64, sigma=0.5, seed=20243):
    rng = np.random.Generator(np.random.Philox(seed))
    X = rng.standard_normal((n, D))
    w = rng.standard_normal(D)
    y = X @ w + sigma * rng.standard_normal(n)
    return X, y, sigma


def blr_params(X, sigma):
    n, D = X.shape
    return np.concatenate([[D, n, sigma], X.ravel()])


# ---- cfg 5: resampling microbenchmark weight shapes ----
def resample_logw(n, shape="normal", seed=20244):
    if shape == "normal":
        rng = np.random.Generator(np.random.Philox(seed))
        return 3.0 * rng.standard_normal(n)
    if shape == "equal":
        return np.zeros(n)
    if shape == "onehot":
        lw = np.full(n, -50.0)
        lw[0] = 0.0
        return lw
    raise ValueError(shape)
