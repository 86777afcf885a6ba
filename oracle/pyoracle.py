"""ctypes binding of the CPU oracle (oracle/gen_oracle.c).  TEST INFRASTRUCTURE ONLY.

Importable from tests/, __graft_entry__.smoke() and bench.py's CPU legs; never from the
product package under gen.jl_b200/.
"""
import ctypes as C
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
if _HERE not in sys.path:
    sys.path.insert(0, _HERE)
import build as _build  # noqa: E402

MODEL_LGSSM, MODEL_SV, MODEL_BEARINGS, MODEL_HMM, MODEL_BLR = 1, 2, 3, 4, 5
PROPOSAL_DEFAULT, PROPOSAL_AFFINE_NORMAL, PROPOSAL_CATEGORICAL_TABLE = 0, 1, 2
RESAMPLE_MULTINOMIAL, RESAMPLE_SYSTEMATIC, RESAMPLE_RESIDUAL, RESAMPLE_MULTINOMIAL_SORTED = 0, 1, 2, 3

_lib = None
_dp = C.POINTER(C.c_double)
_i64p = C.POINTER(C.c_int64)
_u64p = C.POINTER(C.c_uint64)


def lib():
    global _lib
    if _lib is None:
        path = _build.build()
        L = C.CDLL(path)
        d, i, i64, u64, u32, vp = C.c_double, C.c_int, C.c_int64, C.c_uint64, C.c_uint32, C.c_void_p
        sig = {
            "go_logpdf_normal": (d, [d, d, d]),
            "go_logpdf_gamma": (d, [d, d, d]),
            "go_logpdf_bernoulli": (d, [i, d]),
            "go_logpdf_categorical": (d, [i64, _dp, i64]),
            "go_logpdf_mvnormal": (d, [_dp, _dp, _dp, i]),
            "go_random_normal": (d, [d, d, d]),
            "go_random_bernoulli": (i, [d, d]),
            "go_random_categorical": (i64, [_dp, i64, d]),
            "go_random_gamma_philox": (d, [d, d, u64, u64, u32]),
            "go_logsumexp": (d, [_dp, i64]),
            "go_logsumexp2": (d, [d, d]),
            "go_effective_sample_size": (d, [_dp, i64]),
            "go_normalize_weights": (d, [_dp, i64, _dp]),
            "go_philox4x32_10": (None, [C.POINTER(u32), C.POINTER(u32), C.POINTER(u32)]),
            "go_philox_normals": (None, [u64, u32, i64, i64, i, _dp]),
            "go_philox_uniforms": (None, [u64, u32, i64, i64, i, _dp]),
            "go_philox_resample_u0": (u32, [u64, u32]),
            "go_philox_multinomial_u64": (None, [u64, u32, i64, i64, _u64p]),
            "go_dexp": (d, [d]),
            "go_quant_bits": (i, [i64]),
            "go_quantize": (u64, [d, i]),
            "go_quantized_weights": (u64, [_dp, i64, d, i, _u64p]),
            "go_resample_systematic": (i, [_dp, i64, u32, _i64p]),
            "go_resample_multinomial": (i, [_dp, i64, _u64p, _i64p]),
            "go_resample_multinomial_sorted": (i, [_dp, i64, _u64p, _i64p]),
            "go_resample_residual": (i, [_dp, i64, u32, _i64p]),
            "go_kalman_logml": (d, [d, d, d, d, d, _dp, i]),
            "go_hmm_forward": (d, [_dp, _dp, _dp, i, i, _i64p, i]),
            "go_blr_log_evidence": (d, [_dp, _dp, i, i, d]),
            "go_model_state_dim": (i, [i, _dp]),
            "go_model_num_normals": (i, [i, _dp, i, i]),
            "go_model_num_uniforms": (i, [i, _dp, i, i]),
            "go_model_num_obs": (i, [i, _dp]),
            "go_set_threads": (None, [i]),
            "go_set_storage_f32": (None, [i]),
            "go_get_max_threads": (i, []),
            "go_pf_init": (vp, [i, _dp, i, i64, u64, _dp, i, _dp, _dp, _dp]),
            "go_pf_step": (i, [vp, _dp, i, _dp, _dp, _dp, _dp]),
            "go_pf_maybe_resample": (i, [vp, d, i, i, u32, _u64p, _dp]),
            "go_pf_resample_given": (i, [vp, _i64p]),
            "go_pf_rejuvenate": (i64, [vp, i, _dp, d, u32, _dp, _dp, _dp, C.POINTER(C.c_uint8)]),
            "go_pf_enable_window": (None, [vp, i]),
            "go_pf_rejuvenate_window": (i64, [vp, i, i, _dp, d, u32, _dp, _dp, _dp, C.POINTER(C.c_uint8)]),
            "go_pf_log_ml_estimate": (d, [vp]),
            "go_pf_log_weights": (_dp, [vp]),
            "go_pf_parents": (_i64p, [vp]),
            "go_pf_state": (_dp, [vp, i]),
            "go_pf_log_ml_est_field": (d, [vp]),
            "go_pf_num_particles": (i64, [vp]),
            "go_pf_step_index": (i, [vp]),
            "go_pf_copy": (vp, [vp]),
            "go_pf_set_log_weights": (None, [vp, _dp]),
            "go_pf_free": (None, [vp]),
            "go_importance_sampling": (d, [i, _dp, i, i64, u64, _dp, i, _dp, _dp, _dp, _dp, C.POINTER(vp)]),
        }
        for name, (res, args) in sig.items():
            f = getattr(L, name)
            f.restype = res
            f.argtypes = args
        _lib = L
    return _lib


def _d(a):
    """float64 C-contiguous array -> (keepalive, pointer); None -> NULL."""
    if a is None:
        return None, None
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(_dp)


def logsumexp(a):
    a, p = _d(a)
    return lib().go_logsumexp(p, a.size)


def effective_sample_size(lognormw):
    a, p = _d(lognormw)
    return lib().go_effective_sample_size(p, a.size)


def normalize_weights(logw):
    a, p = _d(logw)
    out = np.empty_like(a)
    lse = lib().go_normalize_weights(p, a.size, out.ctypes.data_as(_dp))
    return lse, out


def logpdf_categorical(x, probs):
    a, p = _d(probs)
    return lib().go_logpdf_categorical(int(x), p, a.size)


def logpdf_mvnormal(x, mu, cov):
    xa, xp = _d(x)
    ma, mp = _d(mu)
    ca, cp = _d(cov)
    return lib().go_logpdf_mvnormal(xp, mp, cp, xa.size)


def philox(ctr, key):
    u32 = C.c_uint32
    c = (u32 * 4)(*ctr)
    k = (u32 * 2)(*key)
    o = (u32 * 4)()
    lib().go_philox4x32_10(c, k, o)
    return list(o)


def philox_normals(seed, step, pid0, n, ndraw):
    out = np.empty((ndraw, n), dtype=np.float64)
    lib().go_philox_normals(seed, step, pid0, n, ndraw, out.ctypes.data_as(_dp))
    return out


def philox_uniforms(seed, step, pid0, n, ndraw):
    out = np.empty((ndraw, n), dtype=np.float64)
    lib().go_philox_uniforms(seed, step, pid0, n, ndraw, out.ctypes.data_as(_dp))
    return out


def philox_multinomial_u64(seed, step, k0, n):
    out = np.empty(n, dtype=np.uint64)
    lib().go_philox_multinomial_u64(seed, step, k0, n, out.ctypes.data_as(_u64p))
    return out


def quantized_weights(logw, b=None, m=None):
    a, p = _d(logw)
    if m is None:
        m = float(np.max(a))
    if b is None:
        b = lib().go_quant_bits(a.size)
    q = np.empty(a.size, dtype=np.uint64)
    tot = lib().go_quantized_weights(p, a.size, m, b, q.ctypes.data_as(_u64p))
    return q, tot


def resample(kind, logw, u0=0, u64s=None):
    """0-based ancestors (int64) by the integer-CDF spec; raises if all weights are -inf."""
    a, p = _d(logw)
    anc = np.empty(a.size, dtype=np.int64)
    ap = anc.ctypes.data_as(_i64p)
    if kind == RESAMPLE_SYSTEMATIC:
        rc = lib().go_resample_systematic(p, a.size, u0, ap)
    elif kind == RESAMPLE_RESIDUAL:
        rc = lib().go_resample_residual(p, a.size, u0, ap)
    elif kind == RESAMPLE_MULTINOMIAL_SORTED:
        e = np.ascontiguousarray(u64s, dtype=np.uint64)            # the n + 1 integer spacings
        assert e.size == a.size + 1, "sorted multinomial takes n + 1 spacings"
        rc = lib().go_resample_multinomial_sorted(p, a.size, e.ctypes.data_as(_u64p), ap)
    else:
        u = np.ascontiguousarray(u64s, dtype=np.uint64)
        rc = lib().go_resample_multinomial(p, a.size, u.ctypes.data_as(_u64p), ap)
    if rc != 0:
        raise ValueError("oracle resample failed rc=%d" % rc)
    return anc


def kalman_logml(m0, s0, a, q, r, y):
    ya, yp = _d(y)
    return lib().go_kalman_logml(m0, s0, a, q, r, yp, ya.size)


def hmm_forward(prior, emission_MxK, transition_KxK, obs):
    """emission_MxK[x-1, z-1] = p(x|z); transition_KxK[z-1, prev-1] = p(z|prev) (Gen test layout)."""
    pr, prp = _d(prior)
    em, emp = _d(np.asfortranarray(emission_MxK).ravel(order="F"))
    tr, trp = _d(np.asfortranarray(transition_KxK).ravel(order="F"))
    o = np.ascontiguousarray(obs, dtype=np.int64)
    M, K = np.shape(emission_MxK)
    return lib().go_hmm_forward(prp, emp, trp, K, M, o.ctypes.data_as(_i64p), o.size)


def blr_log_evidence(X, y, sigma):
    Xa, Xp = _d(X)
    ya, yp = _d(y)
    n, D = Xa.shape
    return lib().go_blr_log_evidence(Xp, yp, n, D, sigma)


class OraclePF:
    """Python handle on go_pf_t (mirrors Gen's ParticleFilterState fields)."""

    def __init__(self, ptr, kind, params):
        self._p = ptr
        self.kind = kind
        self.params = np.ascontiguousarray(params, dtype=np.float64)
        self.S = lib().go_model_state_dim(kind, self.params.ctypes.data_as(_dp))

    @classmethod
    def init(cls, kind, params, n, obs, seed=0, proposal_kind=0, pargs=None, normals=None, uniforms=None):
        pa, pp = _d(params)
        oa, op = _d(obs)
        ga, gp = _d(pargs)
        na, np_ = _d(normals)
        ua, up = _d(uniforms)
        ptr = lib().go_pf_init(kind, pp, pa.size, n, seed, op, proposal_kind, gp, np_, up)
        if not ptr:
            raise ValueError("go_pf_init failed")
        return cls(ptr, kind, pa)

    def step(self, obs, proposal_kind=0, pargs=None, normals=None, uniforms=None):
        oa, op = _d(obs)
        ga, gp = _d(pargs)
        na, np_ = _d(normals)
        ua, up = _d(uniforms)
        incr = np.empty(self.n, dtype=np.float64)
        rc = lib().go_pf_step(self._p, op, proposal_kind, gp, np_, up, incr.ctypes.data_as(_dp))
        assert rc == 0
        return incr

    def resample_given(self, parents):
        """The resampling branch of maybe_resample! with the given 1-based parents (shared genealogy, own weights)."""
        p = np.ascontiguousarray(parents, dtype=np.int64)
        rc = lib().go_pf_resample_given(self._p, p.ctypes.data_as(_i64p))
        assert rc == 1, rc

    def maybe_resample(self, ess_threshold=None, kind=RESAMPLE_SYSTEMATIC, u0=None, u64s=None):
        if ess_threshold is None:
            ess_threshold = self.n / 2
        ess = C.c_double()
        up = None
        if u64s is not None:
            u64s = np.ascontiguousarray(u64s, dtype=np.uint64)
            up = u64s.ctypes.data_as(_u64p)
        rc = lib().go_pf_maybe_resample(self._p, float(ess_threshold), kind, 0 if u0 is None else 1,
                                        0 if u0 is None else int(u0), up, C.byref(ess))
        if rc < 0:
            raise ValueError("oracle resample failed rc=%d" % rc)
        self.last_ess = ess.value
        return bool(rc)

    def rejuvenate(self, obs, move=0, drift_std=0.0, move_index=0, normals=None, uniforms=None):
        """mh() on every particle's latest latent choices (src/inference/mh.jl:16-27 / :37-58).
        Returns (#accepted, log acceptance ratios, accepted flags)."""
        oa, op = _d(obs)
        na, np_ = _d(normals)
        ua, up = _d(uniforms)
        alpha = np.empty(self.n, dtype=np.float64)
        acc = np.empty(self.n, dtype=np.uint8)
        r = lib().go_pf_rejuvenate(self._p, move, op, float(drift_std), move_index, np_, up, alpha.ctypes.data_as(_dp),
                                   acc.ctypes.data_as(C.POINTER(C.c_uint8)))
        if r < 0:
            raise ValueError("oracle: MH move unsupported in this state / for this model")
        return r, alpha, acc.astype(bool)

    def enable_window(self, W):
        lib().go_pf_enable_window(self._p, int(W))

    def rejuvenate_window(self, a, b, obs_window, drift_std, move_index=0, normals=None, uniforms=None):
        """One block drift move on the latent choices of steps a..b per particle (docs/src/tutorials/smc.md:492-541).
        Returns (#accepted, log acceptance ratios, accepted flags)."""
        oa, op = _d(obs_window)
        na, np_ = _d(normals)
        ua, up = _d(uniforms)
        alpha = np.empty(self.n, dtype=np.float64)
        acc = np.empty(self.n, dtype=np.uint8)
        r = lib().go_pf_rejuvenate_window(self._p, int(a), int(b), op, float(drift_std), int(move_index), np_, up,
                                          alpha.ctypes.data_as(_dp), acc.ctypes.data_as(C.POINTER(C.c_uint8)))
        if r < 0:
            raise ValueError("oracle: window move unsupported in this state / for this model")
        return r, alpha, acc.astype(bool)

    @property
    def n(self):
        return lib().go_pf_num_particles(self._p)

    @property
    def log_weights(self):
        return np.ctypeslib.as_array(lib().go_pf_log_weights(self._p), shape=(self.n,)).copy()

    @log_weights.setter
    def log_weights(self, lw):
        a, p = _d(lw)
        assert a.size == self.n
        lib().go_pf_set_log_weights(self._p, p)

    @property
    def parents(self):
        return np.ctypeslib.as_array(lib().go_pf_parents(self._p), shape=(self.n,)).copy()

    @property
    def log_ml_est(self):
        return lib().go_pf_log_ml_est_field(self._p)

    def state(self, field):
        return np.ctypeslib.as_array(lib().go_pf_state(self._p, field), shape=(self.n,)).copy()

    def log_ml_estimate(self):
        return lib().go_pf_log_ml_estimate(self._p)

    def copy(self):
        return OraclePF(lib().go_pf_copy(self._p), self.kind, self.params)

    def __del__(self):
        if getattr(self, "_p", None):
            lib().go_pf_free(self._p)
            self._p = None


def importance_sampling(kind, params, n, obs, seed=0, proposal_kind=0, pargs=None, normals=None, uniforms=None):
    pa, pp = _d(params)
    oa, op = _d(obs)
    ga, gp = _d(pargs)
    na, np_ = _d(normals)
    ua, up = _d(uniforms)
    out = np.empty(n, dtype=np.float64)
    ptr = C.c_void_p()
    lml = lib().go_importance_sampling(kind, pp, pa.size, n, seed, op, proposal_kind, gp, np_, up,
                                       out.ctypes.data_as(_dp), C.byref(ptr))
    return OraclePF(ptr.value, kind, pa), out, lml


# ---- the remaining built-in distributions: numpy restatements of the reference formulas ----
def logpdf_broadcasted_normal(x, mu, std):
    """normal.jl:61-68: sum over the broadcast shape of the elementwise normal log density."""
    x, mu, std = np.broadcast_arrays(np.asarray(x, float), np.asarray(mu, float), np.asarray(std, float))
    z = (x - mu) / std
    return float(np.sum(-(z * z + np.log(2 * np.pi)) / 2 - np.log(std)))


def logpdf_dirichlet(x, alpha):
    """dirichlet.jl:12-21 (isapprox(sum(x), 1) with Julia's default rtol = sqrt(eps))."""
    from scipy.special import gammaln
    x, alpha = np.asarray(x, float), np.asarray(alpha, float)
    s = x.sum()
    if len(x) == len(alpha) and abs(s - 1) <= np.sqrt(np.finfo(float).eps) * max(abs(s), 1) and np.all(x >= 0) and np.all(alpha >= 0):
        with np.errstate(divide="ignore", invalid="ignore"):
            ll = np.sum((alpha - 1) * np.log(x))
        return float(ll - (np.sum(gammaln(alpha)) - gammaln(alpha.sum())))
    return -np.inf


def logpdf_piecewise_uniform(x, bounds, probs):
    """piecewise_uniform.jl:19-46."""
    if x <= bounds[0] or x >= bounds[-1]:
        return -np.inf
    b = 0
    while x > bounds[b + 1]:
        b += 1
    return float(np.log(probs[b]) - np.log(bounds[b + 1] - bounds[b]))


def logpdf_beta_uniform(x, theta, a, b):
    """beta_uniform.jl:10-18 with logsumexp(x1, x2) of inference.jl:8-11."""
    from scipy import stats
    if x < 0 or x > 1:
        return -np.inf
    return logsumexp2(np.log(theta) + stats.beta.logpdf(x, a, b), np.log(1.0 - theta))


def logsumexp2(x1, x2):
    return lib().go_logsumexp2(float(x1), float(x2))


def _sample_unimodal(u, mode, p_mode, kmin, kmax, up, down):
    """This repo's sampler spec for poisson / binom / neg_binom (the reference delegates to Distributions.jl on Julia's
    global RNG: parity unpinned): exact inversion of ONE uniform against the masses in the order mode, mode+1, mode-1, ..."""
    if u < p_mode:
        return mode
    u -= p_mode
    pu = pd = p_mode
    ku = kd = mode
    while True:
        moved = False
        if ku < kmax:
            pu *= up(ku); ku += 1; moved = True
            if u < pu:
                return ku
            u -= pu
        if kd > kmin:
            pd *= down(kd); kd -= 1; moved = True
            if u < pd:
                return kd
            u -= pd
        if not moved:
            return ku


def random_poisson(lam, u):
    import math
    mode = int(math.floor(lam))
    pm = math.exp(mode * math.log(lam) - lam - math.lgamma(mode + 1.0))
    return _sample_unimodal(u, mode, pm, 0, 1 << 62, lambda k: lam / (k + 1), lambda k: k / lam)


def random_binom(n, p, u):
    import math
    from scipy import stats
    mode = min(int(math.floor((n + 1.0) * p)), int(n))
    pm = math.exp(math.lgamma(n + 1.0) - math.lgamma(mode + 1.0) - math.lgamma(n - mode + 1.0) + mode * math.log(p) + (n - mode) * math.log1p(-p))
    odds = p / (1.0 - p)
    return _sample_unimodal(u, mode, pm, 0, int(n), lambda k: (n - k) / (k + 1.0) * odds, lambda k: k / (n - k + 1.0) / odds)


def random_neg_binom(r, p, u):
    import math
    q = 1.0 - p
    mode = int(math.floor((r - 1.0) * q / p)) if r > 1.0 else 0
    pm = math.exp(math.lgamma(mode + r) - math.lgamma(mode + 1.0) - math.lgamma(r) + r * math.log(p) + mode * math.log1p(-p))
    return _sample_unimodal(u, mode, pm, 0, 1 << 62, lambda k: (k + r) / (k + 1.0) * q, lambda k: k / (k - 1.0 + r) / q)
