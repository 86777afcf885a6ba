"""Host-side mirror of Gen.jl's particle-filter / importance-sampling interface over libgenb200.

Julia is not available in this image, so this module plays the role of the `GenB200.jl` glue
(INTEGRATION.md): the same function names, argument order, return values and error behaviour as
src/inference/particle_filter.jl and src/inference/importance.jl of the reference, with Julia's
trailing `!` spelled as a trailing underscore.  Every call goes through the C ABI
(include/genb200.h); nothing here computes on the host.

    state = initialize_particle_filter(model, (0,), obs0, N)
    for t in 1..T:
        maybe_resample_(state, ess_threshold=N/2)
        particle_filter_step_(state, (t,), (UnknownChange,), obs_t)
    log_ml_estimate(state)
"""
import ctypes as C
import math

import numpy as np

from . import _lib as L

UnknownChange = "UnknownChange"
NoChange = "NoChange"

_RESAMPLERS = {"multinomial": L.RESAMPLE_MULTINOMIAL, "systematic": L.RESAMPLE_SYSTEMATIC,
               "residual": L.RESAMPLE_RESIDUAL, "multinomial_sorted": L.RESAMPLE_MULTINOMIAL_SORTED}
_DTYPES = {"f64": L.F64, "float64": L.F64, "fp64": L.F64, np.float64: L.F64,
           "f32": L.F32, "float32": L.F32, "fp32": L.F32, np.float32: L.F32}


def _darr(a):
    if a is None:
        return None, None, 0
    a = np.ascontiguousarray(a, dtype=np.float64).reshape(-1)
    return a, a.ctypes.data_as(L.dp), a.size


def _flatten_obs(obs):
    """Observations cross the boundary as a flat Float64 vector in the model's fixed slot order
    (SURVEY.md section 8b).  Accepts a scalar, a sequence, or a choicemap-like dict (values in
    insertion order, e.g. {("chain", t, "z"): z_t})."""
    if isinstance(obs, dict):
        obs = list(obs.values())
    return np.atleast_1d(np.asarray(obs, dtype=np.float64)).reshape(-1)


class DeviceModel:
    """A generative function of the supported class (replaces a GenerativeFunction object)."""

    def __init__(self, kind, params, name):
        self.kind = kind
        self.name = name
        self.params = np.ascontiguousarray(params, dtype=np.float64).reshape(-1)
        h = C.c_void_p()
        L.check(L.lib().gb_model_create(kind, self.params.ctypes.data_as(L.dp), self.params.size, C.byref(h)))
        self._h = h

    @property
    def state_dim(self):
        return L.lib().gb_model_state_dim(self._h)

    @property
    def num_obs(self):
        return L.lib().gb_model_num_obs(self._h)

    def num_normals(self, is_init, proposal_kind=0):
        return L.lib().gb_model_num_normals(self._h, proposal_kind, int(is_init))

    def num_uniforms(self, is_init, proposal_kind=0):
        return L.lib().gb_model_num_uniforms(self._h, proposal_kind, int(is_init))

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                L.lib().gb_model_free(h)
            except Exception:
                pass
            self._h = None


def lgssm_model(m0=0.0, s0=1.0, a=0.9, q=0.5, r=1.0):
    """x0 ~ normal(m0, s0); x_t ~ normal(a x_{t-1}, q); y_t ~ normal(x_t, r)  via Unfold (cfg 1)."""
    return DeviceModel(L.MODEL_LGSSM, [m0, s0, a, q, r], "lgssm")


def sv_model(mu=-1.0, phi=0.95, sigma=0.25):
    """h0 ~ normal(mu, sigma/sqrt(1-phi^2)); h_t ~ normal(mu + phi (h-mu), sigma); y ~ normal(0, exp(h/2))."""
    return DeviceModel(L.MODEL_SV, [mu, phi, sigma, sigma / math.sqrt(1.0 - phi * phi)], "sv")


def bearings_model(x0=(0.01, 0.01), y0=(0.95, 0.01), vx0=(0.002, 0.01), vy0=(-0.013, 0.01),
                   velocity_var=1e-6, measurement_noise=0.005):
    """The tutorial's bearings-only tracker (docs/src/tutorials/smc.md:607-661)."""
    return DeviceModel(L.MODEL_BEARINGS, [*x0, *y0, *vx0, *vy0, velocity_var, measurement_noise], "bearings")


def hmm_model(prior, transition, emission):
    """Discrete HMM in the layout of test/inference/particle_filter.jl:52-78:
    transition[z, prev_z] = p(z | prev_z), emission[x, z] = p(x | z) (columns are distributions)."""
    prior = np.asarray(prior, dtype=np.float64)
    tr = np.asarray(transition, dtype=np.float64)
    em = np.asarray(emission, dtype=np.float64)
    K, M = prior.size, em.shape[0]
    assert tr.shape == (K, K) and em.shape == (M, K)
    p = np.concatenate([[K, M], prior, tr.ravel(order="F"), em.ravel(order="F")])
    return DeviceModel(L.MODEL_HMM, p, "hmm")


def blr_model(X, sigma):
    """w ~ Map(normal)(0, 1) (D weights); y_j ~ normal(x_j . w, sigma)  (cfg 4)."""
    X = np.ascontiguousarray(X, dtype=np.float64)
    n, D = X.shape
    return DeviceModel(L.MODEL_BLR, np.concatenate([[D, n, sigma], X.ravel()]), "blr")


class SpecProgram:
    """Builder of one SSA program of a GB_MODEL_SPEC model (include/genb200.h).  Methods return register numbers."""
    _CODES = dict(add=10, sub=11, mul=12, div=13, atan2=18)
    _UNARY = dict(neg=14, sqrt=15, exp=16, log=17, abs=19)

    def __init__(self):
        self.ops = []
        self.nreg = 0

    def _new(self):
        r = self.nreg
        self.nreg += 1
        if r >= 32:
            raise ValueError("a spec program may use at most 32 registers")
        return r

    def _emit(self, code, dst, a=0, b=0, imm=0.0):
        self.ops.append((code, dst, a, b, float(imm)))
        return dst

    def const(self, v):
        return self._emit(0, self._new(), imm=v)

    def prev(self, field):
        return self._emit(1, self._new(), field)

    def param(self, idx):
        return self._emit(2, self._new(), idx)

    def obs(self, idx=0):
        return self._emit(3, self._new(), idx)

    def time(self):
        return self._emit(4, self._new())

    def __getattr__(self, name):
        if name in SpecProgram._CODES:
            return lambda a, b: self._emit(SpecProgram._CODES[name], self._new(), a, b)
        if name in SpecProgram._UNARY:
            return lambda a: self._emit(SpecProgram._UNARY[name], self._new(), a)
        raise AttributeError(name)

    def sample_normal(self, mu, std):
        return self._emit(20, self._new(), mu, std)

    def observe_normal(self, x, mu, std):
        self._emit(21, x, mu, std)

    def sample_bernoulli(self, p):
        return self._emit(22, self._new(), p)

    def observe_bernoulli(self, x, p):
        self._emit(23, x, p)

    def sample_gamma(self, shape, scale):
        return self._emit(24, self._new(), shape, scale)

    def observe_gamma(self, x, shape, scale):
        self._emit(25, x, shape, scale)

    def sample_categorical(self, offset_reg, length):
        return self._emit(26, self._new(), offset_reg, int(length))

    def observe_categorical(self, x, offset_reg, length):
        self._emit(27, x, offset_reg, int(length))

    def store(self, field, reg):
        self._emit(30, field, reg)

    def add_weight(self, reg):
        self._emit(31, 0, reg)

    def logpdf_normal(self, x, mu, std):
        return self._emit(32, self._new(), mu, std, imm=x)

    # the other scalar built-ins (distributions/*.jl): dist is a name of _DISTS; one-parameter ones ignore b
    def sample(self, dist, a, b=None):
        """dst ~ dist(a, b) by inverse CDF (uniform, exponential, laplace, cauchy, geometric, uniform_discrete)."""
        return self._emit(40, self._new(), a, a if b is None else b, imm=_DISTS[dist])

    def observe(self, dist, x, a, b=None):
        """weight += logpdf(dist, reg x, a, b)."""
        self._emit(41, x, a, a if b is None else b, imm=_DISTS[dist])

    def logpdf(self, dist, x, a, b=None):
        return self._emit(42, self._new(), a, a if b is None else b, imm=_DISTS[dist] + 100 * x)

    # mvnormal(mu, cov) (mvnormal.jl:12-33): vectors are k CONSECUTIVE registers; cov = k*k model parameters (row-major)
    def vector(self, regs):
        """Copy `regs` into k fresh consecutive registers (x + 0) and return the first."""
        if getattr(self, "_zero", None) is None:
            self._zero = self.const(0.0)
        zero = self._zero
        first = None
        for r in regs:
            d = self.add(r, zero)
            first = d if first is None else first
        return first

    def sample_mvnormal(self, mu_first, cov_param_offset, k):
        """k consecutive registers ~ mvnormal(regs mu_first .. mu_first+k-1, params[offset : offset + k*k]); returns the first."""
        first = self.nreg
        for _ in range(int(k)):
            self._new()
        return self._emit(43, first, mu_first, int(cov_param_offset), imm=int(k))

    def observe_mvnormal(self, x_first, mu_first, cov_param_offset, k):
        self._emit(44, x_first, mu_first, int(cov_param_offset), imm=int(k))


def spec_model(state_dim, params, num_obs, generate_program, step_program=None, name="spec"):
    """A model of the supported class from its declarative specification: `generate_program` is the static-IR
    function at t = 0, `step_program` the Unfold kernel (extend-by-one update)."""
    params = np.asarray(params, dtype=np.float64).reshape(-1)
    gops = list(generate_program.ops)
    sops = list(step_program.ops) if step_program is not None else []
    flat = [float(state_dim), float(params.size), float(num_obs), float(len(gops)), float(len(sops))]
    flat += list(params)
    for op in gops + sops:
        flat += [float(v) for v in op]
    return DeviceModel(L.MODEL_SPEC, flat, name)


def spec_jit_source(model):
    """The CUDA C++ translation unit the library generates from a spec model's programs (NVRTC input)."""
    need = C.c_int64()
    L.check(L.lib().gb_spec_jit_source(model._h, None, 0, C.byref(need)))
    buf = C.create_string_buffer(need.value)
    L.check(L.lib().gb_spec_jit_source(model._h, buf, need.value, None))
    return buf.value.decode()


def spec_jit_compile(model, dtype="f64"):
    """Compile (not load) the specialised kernels for sm_100a; returns the cubin size.  Needs libnvrtc, no GPU."""
    nb = C.c_int64()
    L.check(L.lib().gb_spec_jit_compile(model._h, _DTYPES[dtype], C.byref(nb)))
    return nb.value


class AffineNormalProposal:
    """x ~ normal(c0 + c1 * x_prev, std): proposal_args = (c0, c1, std)."""
    kind = L.PROPOSAL_AFFINE_NORMAL


class CategoricalTableProposal:
    """z ~ categorical(table[:, prev_z]): proposal_args = (table,) with K probs at t=0, K x K after."""
    kind = L.PROPOSAL_CATEGORICAL_TABLE


def _proposal(proposal, proposal_args):
    if proposal is None:
        return L.PROPOSAL_DEFAULT, None
    kind = proposal.kind
    if kind == L.PROPOSAL_CATEGORICAL_TABLE:
        tab = np.asarray(proposal_args[0], dtype=np.float64)
        return kind, tab.ravel(order="F") if tab.ndim == 2 else tab
    return kind, np.asarray(proposal_args, dtype=np.float64)


class DeviceTraces:
    """Lazy view of the traces of a device particle store: columns are copied to the host on access."""

    def __init__(self, state):
        self._state = state

    def __len__(self):
        return self._state.num_particles

    def column(self, field):
        return self._state.get_state(field)

    def __getitem__(self, i):
        return tuple(self.column(f)[i] for f in range(self._state.state_dim))


class DeviceParticleFilterState:
    """ParticleFilterState (particle_filter.jl:35-41) held as SoA columns in HBM."""

    def __init__(self, handle, model, dtype):
        self._h = handle
        self.model = model
        self.dtype = dtype
        self.resampler = "multinomial"
        self._last_t = None   # the model's time argument (extend-by-one check)

    # -- fields of the reference struct, materialised lazily --
    @property
    def num_particles(self):
        nl, ng = C.c_int64(), C.c_int64()
        L.check(L.lib().gb_pf_num_particles(self._h, C.byref(nl), C.byref(ng)))
        return nl.value

    @property
    def state_dim(self):
        return self.model.state_dim

    @property
    def log_weights(self):
        return get_log_weights(self)

    @log_weights.setter
    def log_weights(self, lw):
        a, p, n = _darr(lw)
        if n != self.num_particles:
            raise ValueError("log_weights has the wrong length")
        L.check(L.lib().gb_pf_set_log_weights(self._h, p))

    @property
    def parents(self):
        out = np.empty(self.num_particles, dtype=np.int64)
        L.check(L.lib().gb_pf_get_parents(self._h, out.ctypes.data_as(L.i64p)))
        return out

    @property
    def log_ml_est(self):
        v = C.c_double()
        L.check(L.lib().gb_pf_get_log_ml_est(self._h, C.byref(v)))
        return v.value

    @property
    def traces(self):
        return DeviceTraces(self)

    @property
    def step_index(self):
        v = C.c_int()
        L.check(L.lib().gb_pf_step_index(self._h, C.byref(v)))
        return v.value

    def get_state(self, field):
        out = np.empty(self.num_particles, dtype=np.float64)
        L.check(L.lib().gb_pf_get_state(self._h, field, out.ctypes.data_as(L.dp)))
        return out

    def ess(self):
        e, l = C.c_double(), C.c_double()
        L.check(L.lib().gb_pf_get_ess(self._h, C.byref(e), C.byref(l)))
        return e.value

    # -- validation mode (pre-drawn noise; Gen has no RNG hook) --
    def set_noise(self, normals=None, uniforms=None):
        n = self.num_particles
        na = None if normals is None else np.ascontiguousarray(normals, dtype=np.float64).reshape(-1, n)
        ua = None if uniforms is None else np.ascontiguousarray(uniforms, dtype=np.float64).reshape(-1, n)
        L.check(L.lib().gb_pf_set_noise(
            self._h, None if na is None else na.ctypes.data_as(L.dp), 0 if na is None else na.shape[0],
            None if ua is None else ua.ctypes.data_as(L.dp), 0 if ua is None else ua.shape[0]))

    def set_resample_noise(self, u0=None, u64s=None):
        up = None
        if u64s is not None:
            u64s = np.ascontiguousarray(u64s, dtype=np.uint64)
            up = u64s.ctypes.data_as(L.u64p)
        L.check(L.lib().gb_pf_set_resample_noise(self._h, 0 if u0 is None else 1, 0 if u0 is None else int(u0), up))

    def set_resample_spacings(self, spacings):
        """Validation mode of the sorted multinomial resampler: n + 1 integer spacings for the next resample."""
        sp = None
        if spacings is not None:
            spacings = np.ascontiguousarray(spacings, dtype=np.uint64)
            assert spacings.size == self.num_particles + 1, "n + 1 spacings"
            sp = spacings.ctypes.data_as(L.u64p)
        L.check(L.lib().gb_pf_set_resample_spacings(self._h, sp))

    def enable_window(self, W):
        """Keep the last W states of every particle's own lineage (sliding-window rejuvenation); before generate."""
        L.check(L.lib().gb_pf_enable_window(self._h, int(W)))

    def enable_history(self, on=True):
        """Keep per-step snapshots + ancestors so that full trajectories can be reconstructed (call before generate)."""
        L.check(L.lib().gb_pf_enable_history(self._h, int(bool(on))))

    def trajectory(self, field):
        """[T+1, N] array: value of trace column `field` at every time step along each current particle's lineage."""
        steps = C.c_int()
        L.check(L.lib().gb_pf_history_length(self._h, C.byref(steps)))
        out = np.empty((steps.value, self.num_particles), dtype=np.float64)
        L.check(L.lib().gb_pf_get_trajectory(self._h, int(field), out.ctypes.data_as(L.dp)))
        return out

    @property
    def log_incremental_weights(self):
        """The return value of the last particle_filter_step_ (read on demand from the device)."""
        out = np.empty(self.num_particles, dtype=np.float64)
        L.check(L.lib().gb_pf_get_log_increments(self._h, out.ctypes.data_as(L.dp)))
        return out

    @property
    def spec_backend(self):
        """GB_MODEL_SPEC filters: "nvrtc" / "interpreter" for the kernel the last generate/step ran, else None."""
        b = C.c_int()
        L.check(L.lib().gb_pf_spec_backend(self._h, C.byref(b)))
        return {1: "nvrtc", 0: "interpreter"}.get(b.value)

    def set_lazy_gather(self, lazy):
        L.check(L.lib().gb_pf_set_lazy_gather(self._h, int(bool(lazy))))

    def set_stream(self, cuda_stream):
        L.check(L.lib().gb_pf_set_stream(self._h, C.c_void_p(cuda_stream)))

    def sync(self):
        L.check(L.lib().gb_pf_sync(self._h))

    def enable_timing(self, on=True):
        L.check(L.lib().gb_pf_enable_timing(self._h, int(on)))

    def get_timing(self):
        ms = (C.c_double * L.K_COUNT)()
        ln = (C.c_int64 * L.K_COUNT)()
        L.check(L.lib().gb_pf_get_timing(self._h, ms, ln))
        return {L.lib().gb_kernel_name(k).decode(): (ms[k], ln[k]) for k in range(L.K_COUNT)}

    # Base.copy (particle_filter.jl:43-51)
    def copy(self):
        h = C.c_void_p()
        L.check(L.lib().gb_pf_clone(self._h, C.byref(h)))
        c = DeviceParticleFilterState(h, self.model, self.dtype)
        c.resampler = self.resampler
        c._last_t = self._last_t
        return c

    __copy__ = copy

    # Base.:(==) (particle_filter.jl:52-58)
    def __eq__(self, other):
        if not isinstance(other, DeviceParticleFilterState):
            return NotImplemented
        if self.num_particles != other.num_particles or self.log_ml_est != other.log_ml_est:
            return False
        if not np.array_equal(self.log_weights, other.log_weights) or not np.array_equal(self.parents, other.parents):
            return False
        return all(np.array_equal(self.get_state(f), other.get_state(f)) for f in range(self.state_dim))

    # Base.hash (particle_filter.jl:59-63): over the same fields == compares, so equal states hash equal
    def __hash__(self):
        import hashlib
        h = hashlib.blake2b(digest_size=8)
        for f in range(self.state_dim):
            h.update(self.get_state(f).tobytes())
        h.update(self.log_weights.tobytes())
        h.update(np.float64(self.log_ml_est).tobytes())
        h.update(self.parents.tobytes())
        return int.from_bytes(h.digest(), "little", signed=True)

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                L.lib().gb_pf_free(h)
            except Exception:
                pass
            self._h = None


def initialize_particle_filter(model, model_args, observations, *rest, dtype="f64", seed=0):
    """initialize_particle_filter(model, model_args, observations[, proposal, proposal_args], num_particles)
    (particle_filter.jl:118-134, :142-154)."""
    if len(rest) == 1:
        proposal, proposal_args, num_particles = None, (), rest[0]
    elif len(rest) == 3:
        proposal, proposal_args, num_particles = rest
    else:
        raise TypeError("initialize_particle_filter(model, model_args, observations, [proposal, proposal_args,] num_particles)")
    obs = _flatten_obs(observations)
    pk, pargs = _proposal(proposal, proposal_args)
    pa, pp, npa = _darr(pargs)
    h = C.c_void_p()
    dt = _DTYPES[dtype]
    L.check(L.lib().gb_pf_initialize(model._h, int(num_particles), dt, int(seed), obs.ctypes.data_as(L.dp), obs.size,
                                     pk, pp, npa, C.byref(h)))
    st = DeviceParticleFilterState(h, model, dt)
    if len(model_args) >= 1 and isinstance(model_args[0], (int, np.integer)):
        st._last_t = int(model_args[0])
    return st


def create_particle_filter(model, num_particles, dtype="f64", seed=0, n_global=None, pid_offset=0):
    """Allocate an empty store (validation mode needs the handle before `generate` to inject noise)."""
    h = C.c_void_p()
    dt = _DTYPES[dtype]
    ng = int(num_particles) if n_global is None else int(n_global)
    L.check(L.lib().gb_pf_create(model._h, int(num_particles), ng, int(pid_offset), dt, int(seed), C.byref(h)))
    return DeviceParticleFilterState(h, model, dt)


def generate_(state, observations, proposal=None, proposal_args=()):
    """The t = 0 `generate` sweep of initialize_particle_filter on an allocated store."""
    obs = _flatten_obs(observations)
    pk, pargs = _proposal(proposal, proposal_args)
    pa, pp, npa = _darr(pargs)
    L.check(L.lib().gb_pf_generate(state._h, obs.ctypes.data_as(L.dp), obs.size, pk, pp, npa))
    return state


def particle_filter_step_(state, new_args, argdiffs, observations, proposal=None, proposal_args=(), return_incr=True):
    """particle_filter_step!(state, new_args, argdiffs, observations[, proposal, proposal_args])
    -> (log_incremental_weights,)   (particle_filter.jl:186-208, :217-240).
    Only extend-by-one is supported: new_args must be (t,) with t = previous t + 1; anything else would
    produce a non-empty discard, which the reference rejects (:227-229)."""
    if len(new_args) >= 1 and new_args[0] is not None and state._last_t is not None:
        if int(new_args[0]) != state._last_t + 1:
            raise L.GenB200Error(L.ESTATE, "Choices were updated or deleted inside particle filter step "
                                 "(only extension by one time step is supported)")
    obs = _flatten_obs(observations)
    pk, pargs = _proposal(proposal, proposal_args)
    pa, pp, npa = _darr(pargs)
    incr = np.empty(state.num_particles, dtype=np.float64) if return_incr else None
    L.check(L.lib().gb_pf_step(state._h, obs.ctypes.data_as(L.dp), obs.size, pk, pp, npa,
                               None if incr is None else incr.ctypes.data_as(L.dp)))
    if len(new_args) >= 1 and new_args[0] is not None:
        state._last_t = int(new_args[0])
    return (incr,)


def maybe_resample_(state, ess_threshold=None, verbose=False, resampler=None):
    """maybe_resample!(state; ess_threshold=num_particles/2, verbose=false) -> Bool (particle_filter.jl:249-275).
    `resampler` in {"multinomial" (reference semantics), "systematic", "residual"}."""
    n = state.num_particles
    thr = n / 2 if ess_threshold is None else float(ess_threshold)
    kind = _RESAMPLERS[resampler or state.resampler]
    did = C.c_int()
    if verbose:
        print("effective sample size: %s, doing resample: %s" % (state.ess(), state.ess() < thr))
    L.check(L.lib().gb_pf_maybe_resample(state._h, thr, kind, C.byref(did)))
    return bool(did.value)


def rejuvenate_(state, observations, move="resimulate", drift_std=0.0, return_alpha=False):
    """The tutorial's rejuvenation loop `for i in 1:N  state.traces[i], _ = mh(state.traces[i], ...)`
    (docs/src/tutorials/smc.md:459-463, :518-520) for the moves the Markov store supports: on the latent choices of
    the latest time step.  move = "resimulate": mh(trace, select(latest latents)) (mh.jl:16-27);
    move = "drift": mh(trace, gaussian_drift_proposal, (drift_std,)) (mh.jl:37-58).  Call between
    particle_filter_step_ and maybe_resample_; `observations` are the latest step's.
    Returns the number of accepted moves (and the log acceptance ratios with return_alpha=True)."""
    obs = _flatten_obs(observations)
    mv = {"resimulate": L.MH_RESIMULATE, "drift": L.MH_DRIFT}[move]
    acc = C.c_int64()
    alpha = np.empty(state.num_particles, dtype=np.float64) if return_alpha else None
    L.check(L.lib().gb_pf_rejuvenate(state._h, mv, obs.ctypes.data_as(L.dp), obs.size, float(drift_std), C.byref(acc),
                                     None if alpha is None else alpha.ctypes.data_as(L.dp)))
    return (acc.value, alpha) if return_alpha else acc.value


def rejuvenate_window_(state, a, b, observations_window, drift_std, return_alpha=False):
    """The tutorial's sliding-window rejuvenation `for i in 1:N  state.traces[i], _ = perturbation_move(state.traces[i], a, b)`
    (docs/src/tutorials/smc.md:492-541): one block Gaussian-drift MH move per particle on the latent choices of steps
    a..b (1 <= a <= b <= current step, at most W steps back: state.enable_window(W) before generate).
    `observations_window`: the observations of steps a..T.  Returns the number of accepted moves."""
    obs = np.ascontiguousarray([_flatten_obs(o)[0] for o in observations_window], dtype=np.float64)
    acc = C.c_int64()
    alpha = np.empty(state.num_particles, dtype=np.float64) if return_alpha else None
    L.check(L.lib().gb_pf_rejuvenate_window(state._h, int(a), int(b), obs.ctypes.data_as(L.dp), 1, float(drift_std), C.byref(acc),
                                            None if alpha is None else alpha.ctypes.data_as(L.dp)))
    return (acc.value, alpha) if return_alpha else acc.value


def particle_filter_run_(state, observations_seq, ess_threshold=None, rejuvenation=None):
    """T x { maybe_resample!(state, ess_threshold) [systematic]; particle_filter_step!(state, (t,), ..., obs_t) } in one
    call, without a host round trip per step (include/genb200.h gb_pf_run).  Returns the number of resamples.
    rejuvenation = (move, drift_std, sweeps) adds `sweeps` MH moves (rejuvenate_) after every step -- the tutorial's
    particle_filter_rejuv loop (docs/src/tutorials/smc.md:513-531); then (resamples, accepted moves) is returned."""
    obs = np.ascontiguousarray([_flatten_obs(o) for o in observations_seq], dtype=np.float64)
    T, nobs = obs.shape if obs.ndim == 2 else (len(obs), 1)
    thr = state.num_particles / 2 if ess_threshold is None else float(ess_threshold)
    nres = C.c_int()
    if rejuvenation is None:
        L.check(L.lib().gb_pf_run(state._h, int(T), obs.ctypes.data_as(L.dp), int(nobs), thr, C.byref(nres)))
        out = nres.value
    else:
        move, drift_std, sweeps = rejuvenation
        acc = C.c_int64()
        L.check(L.lib().gb_pf_run_mh(state._h, int(T), obs.ctypes.data_as(L.dp), int(nobs), thr,
                                     {"resimulate": L.MH_RESIMULATE, "drift": L.MH_DRIFT}[move], float(drift_std), int(sweeps),
                                     C.byref(nres), C.byref(acc)))
        out = (nres.value, acc.value)
    if state._last_t is not None:
        state._last_t += int(T)
    return out


def log_ml_estimate(state):
    """particle_filter.jl:91-94."""
    v = C.c_double()
    L.check(L.lib().gb_pf_log_ml_estimate(state._h, C.byref(v)))
    return v.value


def get_log_weights(state):
    """particle_filter.jl:82-84 (unnormalised log weights)."""
    out = np.empty(state.num_particles, dtype=np.float64)
    L.check(L.lib().gb_pf_get_log_weights(state._h, out.ctypes.data_as(L.dp)))
    return out


def get_traces(state):
    """particle_filter.jl:70-73."""
    return state.traces


def sample_unweighted_traces(state, num_samples, draw=0):
    """sample_unweighted_traces(state, num_samples) (particle_filter.jl:101-109): traces drawn iid from the weighted
    collection.  Returns (indices [1-based], rows [num_samples x state_dim])."""
    n = int(num_samples)
    idx = np.empty(n, dtype=np.int64)
    rows = np.empty((state.state_dim, n), dtype=np.float64)
    L.check(L.lib().gb_pf_sample_unweighted(state._h, n, int(draw), idx.ctypes.data_as(L.i64p), rows.ctypes.data_as(L.dp)))
    return idx, rows.T.copy()


def importance_resampling(model, model_args, observations, *rest, verbose=False, dtype="f64", seed=0, chunk=0):
    """(trace, lml_est) = importance_resampling(model, model_args, observations, [proposal, proposal_args,] num_samples)
    (importance.jl:77-122), streaming: the proposals are generated and reduced chunk by chunk on the device (O(chunk)
    memory for any num_samples) and folded by the reference's running logsumexp(x1, x2) + Bernoulli swap
    (include/genb200.h gb_importance_resampling).  `trace` is the kept particle's state row."""
    if len(rest) == 1:
        proposal, proposal_args, n = None, (), rest[0]
    elif len(rest) == 3:
        proposal, proposal_args, n = rest
    else:
        raise TypeError("importance_resampling(model, model_args, observations, [proposal, proposal_args,] num_samples)")
    obs = _flatten_obs(observations)
    pk, pargs = _proposal(proposal, proposal_args)
    pa, pp, npa = _darr(pargs)
    row = np.empty(model.state_dim, dtype=np.float64)
    lml = C.c_double()
    L.check(L.lib().gb_importance_resampling(model._h, int(n), _DTYPES[dtype], int(seed), obs.ctypes.data_as(L.dp), obs.size, pk, pp, npa,
                                             int(chunk), row.ctypes.data_as(L.dp), C.byref(lml)))
    return row, lml.value


def importance_sampling(model, model_args, observations, *rest, verbose=False, dtype="f64", seed=0,
                        normals=None, uniforms=None, return_traces=True):
    """(traces, log_norm_weights, lml_est) = importance_sampling(model, model_args, observations,
    [proposal, proposal_args,] num_samples)   (importance.jl:20-59)."""
    if len(rest) == 1:
        proposal, proposal_args, n = None, (), rest[0]
    elif len(rest) == 3:
        proposal, proposal_args, n = rest
    else:
        raise TypeError("importance_sampling(model, model_args, observations, [proposal, proposal_args,] num_samples)")
    n = int(n)
    obs = _flatten_obs(observations)
    pk, pargs = _proposal(proposal, proposal_args)
    pa, pp, npa = _darr(pargs)
    na = None if normals is None else np.ascontiguousarray(normals, dtype=np.float64).reshape(-1, n)
    ua = None if uniforms is None else np.ascontiguousarray(uniforms, dtype=np.float64).reshape(-1, n)
    lnw = np.empty(n, dtype=np.float64)
    lml = C.c_double()
    h = C.c_void_p()
    dt = _DTYPES[dtype]
    L.check(L.lib().gb_importance_sampling(
        model._h, n, dt, int(seed), obs.ctypes.data_as(L.dp), obs.size, pk, pp, npa,
        None if na is None else na.ctypes.data_as(L.dp), 0 if na is None else na.shape[0],
        None if ua is None else ua.ctypes.data_as(L.dp), 0 if ua is None else ua.shape[0],
        lnw.ctypes.data_as(L.dp), C.byref(lml), C.byref(h) if return_traces else None))
    traces = DeviceParticleFilterState(h, model, dt).traces if return_traces else None
    return traces, lnw, lml.value


# ---- raw-array entries (config 5) and the distribution kernels ----
def resample(logw, kind="systematic", u0=0, u64s=None, dtype="f64", seed=0, step=0):
    """0-based ancestors of N log weights by the engine's resampling kernels."""
    a, p, n = _darr(logw)
    out = np.empty(n, dtype=np.int64)
    up = None
    if u64s is not None:
        u64s = np.ascontiguousarray(u64s, dtype=np.uint64)
        up = u64s.ctypes.data_as(L.u64p)
    L.check(L.lib().gb_resample(p, n, _DTYPES[dtype], _RESAMPLERS[kind], int(u0), up, int(seed), int(step),
                                out.ctypes.data_as(L.i64p)))
    return out


def logsumexp(arr, dtype="f64"):
    """inference.jl:3-6 on the device."""
    a, p, n = _darr(arr)
    v = C.c_double()
    L.check(L.lib().gb_logsumexp(p, n, _DTYPES[dtype], C.byref(v), None))
    return v.value


def effective_sample_size_from_logw(arr, dtype="f64"):
    """particle_filter.jl:3-12 (normalize_weights + effective_sample_size) on the device."""
    a, p, n = _darr(arr)
    lse, ess = C.c_double(), C.c_double()
    L.check(L.lib().gb_logsumexp(p, n, _DTYPES[dtype], C.byref(lse), C.byref(ess)))
    return ess.value


_DISTS = {"normal": L.DIST_NORMAL, "gamma": L.DIST_GAMMA, "bernoulli": L.DIST_BERNOULLI,
          "categorical": L.DIST_CATEGORICAL, "mvnormal": L.DIST_MVNORMAL, "uniform": 6, "exponential": 7, "laplace": 8,
          "cauchy": 9, "inv_gamma": 10, "beta": 11, "poisson": 12, "geometric": 13, "binom": 14, "neg_binom": 15,
          "uniform_discrete": 16, "broadcasted_normal": 17, "dirichlet": 18, "piecewise_uniform": 19, "beta_uniform": 20}


def logpdf(dist, x, a, b=None, dtype="f64"):
    """Batched logpdf(dist, x, args...) (modeling_library.jl:22).  normal(mu, std), gamma(shape, scale),
    bernoulli(p), categorical(probs), mvnormal(mu, cov)."""
    k = _DISTS[dist]
    x = np.ascontiguousarray(x, dtype=np.float64)
    a = np.ascontiguousarray(a, dtype=np.float64)
    bb = None if b is None else np.ascontiguousarray(b, dtype=np.float64)
    if k in (L.DIST_MVNORMAL, 18):                 # mvnormal(mu, cov) / dirichlet(alpha): x is [n][len]
        ln = a.size
        n = x.size // ln
        ast = bst = 0
    elif k == L.DIST_CATEGORICAL:
        ln, n, ast, bst = a.size, x.size, 0, 0
    elif k == 19:                                  # piecewise_uniform(bounds, probs)
        ln, n, ast, bst = bb.size, x.size, 0, 0
        if a.size != ln + 1:
            raise L.GenB200Error(L.EINVAL, "Dimension mismatch")       # piecewise_uniform.jl:13-17
    elif k == 20:                                  # beta_uniform(theta, alpha, beta): b = (alpha, beta) pairs
        ln, n = 0, x.size
        ast = 0 if a.size == 1 else 1
        bst = 0 if bb.size == 2 else 1
    else:
        ln, n = 0, x.size
        ast = 0 if a.size == 1 else 1
        bst = 0 if (bb is None or bb.size == 1) else 1
        if k == 17 and n == 1 and (a.size > 1 or (bb is not None and bb.size > 1)):      # scalar x against array parameters
            n = max(a.size, 1 if bb is None else bb.size)
            x = np.full(n, float(x.ravel()[0]))
    out = np.empty(1 if k == 17 else n, dtype=np.float64)
    L.check(L.lib().gb_dist_logpdf(k, _DTYPES[dtype], n, x.ctypes.data_as(L.dp), a.ctypes.data_as(L.dp), ast,
                                   None if bb is None else bb.ctypes.data_as(L.dp), bst, ln, out.ctypes.data_as(L.dp)))
    return float(out[0]) if k == 17 else out


def random(dist, n, a, b=None, dtype="f64", seed=0, step=0):
    """n draws of random(dist, args...) (modeling_library.jl:15) from the engine's Philox stream."""
    k = _DISTS[dist]
    a = np.ascontiguousarray(a, dtype=np.float64)
    bb = None if b is None else np.ascontiguousarray(b, dtype=np.float64)
    if k in (L.DIST_MVNORMAL, L.DIST_CATEGORICAL, 18):
        ln, ast, bst = a.size, 0, 0
    elif k == 19:
        ln, ast, bst = bb.size, 0, 0
    elif k == 20:
        ln, ast, bst = 0, (0 if a.size == 1 else 1), (0 if bb.size == 2 else 1)
    else:
        ln = 0
        ast = 0 if a.size == 1 else 1
        bst = 0 if (bb is None or bb.size == 1) else 1
    rowwise = k in (L.DIST_MVNORMAL, 18)
    out = np.empty(n * (ln if rowwise else 1), dtype=np.float64)
    L.check(L.lib().gb_dist_random(k, _DTYPES[dtype], n, a.ctypes.data_as(L.dp), ast,
                                   None if bb is None else bb.ctypes.data_as(L.dp), bst, ln, int(seed), int(step),
                                   out.ctypes.data_as(L.dp)))
    return out.reshape(n, ln) if rowwise else out


def enumerative_normalize(log_weights, log_vols=None, dtype="f64"):
    """(log_normalized_weights, log_total_weight): the tail of enumerative_inference (enumerative.jl:39-46)."""
    lw = np.ascontiguousarray(log_weights, dtype=np.float64).ravel()
    lv = None if log_vols is None else np.ascontiguousarray(log_vols, dtype=np.float64).ravel()
    out = np.empty(lw.size, dtype=np.float64)
    tot = C.c_double()
    L.check(L.lib().gb_enumerative_normalize(lw.ctypes.data_as(L.dp), None if lv is None else lv.ctypes.data_as(L.dp), lw.size,
                                             _DTYPES[dtype], out.ctypes.data_as(L.dp), C.byref(tot)))
    return out.reshape(np.shape(log_weights)), tot.value


def philox_normals(seed, step, pid0, n, ndraw, dtype="f64"):
    out = np.empty((ndraw, n), dtype=np.float64)
    L.check(L.lib().gb_philox_normals(int(seed), int(step), int(pid0), int(n), int(ndraw), _DTYPES[dtype],
                                      out.ctypes.data_as(L.dp)))
    return out


def philox_uniforms(seed, step, pid0, n, ndraw, dtype="f64"):
    out = np.empty((ndraw, n), dtype=np.float64)
    L.check(L.lib().gb_philox_uniforms(int(seed), int(step), int(pid0), int(n), int(ndraw), _DTYPES[dtype],
                                       out.ctypes.data_as(L.dp)))
    return out


def philox_resample_u0(seed, step):
    return int(L.lib().gb_philox_resample_u0(int(seed), int(step)))


def philox_sorted_spacings(seed, step, k0, n):
    """The integer exponential spacings the sorted multinomial resampler derives from Philox (computed on the device)."""
    out = np.empty(n, dtype=np.uint64)
    L.check(L.lib().gb_philox_sorted_spacings(int(seed), int(step), int(k0), int(n), out.ctypes.data_as(L.u64p)))
    return out


def philox_multinomial_u64(seed, step, k0, n):
    out = np.empty(n, dtype=np.uint64)
    L.check(L.lib().gb_philox_multinomial_u64(int(seed), int(step), int(k0), int(n), out.ctypes.data_as(L.u64p)))
    return out


def cache_trim():
    """Return the allocator's cached device / pinned blocks to the driver."""
    L.check(L.lib().gb_cache_trim())


def cache_stats():
    """(cached bytes, cached blocks, live blocks) of the library's caching allocator."""
    a, b, c = C.c_int64(), C.c_int64(), C.c_int64()
    L.check(L.lib().gb_cache_stats(C.byref(a), C.byref(b), C.byref(c)))
    return a.value, b.value, c.value


def device_count():
    c = C.c_int()
    L.check(L.lib().gb_device_count(C.byref(c)))
    return c.value
