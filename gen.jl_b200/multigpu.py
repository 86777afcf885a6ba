"""One particle filter sharded over several GPUs from ONE process: the host mirror of the gb_spf_* entries of
include/genb200.h (csrc/spf.inc).  No torch, no NCCL: the shards exchange through peer memory inside the kernels.

    st = multigpu.initialize_particle_filter(model, (0,), obs0, n, devices=[0, 1, 2, 3])
    multigpu.maybe_resample_(st); multigpu.particle_filter_step_(st, (1,), (UnknownChange,), obs1)
    multigpu.log_ml_estimate(st)

Same call shapes as api.py (Gen's particle_filter.jl); results are bit-identical to the single-GPU filter for every
shard count.  `devices` may repeat an index (several shards on one GPU), which is how a single-GPU box tests this path.
"""
import ctypes as C

import numpy as np

from . import _lib as L
from . import api


class MultiGpuParticleFilterState:
    """ParticleFilterState (particle_filter.jl:35-41) sharded over the devices behind a gb_spf_t handle."""

    def __init__(self, handle, model, dtype, num_particles):
        self._h = handle
        self.model = model
        self.dtype = dtype
        self.num_particles = int(num_particles)
        self._last_t = None

    @property
    def num_shards(self):
        v = C.c_int()
        L.check(L.lib().gb_spf_num_shards(self._h, C.byref(v)))
        return v.value

    @property
    def state_dim(self):
        return self.model.state_dim

    @property
    def log_weights(self):
        out = np.empty(self.num_particles, dtype=np.float64)
        L.check(L.lib().gb_spf_get_log_weights(self._h, out.ctypes.data_as(L.dp)))
        return out

    @property
    def log_incremental_weights(self):
        out = np.empty(self.num_particles, dtype=np.float64)
        L.check(L.lib().gb_spf_get_log_increments(self._h, out.ctypes.data_as(L.dp)))
        return out

    @property
    def parents(self):
        out = np.empty(self.num_particles, dtype=np.int64)
        L.check(L.lib().gb_spf_get_parents(self._h, out.ctypes.data_as(L.i64p)))
        return out

    @property
    def log_ml_est(self):
        v = C.c_double()
        L.check(L.lib().gb_spf_get_log_ml_est(self._h, C.byref(v)))
        return v.value

    def get_state(self, field):
        out = np.empty(self.num_particles, dtype=np.float64)
        L.check(L.lib().gb_spf_get_state(self._h, int(field), out.ctypes.data_as(L.dp)))
        return out

    def ess(self):
        e = C.c_double()
        L.check(L.lib().gb_spf_get_ess(self._h, C.byref(e), None))
        return e.value

    def set_resample_noise(self, u0=None):
        L.check(L.lib().gb_spf_set_resample_noise(self._h, 0 if u0 is None else 1, 0 if u0 is None else int(u0)))

    def shard(self, r):
        """(borrowed gb_pf_t handle, pid_offset) of shard r -- for instrumentation (gb_pf_enable_timing ...)."""
        h, off = C.c_void_p(), C.c_int64()
        L.check(L.lib().gb_spf_shard(self._h, int(r), C.byref(h), C.byref(off)))
        return h, off.value

    def enable_timing(self, on=True, shard=0):
        L.check(L.lib().gb_pf_enable_timing(self.shard(shard)[0], 1 if on else 0))

    def get_timing(self, shard=0):
        ms = (C.c_double * L.K_COUNT)()
        ln = (C.c_int64 * L.K_COUNT)()
        L.check(L.lib().gb_pf_get_timing(self.shard(shard)[0], ms, ln))
        return {L.lib().gb_kernel_name(k).decode(): (ms[k], ln[k]) for k in range(L.K_COUNT)}

    def sync(self):
        L.check(L.lib().gb_spf_sync(self._h))

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                L.lib().gb_spf_free(h)
            except Exception:
                pass


def _devices(devices, nshards):
    if devices is None:
        return None, int(nshards)
    arr = (C.c_int * len(devices))(*[int(d) for d in devices])
    return arr, len(devices)


def initialize_particle_filter(model, model_args, observations, *rest, dtype="f64", seed=0, devices=None, nshards=None):
    """initialize_particle_filter(model, model_args, observations[, proposal, proposal_args], num_particles)
    (particle_filter.jl:118-154) over `devices` (or nshards devices 0..nshards-1)."""
    if len(rest) == 1:
        proposal, proposal_args, n = None, (), rest[0]
    elif len(rest) == 3:
        proposal, proposal_args, n = rest
    else:
        raise TypeError("initialize_particle_filter(model, model_args, observations, [proposal, proposal_args,] num_particles)")
    obs = api._flatten_obs(observations)
    pk, pargs = api._proposal(proposal, proposal_args)
    pa, pp, npa = api._darr(pargs)
    dev, G = _devices(devices, nshards if nshards is not None else 1)
    dt = api._DTYPES[dtype]
    h = C.c_void_p()
    L.check(L.lib().gb_spf_initialize(model._h, int(n), dt, int(seed), G, dev, obs.ctypes.data_as(L.dp), obs.size, pk, pp, npa,
                                      C.byref(h)))
    st = MultiGpuParticleFilterState(h, model, dt, n)
    if len(model_args) >= 1 and isinstance(model_args[0], (int, np.integer)):
        st._last_t = int(model_args[0])
    return st


def particle_filter_step_(state, new_args, argdiffs, observations, proposal=None, proposal_args=()):
    """particle_filter_step! (particle_filter.jl:186-240), extend-by-one only; the log incremental weights stay on the
    devices (state.log_incremental_weights reads them)."""
    if len(new_args) >= 1 and new_args[0] is not None and state._last_t is not None and int(new_args[0]) != state._last_t + 1:
        raise L.GenB200Error(L.ESTATE, "Choices were updated or deleted inside particle filter step "
                             "(only extension by one time step is supported)")
    obs = api._flatten_obs(observations)
    pk, pargs = api._proposal(proposal, proposal_args)
    pa, pp, npa = api._darr(pargs)
    L.check(L.lib().gb_spf_step(state._h, obs.ctypes.data_as(L.dp), obs.size, pk, pp, npa))
    if len(new_args) >= 1 and new_args[0] is not None:
        state._last_t = int(new_args[0])


def maybe_resample_(state, ess_threshold=None, verbose=False):
    """maybe_resample! (particle_filter.jl:249-275), global-prefix systematic resampling."""
    thr = state.num_particles / 2 if ess_threshold is None else float(ess_threshold)
    if verbose:
        print("effective sample size: %s, doing resample: %s" % (state.ess(), state.ess() < thr))
    did = C.c_int()
    L.check(L.lib().gb_spf_maybe_resample(state._h, thr, C.byref(did)))
    return bool(did.value)


def particle_filter_run_(state, observations_seq, ess_threshold=None):
    """T x { maybe_resample!; particle_filter_step! } with every decision and exchange on the devices (gb_spf_run)."""
    obs = np.ascontiguousarray([api._flatten_obs(o) for o in observations_seq], dtype=np.float64)
    T, nobs = obs.shape if obs.ndim == 2 else (len(obs), 1)
    thr = state.num_particles / 2 if ess_threshold is None else float(ess_threshold)
    nres = C.c_int()
    L.check(L.lib().gb_spf_run(state._h, int(T), obs.ctypes.data_as(L.dp), int(nobs), thr, C.byref(nres)))
    if state._last_t is not None:
        state._last_t += int(T)
    return nres.value


def log_ml_estimate(state):
    v = C.c_double()
    L.check(L.lib().gb_spf_log_ml_estimate(state._h, C.byref(v)))
    return v.value


def get_log_weights(state):
    return state.log_weights


def importance_sampling(model, model_args, observations, *rest, dtype="f64", seed=0, devices=None, nshards=None,
                        return_weights=True):
    """importance_sampling(model, model_args, observations[, proposal, proposal_args], num_samples)
    -> (state, log_normalized_weights, log_ml_estimate)   (importance.jl:20-59), samples sharded over the devices."""
    if len(rest) == 1:
        proposal, proposal_args, n = None, (), rest[0]
    else:
        proposal, proposal_args, n = rest
    obs = api._flatten_obs(observations)
    pk, pargs = api._proposal(proposal, proposal_args)
    pa, pp, npa = api._darr(pargs)
    dev, G = _devices(devices, nshards if nshards is not None else 1)
    dt = api._DTYPES[dtype]
    lnw = np.empty(int(n), dtype=np.float64) if return_weights else None
    lml, h = C.c_double(), C.c_void_p()
    L.check(L.lib().gb_spf_importance_sampling(model._h, int(n), dt, int(seed), G, dev, obs.ctypes.data_as(L.dp), obs.size, pk, pp, npa,
                                               None if lnw is None else lnw.ctypes.data_as(L.dp), C.byref(lml), C.byref(h)))
    return MultiGpuParticleFilterState(h, model, dt, n), lnw, lml.value


def resample(logw, u0=0, dtype="f64", devices=None, nshards=None):
    """Global-prefix systematic resampling of a raw log-weight array over the devices (cfg 5); 0-based ancestors."""
    a = np.ascontiguousarray(logw, dtype=np.float64)
    dev, G = _devices(devices, nshards if nshards is not None else 1)
    anc = np.empty(a.size, dtype=np.int64)
    L.check(L.lib().gb_spf_resample(a.ctypes.data_as(L.dp), a.size, api._DTYPES[dtype], G, dev, int(u0), anc.ctypes.data_as(L.i64p)))
    return anc
