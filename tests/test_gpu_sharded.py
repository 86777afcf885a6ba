"""The sharded pieces of the C ABI on one GPU: K shards of one filter live in the same process and the
host plays the collectives (sums / prefix / row exchange by device copies).  The resampled filter must be
bit-identical to the single-shard filter and to the oracle: global-prefix systematic resampling is
partition independent by construction.  (Real multi-process NCCL runs: tests/run_sharded_multiproc.py under
`gpurun --gpus 2`, and bench.py --gpus N.)"""
import math

import numpy as np
import pytest

import fixtures as F

pytestmark = pytest.mark.gpu


def run_sharded_in_process(g, model, n, K, ys, seed, dtype="f64", fast=False, mh=None):
    import torch
    from genb200 import sharded, _lib as L
    counts, offs = sharded.shard_bounds(n, K)
    bits = L.lib().gb_quant_bits(n)
    shards = [g.create_particle_filter(model, counts[r], dtype=dtype, seed=seed, n_global=n, pid_offset=offs[r]) for r in range(K)]
    ops = [sharded.ShardOps(s) for s in shards]
    for s in shards:
        g.generate_(s, ys[0])
    S = model.state_dim
    td = torch.float64 if dtype == "f64" else torch.float32
    lml_hist = []
    for t in range(1, len(ys)):
        m = max(o.weight_max() for o in ops)                                  # all-reduce(max)
        trip = [o.weight_sums(m, bits) for o in ops]                          # all-gather
        S1, S2 = sum(x[0] for x in trip), sum(x[1] for x in trip)
        lse, ess = m + math.log(S1), S1 * S1 / S2
        if ess < n / 2 or t % 2 == 0:
            u0 = g.philox_resample_u0(seed, t - 1)
            qt, pref, ranges, send = sharded.plan_exchange([x[2] for x in trip], n, offs, u0)
            bufs = {}
            use_fast = fast and K > 1 and all(
                sum(send[s_][d][1] - send[s_][d][0] for s_ in range(K) if s_ != d) <= ops[d].ext_capacity() for d in range(K))
            for r in range(K):
                lo, hi = ops[r].systematic_offspring(m, bits, pref[r], trip[r][2], qt, u0)
                assert (lo, hi) == ranges[r]
                for d in range(K):
                    a, b = send[r][d]
                    if use_fast and d == r:
                        ops[r].offspring_local(a, b)
                    elif b > a:
                        blk = torch.empty(ops[r].block_bytes(b - a), dtype=torch.uint8, device="cuda")
                        ops[r].export_offspring(a, b, blk)
                        bufs[(r, d)] = (a, b, blk)
            ext_off = [0] * K
            for (r, d), (a, b, blk) in sorted(bufs.items()):                 # all-to-all-v
                if use_fast:
                    ops[d].import_ext(a - offs[d], b - a, ext_off[d], blk)
                    ext_off[d] += b - a
                else:
                    ops[d].import_rows(a - offs[d], b - a, blk)
            for o in ops:
                (o.commit_lazy if use_fast else o.commit)(lse - math.log(n))
        for s in shards:
            g.particle_filter_step_(s, (t,), (g.UnknownChange,), ys[t], return_incr=False)
            if mh is not None:                                                # rejuvenation is shard-local
                mh.append(g.rejuvenate_(s, ys[t], move="drift" if t % 2 else "resimulate", drift_std=1e-3))
        m = max(o.weight_max() for o in ops)
        trip = [o.weight_sums(m, bits) for o in ops]
        lml_hist.append(shards[0].log_ml_est + m + math.log(sum(x[0] for x in trip)) - math.log(n))
    cols = [np.concatenate([s.get_state(f) for s in shards]) for f in range(S)]
    lw = np.concatenate([g.get_log_weights(s) for s in shards])
    par = np.concatenate([s.parents for s in shards])
    return cols, lw, par, lml_hist


@pytest.mark.parametrize("fast", [False, True])
def test_sharded_rejuvenation_equals_single_shard(g, fast):
    """MH rejuvenation after each step is shard-local (parents are resident on the shard, through the extension rows
    when offspring migrated in) and keyed by the global particle id: K shards == 1 shard, bit for bit."""
    model, ys, n, seed = g.bearings_model(), F.bearings_series(7), 40009, 5
    acc3, acc1 = [], []
    cols, lw, par, lml = run_sharded_in_process(g, model, n, 3, ys, seed, fast=fast, mh=acc3)
    cols1, lw1, par1, lml1 = run_sharded_in_process(g, model, n, 1, ys, seed, mh=acc1)
    assert all(np.array_equal(a, b) for a, b in zip(cols, cols1))
    assert np.array_equal(lw, lw1) and np.array_equal(par, par1)
    assert [sum(acc3[3 * k:3 * k + 3]) for k in range(len(acc1))] == acc1 and sum(acc1) > 0
    # and it changed the filter
    cols0, _, _, _ = run_sharded_in_process(g, model, n, 1, ys, seed)
    assert not np.array_equal(cols0[2], cols1[2])


def test_sharded_spec_model_equals_builtin_kind(g):
    """A GB_MODEL_SPEC program (NVRTC-specialised) on a sharded filter: Philox is keyed by the global particle id and the
    extension rows work the same way, so 3 shards of the spec model == 1 shard of the hand-written kind."""
    from test_gpu_parity import bearings_spec
    ys, n, seed = F.bearings_series(7), 40009, 12
    cols, lw, par, lml = run_sharded_in_process(g, bearings_spec(g), n, 3, ys, seed, fast=True)
    cols1, lw1, par1, lml1 = run_sharded_in_process(g, g.bearings_model(), n, 1, ys, seed)
    assert all(np.array_equal(a, b) for a, b in zip(cols, cols1)) and np.array_equal(par, par1)
    assert np.allclose(lw, lw1, rtol=1e-10, atol=1e-10) and np.allclose(lml, lml1, rtol=1e-10)


@pytest.mark.parametrize("K", [2, 3, 8])
@pytest.mark.parametrize("name", ["bearings", "lgssm"])
def test_sharded_equals_single_shard_and_oracle(g, O, K, name):
    if name == "bearings":
        model, okind, params, ys = g.bearings_model(), O.MODEL_BEARINGS, F.BEARINGS_PARAMS, F.bearings_series(9)
    else:
        model, okind, params, ys = g.lgssm_model(**F.LGSSM), O.MODEL_LGSSM, F.lgssm_params(), F.lgssm_series(9)
    n, seed = 50021, 31
    cols, lw, par, lml = run_sharded_in_process(g, model, n, K, ys, seed)
    # the fast path (local ancestors + extension rows + deferred gather) gives the same filter
    colsf, lwf, parf, lmlf = run_sharded_in_process(g, model, n, K, ys, seed, fast=True)
    assert all(np.array_equal(a, b) for a, b in zip(cols, colsf))
    assert np.array_equal(lw, lwf) and np.array_equal(par, parf) and lml == lmlf
    # single shard through the same pieces (K = 1) and through the ordinary API
    cols1, lw1, par1, lml1 = run_sharded_in_process(g, model, n, 1, ys, seed)
    for a, b in zip(cols, cols1):
        assert np.array_equal(a, b)
    assert np.array_equal(lw, lw1) and np.array_equal(par, par1)
    assert np.allclose(lml, lml1, rtol=1e-12, atol=1e-12)
    # the oracle on the exported Philox noise
    nn0, nn1 = model.num_normals(True), model.num_normals(False)
    ref = O.OraclePF.init(okind, params, n, [ys[0]], normals=g.philox_normals(seed, 0, 0, n, nn0))
    for t in range(1, len(ys)):
        lse, lnw = O.normalize_weights(ref.log_weights)
        thr = n + 1 if t % 2 == 0 else n / 2
        ref.maybe_resample(ess_threshold=thr, kind=O.RESAMPLE_SYSTEMATIC, u0=g.philox_resample_u0(seed, t - 1))
        ref.step([ys[t]], normals=g.philox_normals(seed, t, 0, n, nn1))
    for f in range(model.state_dim):
        assert np.array_equal(cols[f], ref.state(f))
    assert np.array_equal(par, ref.parents)
    assert np.allclose(lw, ref.log_weights, rtol=1e-10, atol=1e-10)
    assert abs(lml[-1] - ref.log_ml_estimate()) <= 1e-10 * abs(ref.log_ml_estimate())


def test_sharded_class_world1(g, O):
    """ShardedParticleFilter without a process group (world = 1) equals the plain API."""
    import torch
    from genb200 import sharded
    ys = F.bearings_series(6)
    n, seed = 30000, 5
    spf = sharded.ShardedParticleFilter(g.bearings_model(), n, seed=seed).initialize(ys[0])
    from genb200 import _lib as L
    for c in (0, 1, 7, 1000):
        assert spf.ops.block_bytes(c) == L.lib().gb_pf_block_bytes(spf.local._h, c)
    st = g.initialize_particle_filter(g.bearings_model(), (0,), ys[0], n, seed=seed)
    for t in range(1, len(ys)):
        a = spf.maybe_resample_(ess_threshold=n + 1)
        b = g.maybe_resample_(st, ess_threshold=n + 1, resampler="systematic")
        assert a == b
        spf.particle_filter_step_((t,), ys[t])
        g.particle_filter_step_(st, (t,), (g.UnknownChange,), ys[t])
    for f in range(4):
        assert np.array_equal(spf.get_state(f), st.get_state(f))
    assert np.array_equal(spf.get_parents(), st.parents)
    assert abs(spf.log_ml_estimate() - g.log_ml_estimate(st)) < 1e-10


def run_async_in_process(g, model, n, K, ys, seed, thr):
    """K shards in one process (one GPU, one stream) driven through the ASYNCHRONOUS sharded entries.  Nothing is played
    by the host: the shards' maxima, sums and "rows delivered" flags go through the peer-memory mailboxes, offspring
    through the export kernel's peer stores (here: pointers of the other handles in the same process).  With one
    stream the host enqueues phase by phase over the shards, so every post precedes the wait that needs it."""
    from genb200 import sharded
    counts, offs = sharded.shard_bounds(n, K)
    shards = [g.create_particle_filter(model, counts[r], seed=seed, n_global=n, pid_offset=offs[r]) for r in range(K)]
    ops = [sharded.ShardOps(s) for s in shards]
    for s in shards:
        g.generate_(s, ys[0])
    for r, o in enumerate(ops):
        for d, p in enumerate(ops):
            if d != r:
                o.attach_local(d, p)
    for r, o in enumerate(ops):
        o.async_setup(K, r, offs, thr)
    for t in range(1, len(ys)):
        u0 = g.philox_resample_u0(seed, t - 1)
        for phase in ("xchg_post", "xchg_stats", ("async_plan", u0, float("nan"), 1), "async_sweep", "async_export", "async_wait",
                      ("async_step", ys[t])):
            for o in ops:
                if isinstance(phase, tuple):
                    getattr(o, phase[0])(*phase[1:])
                else:
                    getattr(o, phase)()
    nres = [o.async_finish() for o in ops]
    assert len(set(nres)) == 1
    cols = [np.concatenate([s.get_state(f) for s in shards]) for f in range(model.state_dim)]
    lw = np.concatenate([g.get_log_weights(s) for s in shards])
    par = np.concatenate([s.parents for s in shards])
    return cols, lw, par, shards[0].log_ml_est, nres[0]


@pytest.mark.parametrize("K", [1, 2, 3, 8])
def test_async_sharded_run_equals_single_gpu(g, K):
    model, ys = g.bearings_model(), F.bearings_series(9)
    n, seed = 60013, 23
    for thr in (n / 2, n + 1):
        ref = g.initialize_particle_filter(model, (0,), ys[0], n, seed=seed)
        nres_ref = g.particle_filter_run_(ref, ys[1:], ess_threshold=thr)
        cols, lw, par, lml, nres = run_async_in_process(g, model, n, K, ys, seed, thr)
        assert nres == nres_ref
        for f in range(4):
            assert np.array_equal(cols[f], ref.get_state(f))
        assert np.array_equal(lw, ref.log_weights) and np.array_equal(par, ref.parents)
        assert abs(lml - ref.log_ml_est) <= 1e-12 * max(1.0, abs(lml))


def test_async_sharded_overflow_is_reported(g):
    """A shard that would receive more rows than its extension rows hold (here: all the weight sits in shard 0) must make
    the asynchronous run fail loudly, never silently."""
    import torch
    from genb200 import sharded, _lib as L
    model, ys = g.lgssm_model(**F.LGSSM), F.lgssm_series(3)
    n, K = 40000, 2
    counts, offs = sharded.shard_bounds(n, K)
    shards = [g.create_particle_filter(model, counts[r], seed=1, n_global=n, pid_offset=offs[r]) for r in range(K)]
    ops = [sharded.ShardOps(s) for s in shards]
    for s in shards:
        g.generate_(s, ys[0])
    shards[1].log_weights = np.full(counts[1], -800.0)            # shard 1 has no weight: all its slots come from shard 0
    for r, o in enumerate(ops):
        o.attach_local(1 - r, ops[1 - r])
    for r, o in enumerate(ops):
        o.async_setup(K, r, offs, n + 1)
    for phase in ("xchg_post", "xchg_stats", ("async_plan", 7, float("nan"), 1), "async_sweep", "async_export", "async_wait",
                  ("async_step", ys[1])):
        for o in ops:
            if isinstance(phase, tuple):
                getattr(o, phase[0])(*phase[1:])
            else:
                getattr(o, phase)()
    with pytest.raises(g.GenB200Error):
        ops[0].async_finish()
