"""World-size-2 (and 3) gloo tests of the sharded maybe_resample! sequencing (gen.jl_b200/sharded.py) on the
CPU: the collective logic, the exchange plan and the all-to-all-v are the product's; the shard itself is
an oracle-backed test double (the CUDA shard needs a GPU and is covered by tests/test_gpu_sharded.py).
The result must equal the single-process oracle resample bit for bit (global-prefix systematic)."""
import math
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class OracleShard:
    """Test double for ShardOps: the shard's arithmetic restated with the oracle + Python big integers."""

    def __init__(self, spf, O, logw, state_cols, log_ml_est):
        self.spf, self.O = spf, O
        lo, hi = spf.offs[spf.rank], spf.offs[spf.rank + 1]
        self.logw = np.array(logw[lo:hi])
        self.cols = [np.array(c[lo:hi]) for c in state_cols]
        self.state_dim = len(state_cols)
        self.new_cols = [np.empty_like(c) for c in self.cols]
        self.parents = np.arange(lo + 1, hi + 1, dtype=np.int64)
        self.new_parents = np.empty_like(self.parents)
        self._lml = log_ml_est
        self.anc = None

    def weight_max(self):
        return float(self.logw.max())

    def weight_sums(self, m, bits):
        e = np.exp(self.logw - m)
        q, tot = self.O.quantized_weights(self.logw, b=bits, m=m)
        self.q, self.m, self.bits = [int(v) for v in q], m, bits
        return float(e.sum()), float((e * e).sum()), int(tot)

    def block_bytes(self, count):
        return (count * (8 + 8 * self.state_dim) + 15) // 16 * 16

    def systematic_offspring(self, m, bits, q_offset, q_local, q_total, u0):
        assert q_local == sum(self.q)
        n = self.spf.n_global

        def below(Cv):
            if Cv == 0:
                return 0
            X = (Cv * (n << 32) - 1) // q_total
            return 0 if X < u0 else ((X - u0) >> 32) + 1
        cdf = q_offset
        lo = below(cdf)
        anc = []
        for j, qj in enumerate(self.q):
            k0 = below(cdf)
            cdf += qj
            anc += [j] * (below(cdf) - k0)
        self.anc, self.off_lo = np.array(anc, dtype=np.int64), lo
        return lo, lo + len(anc)

    def export_offspring(self, a, b, block):
        # block = [parents int64 x c][rows float64 x S x c] (include/genb200.h "Exchange block")
        idx = self.anc[a - self.off_lo:b - self.off_lo]
        c = b - a
        buf = block.numpy()
        buf[:8 * c].view(np.int64)[:] = self.spf.offs[self.spf.rank] + idx + 1
        rows = buf[8 * c:8 * c + 8 * c * self.state_dim].view(np.float64)
        for s in range(self.state_dim):
            rows[s * c:(s + 1) * c] = self.cols[s][idx]

    def import_rows(self, dst_lo, count, block):
        buf = block.numpy()
        self.new_parents[dst_lo:dst_lo + count] = buf[:8 * count].view(np.int64)
        rows = buf[8 * count:8 * count + 8 * count * self.state_dim].view(np.float64)
        for s in range(self.state_dim):
            self.new_cols[s][dst_lo:dst_lo + count] = rows[s * count:(s + 1) * count]

    def commit(self, incr):
        self.cols, self.new_cols = self.new_cols, self.cols
        self.parents = self.new_parents.copy()
        self.logw = np.zeros_like(self.logw)
        self._lml += incr

    def log_ml_est(self):
        return self._lml


def _worker(rank, world, port, n, shape, out_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import __graft_entry__ as entry
        import fixtures as F
        g = entry.load_package()
        O = entry.load_oracle()
        from genb200 import sharded
        rng = np.random.default_rng(123)                      # identical "global" filter in every process
        logw = F.resample_logw(n, shape)
        cols = [rng.standard_normal(n) for _ in range(3)]
        seed, step = 77, 5
        spf = sharded.ShardedParticleFilter(None, n, seed=seed,
                                            ops_factory=lambda s: OracleShard(s, O, logw, cols, -1.25))
        spf.step = step
        lml_before = spf.log_ml_estimate()
        did_skip = spf.maybe_resample_(ess_threshold=0.0)      # never resample
        did = spf.maybe_resample_(ess_threshold=n + 1)         # always resample
        # single-process reference: oracle systematic resample with the same u0
        u0 = O.lib().go_philox_resample_u0(seed, step)
        anc = O.resample(O.RESAMPLE_SYSTEMATIC, logw, u0=u0)
        lo, hi = spf.offs[rank], spf.offs[rank + 1]
        ok = (not did_skip) and did
        ok = ok and np.array_equal(spf.ops.parents, anc[lo:hi] + 1)
        for s in range(3):
            ok = ok and np.array_equal(spf.ops.cols[s], cols[s][anc[lo:hi]])
        lse = O.logsumexp(logw)
        ok = ok and abs(lml_before - (-1.25 + lse - math.log(n))) < 1e-12
        ok = ok and abs(spf.ops.log_ml_est() - (-1.25 + lse - math.log(n))) < 1e-12
        ok = ok and abs(spf.log_ml_estimate() - spf.ops.log_ml_est()) < 1e-12     # zero weights: lse = log N
        out_q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n,shape", [(2, 5000, "normal"), (2, 4097, "onehot"), (3, 1000, "normal"), (2, 64, "equal")])
def test_sharded_resample_matches_single_process(world, n, shape):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, shape, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
    res = sorted(q.get(timeout=5) for _ in range(world))
    assert res == [(r, True) for r in range(world)]
    assert all(p.exitcode == 0 for p in procs)


def test_plan_exchange_covers_every_slot_once():
    import __graft_entry__ as entry
    entry.load_package()
    from genb200 import sharded
    rng = np.random.default_rng(0)
    for world in (1, 2, 4, 8):
        n = 100003
        counts, offs = sharded.shard_bounds(n, world)
        assert sum(counts) == n and offs[-1] == n and max(counts) - min(counts) <= 1
        q = [int(v) for v in rng.integers(1, 2 ** 50, size=world)]
        if world > 2:
            q[1] = 0                                            # a shard without weight
        qt, pref, ranges, send = sharded.plan_exchange(q, n, offs, u0=123456)
        assert ranges[0][0] == 0 and ranges[-1][1] == n
        assert all(ranges[r][1] == ranges[r + 1][0] for r in range(world - 1))
        for d in range(world):                                  # every destination slot is received exactly once
            got = sum(send[s][d][1] - send[s][d][0] for s in range(world))
            assert got == counts[d]


def test_plan_exchange_matches_oracle_ancestors_property(O):
    """Property (hypothesis): for arbitrary weights, shard counts and u0, the slot ranges the host plan assigns to the
    shards are exactly the runs of the oracle's global systematic ancestors that fall into each shard."""
    from hypothesis import given, settings, strategies as st
    import __graft_entry__ as entry
    entry.load_package()
    from genb200 import sharded

    @settings(max_examples=60, deadline=None)
    @given(st.integers(1, 400), st.integers(1, 6), st.integers(0, 2 ** 32 - 1), st.integers(0, 2 ** 31),
           st.sampled_from(["normal", "skewed", "equal", "onehot"]))
    def check(n, world, u0, seed, shape):
        world = min(world, n)
        rng = np.random.default_rng(seed)
        if shape == "normal":
            lw = 3.0 * rng.standard_normal(n)
        elif shape == "skewed":
            lw = -rng.exponential(30.0, n)
        elif shape == "equal":
            lw = np.zeros(n)
        else:
            lw = np.full(n, -60.0)
            lw[rng.integers(0, n)] = 0.0
        counts, offs = sharded.shard_bounds(n, world)
        q, total = O.quantized_weights(lw)
        q_locals = [int(sum(int(v) for v in q[offs[r]:offs[r + 1]])) for r in range(world)]
        assert sum(q_locals) == total
        anc = O.resample(O.RESAMPLE_SYSTEMATIC, lw, u0=u0)
        qt, pref, ranges, send = sharded.plan_exchange(q_locals, n, offs, u0)
        owner = np.searchsorted(np.array(offs[1:]), anc, side="right")      # shard holding each slot's ancestor
        for r in range(world):
            lo, hi = ranges[r]
            assert np.all(owner[lo:hi] == r) and np.count_nonzero(owner == r) == hi - lo
            for d in range(world):                                            # and what r sends to d is d's part of it
                a, b = send[r][d]
                assert b - a == np.count_nonzero((owner == r) & (np.arange(n) >= offs[d]) & (np.arange(n) < offs[d + 1]))
    check()
