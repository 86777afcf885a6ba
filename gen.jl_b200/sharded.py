"""Particle filter sharded over GPUs: one process per GPU (torch.distributed, NCCL over NVLink).

Particles are split into contiguous shards (global id = shard offset + local index).  Propose / weight
need no communication.  maybe_resample! needs exactly

  * one all-reduce(max) of the shard maxima (8 bytes),
  * one all-gather of (sum exp(lw-m), sum exp(2(lw-m)), integer CDF mass) per shard (24 bytes) -> global
    log-sum-exp, ESS and the shard's offset into the global integer CDF (global-prefix systematic
    resampling: every shard sweeps its own particles against the SAME global slot grid, so the ancestors
    are bit-identical to the single-GPU / sequential result by construction),
  * one all-to-all-v of the offspring rows whose destination slot lives on another shard.

All heavy lifting is in libgenb200 (include/genb200.h "sharded pieces"); this file only sequences the
collectives.  `ShardOps` is the interface to a shard; the CPU test-suite substitutes an oracle-backed
implementation to exercise the collective logic with the gloo backend (tests/test_sharded_gloo.py).
"""
import ctypes as C
import math

import numpy as np

from . import _lib as L
from . import api


def shard_bounds(n_global, world):
    """Contiguous shards; the first (n_global % world) shards hold one extra particle."""
    base, rem = divmod(int(n_global), int(world))
    counts = [base + (1 if r < rem else 0) for r in range(world)]
    offs = [0]
    for c in counts:
        offs.append(offs[-1] + c)
    return counts, offs


def slots_below(C_, n_global, q_total, u0):
    """#{k : floor((k 2^32 + u0) Q / (N 2^32)) < C}: host instance of the kernels' exact integer function."""
    return int(L.lib().gb_selftest_slots_below(int(C_), int(n_global), int(q_total), int(u0)))


def plan_exchange(q_locals, n_global, offs, u0):
    """From the gathered per-shard integer CDF masses: every shard's offspring slot range and the matrix
    send[src][dst] = (k_lo, k_hi) of global slots produced by src that are owned by dst."""
    world = len(q_locals)
    q_total = int(sum(int(q) for q in q_locals))
    pref = [0]
    for q in q_locals:
        pref.append(pref[-1] + int(q))
    ranges = [(slots_below(pref[r], n_global, q_total, u0), slots_below(pref[r + 1], n_global, q_total, u0))
              for r in range(world)]
    send = [[(0, 0)] * world for _ in range(world)]
    for s, (lo, hi) in enumerate(ranges):
        for d in range(world):
            a, b = max(lo, offs[d]), min(hi, offs[d + 1])
            send[s][d] = (a, b) if b > a else (a, a)
    return q_total, pref, ranges, send


class _DevWords:
    """Device memory as a torch int64 tensor (through __cuda_array_interface__; no copy)."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i8", "data": (int(ptr), False), "version": 3}


class ShardOps:
    """A shard of the particle store behind the C ABI."""

    def __init__(self, state):
        self.state = state
        self.h = state._h
        self.state_dim = state.state_dim
        self.esz = 8 if state.dtype == L.F64 else 4

    def weight_max(self):
        v = C.c_double()
        L.check(L.lib().gb_pf_weight_max(self.h, C.byref(v)))
        return v.value

    def weight_sums(self, m_ref, bits):
        s1, s2, q = C.c_double(), C.c_double(), C.c_uint64()
        L.check(L.lib().gb_pf_weight_sums(self.h, m_ref, bits, C.byref(s1), C.byref(s2), C.byref(q)))
        return s1.value, s2.value, q.value

    # device-resident statistics record {m, s1, s2, lse, ess, q_total} for collectives without host round trips
    def stats_tensor(self, torch):
        p, n = C.c_void_p(), C.c_int()
        L.check(L.lib().gb_pf_stats_device(self.h, C.byref(p), C.byref(n)))
        return torch.as_tensor(_DevWords(p.value, n.value), device="cuda")

    def weight_max_device(self):
        L.check(L.lib().gb_pf_weight_max_device(self.h))

    def weight_sums_device(self, bits):
        L.check(L.lib().gb_pf_weight_sums_device(self.h, bits))

    def systematic_offspring(self, m_ref, bits, q_offset, q_local, q_total, u0):
        lo, hi = C.c_int64(), C.c_int64()
        L.check(L.lib().gb_pf_systematic_offspring(self.h, m_ref, bits, q_offset, q_local, q_total, u0, C.byref(lo), C.byref(hi)))
        return lo.value, hi.value

    def block_bytes(self, count):
        """== gb_pf_block_bytes: [parents int64 x count][rows S x count x esz], rounded up to 16 bytes."""
        return (int(count) * (8 + self.state_dim * self.esz) + 15) // 16 * 16

    def export_offspring(self, k_lo, k_hi, block):
        L.check(L.lib().gb_pf_export_offspring(self.h, k_lo, k_hi, C.c_void_p(block.data_ptr())))

    def import_rows(self, dst_lo, count, block):
        L.check(L.lib().gb_pf_import_rows(self.h, dst_lo, count, C.c_void_p(block.data_ptr())))

    def commit(self, log_ml_increment):
        L.check(L.lib().gb_pf_commit_resample(self.h, log_ml_increment))

    # fast path: local slots keep their ancestors, incoming rows are parked in the extension rows
    def ext_capacity(self):
        v = C.c_int64()
        L.check(L.lib().gb_pf_ext_capacity(self.h, C.byref(v)))
        return v.value

    def offspring_local(self, k_lo, k_hi):
        L.check(L.lib().gb_pf_offspring_local(self.h, k_lo, k_hi))

    def import_ext(self, dst_lo, count, ext_offset, block):
        L.check(L.lib().gb_pf_import_ext(self.h, dst_lo, count, ext_offset, C.c_void_p(block.data_ptr())))

    def commit_lazy(self, log_ml_increment):
        L.check(L.lib().gb_pf_commit_resample_lazy(self.h, log_ml_increment))

    def async_commit(self, log_ml_increment):
        L.check(L.lib().gb_pf_async_commit(self.h, log_ml_increment))

    def log_ml_est(self):
        return self.state.log_ml_est

    # asynchronous run (plan / decision / log-ML on the device, offspring by peer stores)
    def ipc_bundle(self):
        n = L.lib().gb_pf_ipc_bundle_bytes()
        buf = (C.c_char * n)()
        nb = C.c_int()
        L.check(L.lib().gb_pf_ipc_export(self.h, buf, C.byref(nb)))
        return bytes(buf[:nb.value])

    def attach_ipc(self, peer_rank, bundle):
        L.check(L.lib().gb_pf_async_attach_ipc(self.h, int(peer_rank), C.c_char_p(bundle)))

    def attach_local(self, peer_rank, peer_ops):
        L.check(L.lib().gb_pf_async_attach_local(self.h, int(peer_rank), peer_ops.h))

    def async_cur(self):
        v = C.c_int()
        L.check(L.lib().gb_pf_async_cur(self.h, C.byref(v)))
        return v.value

    def async_setup(self, world, rank, offs, ess_threshold):
        o = np.ascontiguousarray(offs, dtype=np.int64)
        L.check(L.lib().gb_pf_async_setup(self.h, world, rank, o.ctypes.data_as(L.i64p), float(ess_threshold)))

    def async_arm(self, ess_threshold):
        L.check(L.lib().gb_pf_async_arm(self.h, float(ess_threshold)))

    def xchg_post(self):
        L.check(L.lib().gb_pf_xchg_post(self.h))

    def xchg_stats(self):
        L.check(L.lib().gb_pf_xchg_stats(self.h))

    def async_plan(self, u0, ess_threshold, do_scan):
        L.check(L.lib().gb_pf_async_plan(self.h, int(u0), float(ess_threshold), int(do_scan)))

    def xchg_read(self, world):
        """Global (m, lse, ess) and every shard's integer CDF mass after gb_pf_xchg_stats + gb_pf_async_plan: one D2H read."""
        lse, ess, m = C.c_double(), C.c_double(), C.c_double()
        qs = (C.c_uint64 * world)()
        L.check(L.lib().gb_pf_xchg_read(self.h, C.byref(lse), C.byref(ess), None, None))
        L.check(L.lib().gb_pf_xchg_records(self.h, C.byref(m), qs))
        return m.value, lse.value, ess.value, [int(q) for q in qs]

    def async_wait(self):
        L.check(L.lib().gb_pf_async_wait(self.h))

    def async_iterate(self, obs):
        o = api._flatten_obs(obs)
        L.check(L.lib().gb_pf_async_iterate(self.h, o.ctypes.data_as(L.dp), o.size))

    def async_sweep(self):
        L.check(L.lib().gb_pf_async_sweep(self.h))

    def async_export(self):
        L.check(L.lib().gb_pf_async_export(self.h))

    def async_step(self, obs):
        o = api._flatten_obs(obs)
        L.check(L.lib().gb_pf_async_step(self.h, o.ctypes.data_as(L.dp), o.size))

    def async_finish(self):
        nres, ovf = C.c_int(), C.c_int()
        L.check(L.lib().gb_pf_async_finish(self.h, C.byref(nres), C.byref(ovf)))
        return nres.value


class ShardedParticleFilter:
    """Gen's particle-filter calls over a filter whose particles are sharded across the process group."""

    def __init__(self, model, n_global, dtype="f64", seed=0, group=None, ops_factory=None, device=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.n_global = int(n_global)
        self.counts, self.offs = shard_bounds(self.n_global, self.world)
        self.n_local = self.counts[self.rank]
        self.seed = int(seed)
        self.dtype = dtype
        self.model = model
        self.step = 0
        self.bits = int(L.lib().gb_quant_bits(self.n_global))
        if ops_factory is None:
            self.local = api.create_particle_filter(model, self.n_local, dtype=dtype, seed=seed, n_global=self.n_global,
                                                    pid_offset=self.offs[self.rank])
            self.ops = ShardOps(self.local)
            self.device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
            self.tdtype = torch.float64 if api._DTYPES[dtype] == L.F64 else torch.float32
        else:                                   # CPU test double (oracle-backed shard, gloo backend)
            self.local = None
            self.ops = ops_factory(self)
            self.device = torch.device("cpu")
            self.tdtype = torch.float64
        self.state_dim = self.ops.state_dim
        self.last_ess = None
        self._peers_attached = False
        self._bufs = {}
        # extension-row capacity of every shard (same formula on every rank: no collective needed)
        self.ext_caps = None
        if ops_factory is None and self.world > 1:
            self.ext_caps = [((c // 4 + 2047) // 2048 + 1) * 2048 for c in self.counts]
            assert self.ops.ext_capacity() == self.ext_caps[self.rank]
        self.force_staged = False
        self.use_p2p = True           # synchronous maybe_resample_: exchange by peer stores (False: all-to-all-v of blocks)
        # global weight statistics are cached until the weights change (log_ml_estimate followed by maybe_resample!
        # costs one statistics pass and one pair of collectives, as on a single GPU); every rank makes the same
        # sequence of calls, so all ranks hit or miss together
        self._wver = 0
        self._stats_cache = (-1, None)

    # ---- collectives (tiny) ----
    def _allreduce_max(self, x):
        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX, group=self.group)
        return float(t.item())

    def _allgather_triple(self, s1, s2, q):
        """(s1, s2, q) of every shard; q travels exactly as int64 (q < 2^62)."""
        loc = np.array([np.float64(s1).view(np.int64), np.float64(s2).view(np.int64), np.int64(q)], dtype=np.int64)
        if self.world == 1:
            allv = loc[None, :]
        else:
            t = self.torch.from_numpy(loc).to(self.device)
            out = self.torch.empty(self.world * 3, dtype=self.torch.int64, device=self.device)
            self.dist.all_gather_into_tensor(out, t, group=self.group)
            allv = out.cpu().numpy().reshape(self.world, 3)
        s1s = allv[:, 0].copy().view(np.float64)
        s2s = allv[:, 1].copy().view(np.float64)
        return s1s, s2s, [int(v) for v in allv[:, 2]]

    def _all_to_all_v(self, send, send_counts, recv_counts):
        torch, dist = self.torch, self.dist
        recv = self._buffer(int(sum(recv_counts)), "recv")[:int(sum(recv_counts))] if send.dtype == torch.uint8 \
            else torch.empty(int(sum(recv_counts)), dtype=send.dtype, device=send.device)
        if self.world == 1:
            recv.copy_(send)
            return recv
        if send.device.type == "cuda":
            dist.all_to_all_single(recv, send, list(map(int, recv_counts)), list(map(int, send_counts)), group=self.group)
            return recv
        # gloo / CPU: pairwise exchange
        so = np.concatenate([[0], np.cumsum(send_counts)]).astype(int)
        ro = np.concatenate([[0], np.cumsum(recv_counts)]).astype(int)
        recv[ro[self.rank]:ro[self.rank + 1]] = send[so[self.rank]:so[self.rank + 1]]
        reqs = []
        for peer in range(self.world):
            if peer == self.rank:
                continue
            if send_counts[peer]:
                reqs.append(dist.isend(send[so[peer]:so[peer + 1]].contiguous(), peer, group=self.group))
            if recv_counts[peer]:
                reqs.append(dist.irecv(recv[ro[peer]:ro[peer + 1]], peer, group=self.group))
        for r in reqs:
            r.wait()
        return recv

    # ---- Gen API ----
    def initialize(self, observations, proposal=None, proposal_args=()):
        api.generate_(self.local, observations, proposal, proposal_args)
        self.step = 0
        self._wver += 1
        return self

    def particle_filter_step_(self, new_args, observations, proposal=None, proposal_args=(), return_incr=False):
        out = api.particle_filter_step_(self.local, new_args, (api.UnknownChange,), observations, proposal, proposal_args,
                                        return_incr=return_incr)
        self.step += 1
        self._wver += 1
        return out

    def rejuvenate_(self, observations, move="resimulate", drift_std=0.0):
        """MH rejuvenation of every particle's latest latent choices (api.rejuvenate_): shard-local work, keyed by the
        global particle id, so the result does not depend on the GPU count.  Returns the global number of accepted moves."""
        acc = api.rejuvenate_(self.local, observations, move=move, drift_std=drift_std)
        if self.world == 1:
            return acc
        t = self.torch.tensor([acc], dtype=self.torch.int64, device=self.device)
        self.dist.all_reduce(t, group=self.group)
        return int(t.item())

    def _attach_peers(self):
        """Exchange CUDA IPC handles of the shards' buffers once (collective), so that offspring can be written straight
        into the owning shard's memory over NVLink."""
        if self._peers_attached:
            return
        if self.world == 1:
            self.ops.async_setup(1, 0, self.offs, math.inf)
            self._peers_attached = True
            return
        bundles = [None] * self.world
        self.dist.all_gather_object(bundles, (self.ops.ipc_bundle(), self.ops.async_cur()), group=self.group)
        assert len({b[1] for b in bundles}) == 1, "trace buffers of the shards are out of lock step"
        for r, (bundle, _) in enumerate(bundles):
            if r != self.rank:
                self.ops.attach_ipc(r, bundle)
        # exchange control block + mailbox reset, ONCE; nobody may post before every shard has reset its mailbox
        self.ops.async_setup(self.world, self.rank, self.offs, math.inf)
        self.torch.cuda.synchronize()
        self.dist.barrier(group=self.group)
        self._peers_attached = True

    def run_(self, observations_seq, ess_threshold=None):
        """T x { maybe_resample!; particle_filter_step! } over the sharded filter WITHOUT any host round trip and without
        collectives: the shards' maxima, sums and "rows delivered" flags travel through peer-memory mailboxes written and
        awaited by the kernels themselves, the exchange plan, the ESS test and the log-ML update run in a device kernel,
        and offspring owned by another shard are written straight into that shard's memory over NVLink (CUDA IPC peer
        pointers) -- include/genb200.h "Sharded maybe_resample! WITHOUT collectives".  The host only enqueues (one C call
        per step).  Same results as the synchronous path bit for bit, as long as no shard receives more rows than its
        extension rows hold (25 % of the shard; otherwise GenB200Error and the synchronous staged path must be used).
        Returns the number of resamples."""
        thr = self.n_global / 2 if ess_threshold is None else float(ess_threshold)
        self._attach_peers()
        self.ops.async_arm(thr)
        for obs in observations_seq:
            self.ops.async_iterate(obs)
            self.step += 1
        nres = self.ops.async_finish()
        self._wver += 1
        if self.local is not None and self.local._last_t is not None:
            self.local._last_t += len(observations_seq)
        return nres

    def _global_stats(self):
        """Global (m, lse, ess) and the gathered integer masses (particle_filter.jl:3-12 over all shards)."""
        if self._stats_cache[0] == self._wver:
            return self._stats_cache[1]
        out = self._global_stats_uncached()
        self._stats_cache = (self._wver, out)
        return out

    def _global_stats_uncached(self):
        if self.device.type == "cuda" and self.use_p2p:
            # peer-memory exchange (no collective): statistics pass w.r.t. the global max, every shard's record in every
            # mailbox, the global record assembled by a device kernel, ONE device->host read
            self._attach_peers()
            self.ops.xchg_stats()
            self.ops.async_plan(0, -math.inf, 0)
            return self.ops.xchg_read(self.world)
        else:
            m = self._allreduce_max(self.ops.weight_max())
            s1, s2, q = self.ops.weight_sums(m, self.bits)
            s1s, s2s, qs = self._allgather_triple(s1, s2, q)
        S1, S2 = float(np.sum(s1s)), float(np.sum(s2s))      # fixed rank order: identical on every rank
        lse = -math.inf if m == -math.inf else m + math.log(S1)
        ess = (S1 * S1) / S2 if S2 > 0 else float("nan")
        return m, lse, ess, qs

    def log_ml_estimate(self):
        """particle_filter.jl:91-94 with the global log-sum-exp."""
        _, lse, _, _ = self._global_stats()
        return self.ops.log_ml_est() + lse - math.log(self.n_global)

    def maybe_resample_(self, ess_threshold=None, verbose=False):
        """maybe_resample! (particle_filter.jl:249-275) with global-prefix systematic resampling."""
        thr = self.n_global / 2 if ess_threshold is None else float(ess_threshold)
        m, lse, ess, qs = self._global_stats()
        self.last_ess = ess
        if verbose and self.rank == 0:
            print("effective sample size: %s, doing resample: %s" % (ess, ess < thr))
        if not (ess < thr):                                  # :256, strict; NaN compares false
            return False
        u0 = api.philox_resample_u0(self.seed, self.step)
        me, W = self.rank, self.world
        if sum(int(q) for q in qs) == 0:                     # no representable weight mass: nothing to resample from
            return False
        self._wver += 1                                      # the weights are about to become 0 (:266)
        q_total, pref, ranges, send = plan_exchange(qs, self.n_global, self.offs, u0)
        # Fast path when every shard can park its incoming rows in its extension rows (all ranks evaluate
        # the same predicate from the gathered plan, so no extra collective is needed to agree on it)
        fast = self.ext_caps is not None and not self.force_staged and all(
            sum(send[s][d][1] - send[s][d][0] for s in range(W) if s != d) <= self.ext_caps[d] for d in range(W))
        if self.device.type == "cuda" and self.use_p2p and (fast or W == 1):
            # The decision was taken here (the Bool is returned); the exchange itself is only enqueued: plan on the device
            # from the records in the mailbox, ancestors, offspring written into the owning shards' memory by peer stores
            # over NVLink, flags as the barrier -- no staging blocks, no all-to-all, no collective
            self.ops.async_plan(u0, math.inf, 1)
            self.ops.async_sweep()
            self.ops.async_export()
            self.ops.async_wait()
            self.ops.async_commit(lse - math.log(self.n_global))   # :263
            return True
        lo, hi = self.ops.systematic_offspring(m, self.bits, pref[me], qs[me], q_total, u0)
        assert (lo, hi) == ranges[me], "offspring range mismatch between host plan and device sweep"
        send_cnt = [send[me][d][1] - send[me][d][0] for d in range(W)]
        recv_cnt = [send[s][me][1] - send[s][me][0] for s in range(W)]
        if fast:
            a, b = send[me][me]
            self.ops.offspring_local(a, b)                   # stays here: ancestors only, gather deferred
            send_cnt[me] = 0
            recv_cnt[me] = 0
        sbytes = [self.ops.block_bytes(c) for c in send_cnt]
        rbytes = [self.ops.block_bytes(c) for c in recv_cnt]
        any_traffic = any(send[s][d][1] > send[s][d][0] for s in range(W) for d in range(W) if not (fast and s == d))
        if any_traffic:
            sbuf = self._buffer(sum(sbytes), "send")
            off = 0
            for d in range(W):                               # per destination one block, in rank order
                if send_cnt[d]:
                    a, b = send[me][d]
                    self.ops.export_offspring(a, b, sbuf[off:off + sbytes[d]])
                off += sbytes[d]
            rbuf = self._all_to_all_v(sbuf[:sum(sbytes)], sbytes, rbytes)
            off = ext = 0
            for s_ in range(W):
                if recv_cnt[s_]:
                    a, b = send[s_][me]
                    blk = rbuf[off:off + rbytes[s_]]
                    if fast:
                        self.ops.import_ext(a - self.offs[me], recv_cnt[s_], ext, blk)
                        ext += recv_cnt[s_]
                    else:
                        self.ops.import_rows(a - self.offs[me], recv_cnt[s_], blk)
                off += rbytes[s_]
        if fast:
            self.ops.commit_lazy(lse - math.log(self.n_global))    # :263
        else:
            assert sum(recv_cnt) == self.n_local, "resampled shard is not full"
            self.ops.commit(lse - math.log(self.n_global))         # :263
        return True

    def _buffer(self, nbytes, tag):
        """Reusable byte buffer for the exchange (grown geometrically) instead of one allocation per resample."""
        cur = self._bufs.get(tag)
        if cur is None or cur.numel() < nbytes:
            cur = self.torch.empty(max(int(nbytes * 1.5), 4096), dtype=self.torch.uint8, device=self.device)
            self._bufs[tag] = cur
        return cur

    # ---- accessors (gather to every rank; for tests / small filters) ----
    def _gather_array(self, arr, np_dtype):
        torch, dist = self.torch, self.dist
        if self.world == 1:
            return arr
        cmax = max(self.counts)                      # equal-sized pieces for every backend
        pad = np.zeros(cmax, dtype=arr.dtype)
        pad[:len(arr)] = arr
        t = torch.from_numpy(pad).to(self.device)
        outs = [torch.empty(cmax, dtype=t.dtype, device=self.device) for _ in self.counts]
        dist.all_gather(outs, t, group=self.group)
        return np.concatenate([o.cpu().numpy()[:c] for o, c in zip(outs, self.counts)]).astype(np_dtype)

    def get_log_weights(self):
        return self._gather_array(api.get_log_weights(self.local), np.float64)

    def get_state(self, field):
        return self._gather_array(self.local.get_state(field), np.float64)

    def get_parents(self):
        return self._gather_array(self.local.parents, np.int64)
