"""One process per GPU (needs N >= 2 GPUs; spawned by tests/test_gpu_multigpu.py::test_multiprocess_ipc_sharded_filter):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/run_sharded_multiproc.py
Every rank drives its shard through ShardedParticleFilter -- peer-memory exchange over CUDA IPC pointers (no collective
in the step), the all-to-all-v fallback and the staged fallback; rank 0 also runs the same filter on one GPU through the
ordinary API and the results must be bit-identical.  torch.distributed only carries the IPC handles, the result gathers
and (fallback paths) the block exchange."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as entry  # noqa: E402
import fixtures as F  # noqa: E402


def main():
    lr = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    g = entry.load_package()
    from genb200 import sharded, _lib as L
    L.check(L.lib().gb_set_device(lr))
    rank, world = dist.get_rank(), dist.get_world_size()
    # synchronous path: exchange by peer stores (default), by all-to-all-v of blocks, by the staged fallback
    for n, staged, p2p in ((100003, False, True), (100003, False, False), (100003, True, False), (4_000_000, False, True)):
        ys = F.bearings_series(8)
        seed = 17
        spf = sharded.ShardedParticleFilter(g.bearings_model(), n, seed=seed).initialize(ys[0])
        spf.force_staged = staged
        spf.use_p2p = p2p
        nres = nacc = nacc1 = 0
        for t in range(1, len(ys)):
            nres += spf.maybe_resample_(ess_threshold=(n + 1 if t % 2 else n / 2))
            spf.particle_filter_step_((t,), ys[t])
            nacc += spf.rejuvenate_(ys[t], move="drift" if t % 2 else "resimulate", drift_std=1e-3)   # shard-local MH moves
        cols = [spf.get_state(f) for f in range(4)]
        par = spf.get_parents()
        lml = spf.log_ml_estimate()
        if rank == 0:
            st = g.initialize_particle_filter(g.bearings_model(), (0,), ys[0], n, seed=seed)
            for t in range(1, len(ys)):
                g.maybe_resample_(st, ess_threshold=(n + 1 if t % 2 else n / 2), resampler="systematic")
                g.particle_filter_step_(st, (t,), (g.UnknownChange,), ys[t])
                nacc1 += g.rejuvenate_(st, ys[t], move="drift" if t % 2 else "resimulate", drift_std=1e-3)
            ok = nacc == nacc1 and nacc > 0 and all(np.array_equal(cols[f], st.get_state(f)) for f in range(4)) and np.array_equal(par, st.parents)
            ok = ok and abs(lml - g.log_ml_estimate(st)) <= 1e-10 * abs(lml)
            print("multi-process sharded check world=%d n=%d staged=%s p2p=%s resamples=%d mh-accepts=%d: %s (log ML %.12f)" % (world, n, staged, p2p, nres, nacc, "OK" if ok else "MISMATCH", lml))
            if not ok:
                sys.exit(1)
    # asynchronous run (device-side plan and exchange, nothing but enqueues on the host) vs the fused single-GPU run
    for n in (100003, 4_000_000):
        ys = F.bearings_series(10)
        spf = sharded.ShardedParticleFilter(g.bearings_model(), n, seed=9).initialize(ys[0])
        nres = spf.run_(ys[1:], ess_threshold=n / 2)
        cols = [spf.get_state(f) for f in range(4)]
        par = spf.get_parents()
        lml = spf.log_ml_estimate()
        if rank == 0:
            st = g.initialize_particle_filter(g.bearings_model(), (0,), ys[0], n, seed=9)
            nres1 = g.particle_filter_run_(st, ys[1:], ess_threshold=n / 2)
            ok = nres == nres1 and all(np.array_equal(cols[f], st.get_state(f)) for f in range(4)) and np.array_equal(par, st.parents)
            ok = ok and abs(lml - g.log_ml_estimate(st)) <= 1e-10 * abs(lml)
            print("multi-process async sharded run world=%d n=%d resamples=%d: %s (log ML %.12f)" % (world, n, nres, "OK" if ok else "MISMATCH", lml))
            if not ok:
                sys.exit(1)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
