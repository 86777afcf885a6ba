"""BASELINE configs 4 and 5 over several GPUs through the single-process gb_spf_* entries (no NCCL): wall-clock per call with
the state resident on the devices.  Run on a multi-GPU box: python tools/multigpu_configs.py [max_gpus]
  cfg 4  importance_sampling of the Bayesian linear regression (D = 64, n = 64), 1e7 samples sharded over G GPUs
         (generate is shard-local, the log-sum-exp is one peer-memory exchange)
  cfg 5  maybe_resample! (global-prefix systematic) of a filter holding 1e8 log weights sharded over G GPUs"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as entry  # noqa: E402
import fixtures as F  # noqa: E402

g = entry.load_package()
from genb200 import multigpu as M  # noqa: E402

gmax = int(sys.argv[1]) if len(sys.argv) > 1 else g.device_count()
out = []
X, y, sigma = F.blr_problem(D=64, n=64)
blr = g.blr_model(X, sigma)
G = 1
while G <= min(gmax, g.device_count()):
    devs = list(range(G))
    # ---- cfg 4 ----
    n = 10 ** 7
    for rep in range(3):
        t0 = time.perf_counter()
        st, _, lml = M.importance_sampling(blr, (), y, n, devices=devs, seed=3, return_weights=False)
        st.sync()
        dt = time.perf_counter() - t0
        del st
    out.append({"config": "configs[3] BLR importance_sampling 1e7 samples fp64", "gpus": G, "ms": 1e3 * dt, "samples_per_s": n / dt, "log_ml": lml})
    # ---- cfg 5 ----
    n = 10 ** 8
    ys = F.lgssm_series(12)
    st = M.initialize_particle_filter(g.lgssm_model(**F.LGSSM), (0,), ys[0], n, devices=devs, seed=1)
    ts = []
    for t in range(1, 9):
        st.sync()
        t0 = time.perf_counter()
        did = M.maybe_resample_(st, ess_threshold=n + 1)
        st.sync()
        ts.append(time.perf_counter() - t0)
        assert did
        M.particle_filter_step_(st, (t,), (g.UnknownChange,), ys[t])
    dt = sorted(ts[2:])[len(ts[2:]) // 2]
    out.append({"config": "configs[4] systematic maybe_resample! of 1e8 fp64 log weights (statistics + exchange + sweep + export; "
                          "the gather is deferred into the next step)", "gpus": G, "ms": 1e3 * dt, "weights_per_s": n / dt})
    del st
    g.cache_trim()
    G *= 2
for o in out:
    print(json.dumps(o))
