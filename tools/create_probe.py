"""Cost of creating / initialising / freeing a small filter (config 1 creates one per seed)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as entry, fixtures as F
g = entry.load_package()
ys = F.lgssm_series(101)
model = g.lgssm_model(**F.LGSSM)
st = g.initialize_particle_filter(model, (0,), ys[0], 1000, seed=0); del st
for n in (1000, 10 ** 6):
    t0 = time.perf_counter()
    for s in range(20):
        st = g.create_particle_filter(model, n, seed=s)
        del st
    t1 = time.perf_counter()
    for s in range(20):
        st = g.initialize_particle_filter(model, (0,), ys[0], n, seed=s)
        st.sync()
        del st
    t2 = time.perf_counter()
    st = g.initialize_particle_filter(model, (0,), ys[0], n, seed=1)
    t3 = time.perf_counter()
    for s in range(20):
        g.particle_filter_run_(st, ys[1:], ess_threshold=n / 2)
        g.log_ml_estimate(st)
    t4 = time.perf_counter()
    print("n=%d: create+free %.2f ms, initialize+free %.2f ms, run 100 steps + log_ml %.2f ms" % (n, 50 * (t1 - t0), 50 * (t2 - t1), 50 * (t4 - t3)))
