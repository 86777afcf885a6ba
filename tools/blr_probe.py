"""Config 4 (Bayesian linear regression importance sampling, D = 64, n = 64): where the time goes."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as entry, fixtures as F
g = entry.load_package()
import torch
X, y, sigma = F.blr_problem()
model = g.blr_model(X, sigma)
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10 ** 7
for dtype in ("f64", "f32"):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    st = g.create_particle_filter(model, n, dtype=dtype, seed=1)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    st.enable_timing(True)
    g.generate_(st, y)
    st.sync(); t2 = time.perf_counter()
    lml = g.log_ml_estimate(st)
    t3 = time.perf_counter()
    lw = st.log_weights
    t4 = time.perf_counter()
    tm = st.get_timing()
    print(dtype, "alloc %.1f ms, generate %.1f ms, log_ml_estimate %.1f ms, D2H log weights %.1f ms; kernels:" % (1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2), 1e3 * (t4 - t3)),
          {k: round(v[0], 3) for k, v in tm.items() if v[1]}, "lml", lml)
    st.enable_timing(False)
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        g.generate_(st, y); st.sync()
        print("   generate again: %.2f ms" % (1e3 * (time.perf_counter() - t0)))
    del st
