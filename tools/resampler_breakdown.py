"""Per-kernel time of maybe_resample! for the resamplers at N = 1e8 (LGSSM filter, weights after one step)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as entry, fixtures as F
g = entry.load_package()
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10 ** 8
ys = F.lgssm_series(8)
for kind in ("systematic", "residual", "multinomial", "multinomial_sorted"):
    st = g.initialize_particle_filter(g.lgssm_model(**F.LGSSM), (0,), ys[0], n, seed=1)
    for t in (1, 2):
        g.maybe_resample_(st, ess_threshold=n + 1, resampler=kind)
        g.particle_filter_step_(st, (t,), (g.UnknownChange,), ys[t], return_incr=False)
    st.enable_timing(True); st.get_timing()
    for t in (3, 4, 5):
        g.maybe_resample_(st, ess_threshold=n + 1, resampler=kind)
        g.particle_filter_step_(st, (t,), (g.UnknownChange,), ys[t], return_incr=False)
    tm = st.get_timing()
    print(kind, {k: (round(v[0] / 3, 3), v[1] // 3) for k, v in tm.items() if v[1]})
    del st
