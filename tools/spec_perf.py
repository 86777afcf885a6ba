"""Step-kernel time of the bearings model: specialised kind vs the same model as a GB_MODEL_SPEC program."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import __graft_entry__ as entry, fixtures as F
g = entry.load_package()
import test_gpu_parity as T
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
zs = F.bearings_series(12)
for label, model, jit in (("specialised kind (ahead of time)", g.bearings_model(), "1"), ("spec program, NVRTC-specialised", T.bearings_spec(g), "1"),
                          ("spec program, interpreter", T.bearings_spec(g), "0")):
    os.environ["GENB200_SPEC_JIT"] = jit
    st = g.initialize_particle_filter(model, (0,), zs[0], n, seed=1)
    for t in range(1, 4):
        g.maybe_resample_(st, ess_threshold=n + 1, resampler="systematic")
        g.particle_filter_step_(st, (t,), (g.UnknownChange,), zs[t], return_incr=False)
    st.enable_timing(True); st.get_timing()
    for t in range(4, 10):
        g.maybe_resample_(st, ess_threshold=n + 1, resampler="systematic")
        g.particle_filter_step_(st, (t,), (g.UnknownChange,), zs[t], return_incr=False)
    tm = st.get_timing()
    print(label, {k: round(v[0] / max(1, v[1]), 4) for k, v in tm.items() if v[1]}, "log ML", g.log_ml_estimate(st), "backend", st.spec_backend)
    del st
