import sys, os, math, numpy as np
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import __graft_entry__ as entry, fixtures as F
g = entry.load_package(); O = entry.load_oracle()
for name, model, okind, params, ys in (("lgssm", g.lgssm_model(**F.LGSSM), O.MODEL_LGSSM, F.lgssm_params(), F.lgssm_series(12)),
                                       ("sv", g.sv_model(**F.SV), O.MODEL_SV, F.sv_params(), F.sv_series(12)),
                                       ("bearings", g.bearings_model(), O.MODEL_BEARINGS, F.BEARINGS_PARAMS, F.bearings_series(12))):
    n = 200000
    rng = np.random.default_rng(3)
    nn0, nn1 = model.num_normals(True), model.num_normals(False)
    st = g.create_particle_filter(model, n, dtype="f32")
    eps = rng.standard_normal((nn0, n)).astype(np.float32).astype(np.float64)
    st.set_noise(eps); g.generate_(st, ys[0])
    ref = O.OraclePF.init(okind, params, n, [ys[0]], normals=eps)
    worst_lml = worst_w = 0.0
    for t in range(1, 10):
        ref.log_weights = st.log_weights
        st.set_resample_noise(u0=1234 + t)
        g.maybe_resample_(st, ess_threshold=n + 1, resampler="systematic")
        ref.maybe_resample(ess_threshold=n + 1, kind=O.RESAMPLE_SYSTEMATIC, u0=1234 + t)
        eps = rng.standard_normal((nn1, n)).astype(np.float32).astype(np.float64)
        st.set_noise(eps)
        incr, = g.particle_filter_step_(st, (t,), (g.UnknownChange,), ys[t])
        incr_ref = ref.step([ys[t]], normals=eps)
        lml, lml_ref = g.log_ml_estimate(st), ref.log_ml_estimate()
        worst_lml = max(worst_lml, abs(lml - lml_ref) / max(1.0, abs(lml_ref)))
        worst_w = max(worst_w, np.max(np.abs(incr - incr_ref) / np.maximum(1.0, np.abs(incr_ref))))
        for f in range(model.state_dim):
            np.ctypeslib.as_array(O.lib().go_pf_state(ref._p, f), shape=(n,))[:] = st.get_state(f)
    print("%-9s fp32: worst rel log-ML error %.2e, worst rel per-particle incr error %.2e" % (name, worst_lml, worst_w))
