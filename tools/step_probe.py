"""Device time of the step kernel (deferred gather) for each model kind and dtype at one N (A/B of step-kernel variants).
Run on the GPU box: [GENB200_LIB=...] python tools/step_probe.py [N]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as entry  # noqa: E402
import fixtures as F  # noqa: E402

g = entry.load_package()
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 2 * 10 ** 7
zoo = {"sv": (g.sv_model(**F.SV), F.sv_series(40)), "lgssm": (g.lgssm_model(**F.LGSSM), F.lgssm_series(40)),
       "bearings": (g.bearings_model(), F.bearings_series(40))}
for name, (model, ys) in zoo.items():
    for dtype in ("f64", "f32"):
        st = g.initialize_particle_filter(model, (0,), ys[0], n, dtype=dtype, seed=3)
        for t in range(1, 4):
            g.maybe_resample_(st, ess_threshold=n + 1, resampler="systematic")
            g.particle_filter_step_(st, (t,), (g.UnknownChange,), ys[t], return_incr=False)
        st.enable_timing(True)
        st.get_timing()
        T = 12
        for t in range(4, 4 + T):
            g.maybe_resample_(st, ess_threshold=n + 1, resampler="systematic")
            g.particle_filter_step_(st, (t,), (g.UnknownChange,), ys[t], return_incr=False)
        st.sync()
        tm = st.get_timing()
        print("N=%d %-8s %s: " % (n, name, dtype) + ", ".join("%s %.4f ms" % (k.split("(")[0], ms / ln) for k, (ms, ln) in tm.items() if ln))
        # the device-controlled run (gather decided on the device: the GATHER == 2 instantiation of the step kernel)
        st.get_timing()
        g.particle_filter_run_(st, ys[16:16 + T], ess_threshold=n + 1)
        st.sync()
        tm = st.get_timing()
        print("   run_: " + ", ".join("%s %.4f ms" % (k.split("(")[0], ms / ln) for k, (ms, ln) in tm.items() if ln))
        del st
