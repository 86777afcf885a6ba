"""Per-kernel device time of one SV particle-filter step at small N (the launch-latency regime of config 2).
Run on the GPU box: python tools/small_n_probe.py [N]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as entry  # noqa: E402
import fixtures as F  # noqa: E402

g = entry.load_package()
import torch  # noqa: E402

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10 ** 6
ys = F.sv_series(401)
model = g.sv_model(**F.SV)
for dtype in ("f64", "f32"):
    for mode in ("api", "fused"):
        st = g.initialize_particle_filter(model, (0,), ys[0], n, dtype=dtype, seed=3)
        g.particle_filter_run_(st, ys[1:20], ess_threshold=n + 1)
        st.enable_timing(True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        T = 300
        if mode == "fused":
            g.particle_filter_run_(st, ys[20:20 + T], ess_threshold=n + 1)
        else:
            for t in range(20, 20 + T):
                g.maybe_resample_(st, ess_threshold=n + 1, resampler="systematic")
                g.particle_filter_step_(st, (t,), (g.UnknownChange,), ys[t], return_incr=False)
        g.log_ml_estimate(st)
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) / T
        tm = st.get_timing()
        tot = sum(v[0] for v in tm.values()) / T
        print("N=%d %s %s: wall %.1f us/step (with per-kernel events), kernels %.1f us/step" % (n, dtype, mode, wall * 1e6, tot * 1e3))
        for k, (ms, ln) in tm.items():
            if ln:
                print("    %-36s %6d launches  %.2f us each  %.1f us/step" % (k, ln, 1e3 * ms / ln, 1e3 * ms / T))
        st.enable_timing(False)
        st2 = g.initialize_particle_filter(model, (0,), ys[0], n, dtype=dtype, seed=3)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        if mode == "fused":
            g.particle_filter_run_(st2, ys[1:1 + T], ess_threshold=n + 1)
        else:
            for t in range(1, 1 + T):
                g.maybe_resample_(st2, ess_threshold=n + 1, resampler="systematic")
                g.particle_filter_step_(st2, (t,), (g.UnknownChange,), ys[t], return_incr=False)
        g.log_ml_estimate(st2)
        torch.cuda.synchronize()
        print("    untimed wall: %.1f us/step" % ((time.perf_counter() - t0) / T * 1e6))
