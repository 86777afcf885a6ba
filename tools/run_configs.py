"""Measure the BASELINE.json parity configs (1, 2, 4, 5) on one B200; prints a markdown table.
Run on the GPU box: python tools/run_configs.py > gpurun_out/configs.md"""
import math
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as entry  # noqa: E402
import fixtures as F  # noqa: E402

g = entry.load_package()
import torch  # noqa: E402  (events only)

PEAK = 6549.4
rows = []


def timed(fn, reps=1):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        r = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps, r


def pf_run(model, ys, n, dtype, resampler="systematic", thr=None, seed=0):
    st = g.initialize_particle_filter(model, (0,), ys[0], n, dtype=dtype, seed=seed)
    for t in range(1, len(ys)):
        g.maybe_resample_(st, ess_threshold=thr, resampler=resampler)
        g.particle_filter_step_(st, (t,), (g.UnknownChange,), ys[t], return_incr=False)
    return g.log_ml_estimate(st)


# ---- cfg 1: LGSSM N=1000, T=100, 32 seeds vs Kalman ----
ys = F.lgssm_series(101)
O = entry.load_oracle()
kf = O.kalman_logml(0.0, 1.0, 0.9, 0.5, 1.0, ys)
model = g.lgssm_model(**F.LGSSM)
pf_run(model, ys, 1000, "f64")
t0 = time.perf_counter()
ests = [pf_run(model, ys, 1000, "f64", seed=s) for s in range(32)]
dt = (time.perf_counter() - t0) / 32
rows.append(("1: LGSSM N=1e3 T=100 fp64, systematic @ ESS<N/2", "%.3g" % (1000 * 100 / dt), "%.1f us/step" % (1e6 * dt / 100),
             "mean log-ML error vs Kalman %.4f (sd %.3f over 32 seeds)" % (np.mean(ests) - kf, np.std(ests))))

def pf_run_fused(model, ys, n, dtype, thr, seed):
    st = g.initialize_particle_filter(model, (0,), ys[0], n, dtype=dtype, seed=seed)
    g.particle_filter_run_(st, ys[1:], ess_threshold=thr)
    return g.log_ml_estimate(st)


pf_run_fused(model, ys, 1000, "f64", 500, 0)
t0 = time.perf_counter()
ests = [pf_run_fused(model, ys, 1000, "f64", 500, s) for s in range(32)]
dt = (time.perf_counter() - t0) / 32
rows.append(("1: same, fused gb_pf_run (N <= 4096: the whole loop in one single-CTA kernel)", "%.3g" % (1000 * 100 / dt),
             "%.1f us/step" % (1e6 * dt / 100), "mean log-ML error vs Kalman %.4f (sd %.3f over 32 seeds); includes filter "
             "creation and the log-ML read per seed" % (np.mean(ests) - kf, np.std(ests))))

# ---- cfg 2: SV N=1e6, T=1000 ----
ys = F.sv_series(1001)
model = g.sv_model(**F.SV)
res = {}
for dtype, E in (("f64", 8), ("f32", 4)):
    pf_run(model, ys[:20], 10 ** 6, dtype)
    dt, lml = timed(lambda: pf_run(model, ys, 10 ** 6, dtype, thr=10 ** 6 + 1, seed=3))
    res[dtype] = lml
    bstep = 5 * 1 * E * 0 + (4 * 1 + 5) * E + 8
    rows.append(("2: SV N=1e6 T=1000 %s, systematic every step" % dtype, "%.3g" % (1e6 * 1000 / dt), "%.1f us/step" % (1e6 * dt / 1000),
                 "log-ML %.6f; %.0f B/particle-step -> %.2f of measured HBM peak (latency regime)" % (lml, bstep, bstep * 1e6 / (dt / 1000) / 1e9 / PEAK)))
for dtype, E in (("f64", 8), ("f32", 4)):
    def fused():
        st = g.initialize_particle_filter(model, (0,), ys[0], 10 ** 6, dtype=dtype, seed=3)
        g.particle_filter_run_(st, ys[1:], ess_threshold=10 ** 6 + 1)
        return g.log_ml_estimate(st)
    fused()
    dt, lml = timed(fused)
    bstep = (4 * 1 + 5) * E + 8
    rows.append(("2: SV N=1e6 T=1000 %s, fused gb_pf_run (no host round trip per step)" % dtype, "%.3g" % (1e6 * 1000 / dt),
                 "%.1f us/step" % (1e6 * dt / 1000), "log-ML %.6f (== API loop: %s); %.2f of measured HBM peak" % (lml, lml == res[dtype], bstep * 1e6 / (dt / 1000) / 1e9 / PEAK)))
rows.append(("2: fp32 vs fp64 log-ML", "", "", "relative difference %.2e (different noise streams: statistical)" % abs((res["f32"] - res["f64"]) / res["f64"])))

# ---- cfg 4: BLR importance sampling D=64, 1e7 samples ----
X, y, sigma = F.blr_problem()
model = g.blr_model(X, sigma)
for dtype, E in (("f64", 8), ("f32", 4)):
    g.importance_sampling(model, (X,), y, 10 ** 5, dtype=dtype, return_traces=False)
    dt, out = timed(lambda: g.importance_sampling(model, (X,), y, 10 ** 7, dtype=dtype, seed=1, return_traces=False))
    rows.append(("4: BLR importance sampling D=64 n=64, 1e7 samples %s (incl. alloc + 80 MB D2H of log weights)" % dtype,
                 "%.3g samples/s" % (1e7 / dt), "%.1f ms" % (1e3 * dt),
                 "lml estimate %.3f; %d B/sample -> %.2f of HBM peak; %.1f GFLOP/s likelihood" % (out[2], 65 * E, 65 * E * 1e7 / dt / 1e9 / PEAK, 8192e7 / dt / 1e9)))

# ---- cfg 5: resampling microbenchmark (device-resident weights) ----
from genb200 import _lib as L  # noqa: E402
import ctypes as C  # noqa: E402
for n in (10 ** 4, 10 ** 6, 10 ** 8, 10 ** 9):
    for shape in ("normal", "equal", "onehot"):
        if n == 10 ** 9 and shape != "normal":
            continue
        ld = (n + 2047) // 2048 * 2048
        lw = torch.zeros(ld, dtype=torch.float64, device="cuda")
        if shape == "normal":
            gen = torch.Generator(device="cuda").manual_seed(20244)
            lw[:n] = 3.0 * torch.randn(n, dtype=torch.float64, device="cuda", generator=gen)
        elif shape == "onehot":
            lw[:n] = -50.0
            lw[n // 3] = 0.0
        anc = torch.empty(ld, dtype=torch.int32, device="cuda")
        for kind, kname in ((1, "systematic"), (2, "residual"), (0, "multinomial")):
            if n >= 10 ** 9 and kind != 1:
                continue
            if n >= 10 ** 8 and kind == 0 and shape != "normal":
                continue

            def call():
                L.check(L.lib().gb_resample_device(C.c_void_p(lw.data_ptr()), n, L.F64, kind, 12345, None, 7, 1,
                                                   C.c_void_p(anc.data_ptr()), None))
            call()
            reps = 20 if n <= 10 ** 6 else 3
            dt, _ = timed(call, reps)
            a = anc[:n]
            ok = bool((a.min() >= 0) and (a.max() < n))
            if kind != 0:
                ok = ok and bool((a[1:] >= a[:-1]).all())
            rows.append(("5: resample %s N=%.0e %s fp64" % (kname, n, shape), "%.3g weights/s" % (n / dt), "%.3f ms" % (1e3 * dt),
                         "12 B/weight -> %.2f of measured HBM peak; valid=%s" % (12 * n / dt / 1e9 / PEAK, ok)))
        del lw, anc
        torch.cuda.empty_cache()

print("| config | throughput | time | notes |\n|---|---|---|---|")
for r in rows:
    print("| " + " | ".join(r) + " |")
