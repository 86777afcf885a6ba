#!/usr/bin/env python
"""bench.py -- particle-steps/sec of the SMC hot path (BASELINE.json metric) on N B200s.

Workload (config.workload): BASELINE.json configs[2], the 2-D bearings-only tracker
(docs/src/tutorials/smc.md:607-661; 4-D state, nonlinear observation) bootstrap particle filter with
N = 10^8 particles (fp64) and systematic resampling EVERY step.  One "step" = one particle-filter
step over all particles: maybe_resample! (ESS, ancestors, resampled-trace gather) followed by
particle_filter_step! (propose + weight + log-sum-exp partials).  At --gpus G > 1 the 10^8 particles are
sharded G ways (strong scaling, as BASELINE.json words the config; one process per GPU, global-prefix
systematic resampling, every exchange through peer memory inside the kernels -- no collective in the
step); a weak-scaling run (10^8 per GPU) rides in the same JSON line under "weak".

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl genb200|reference]
                    [--particles P] [--dtype f64|f32] [--scaling strong|weak]

Prints ONE JSON line on rank 0.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as entry  # noqa: E402
import fixtures as F  # noqa: E402

S_BEARINGS = 4


def algorithmic_bytes_per_particle_step(S, E):
    """B_step = (4S+5) E + 8 (int32 ancestors, resample every step, history off)."""
    return (4 * S + 5) * E + 8


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.windows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "10"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def wait_ready(self, timeout=3.0):
        """nvidia-smi takes 0.1-0.5 s to print its first row: the timed region starts only once it is sampling."""
        t_end = time.perf_counter() + timeout
        while self.proc and not self.rows and time.perf_counter() < t_end:
            time.sleep(0.005)

    def window(self, t0, t1):
        """A timed region (host clock): only samples taken inside a region count."""
        self.windows.append((t0, t1))

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, r in self.rows:
            if self.windows and not any(a <= ts <= b for a, b in self.windows):
                continue
            f = [x.strip() for x in r.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm),
                "window": "nvidia-smi every 10 ms; samples inside the device-timed region"}


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle (a C restatement of Gen's particle filter; Gen.jl itself cannot run here: no Julia)
# ------------------------------------------------------------------------------------------------
def cpu_pf_rate(n, steps, threads, zs):
    """particle-steps/s of the CPU oracle on a bounded sample of the bearings workload."""
    O = entry.load_oracle()
    O.lib().go_set_threads(threads)
    pf = O.OraclePF.init(O.MODEL_BEARINGS, F.BEARINGS_PARAMS, n, [zs[0]], seed=1)
    lib = O.lib()
    import ctypes as C
    obs = (C.c_double * 1)()

    def one(t):
        lib.go_pf_maybe_resample(pf._p, float(n + 1), O.RESAMPLE_SYSTEMATIC, 0, 0, None, None)
        obs[0] = zs[t % len(zs)]
        lib.go_pf_step(pf._p, obs, 0, None, None, None, None)

    one(1)
    t0 = time.perf_counter()
    for t in range(2, 2 + steps):
        one(t)
    dt = time.perf_counter() - t0
    return n * steps / dt, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    zs = F.bearings_series(64)
    O = entry.load_oracle()
    threads = O.lib().go_get_max_threads()
    n = args.ref_particles
    O.lib().go_set_threads(threads)
    pf = O.OraclePF.init(O.MODEL_BEARINGS, F.BEARINGS_PARAMS, n, [zs[0]], seed=1)
    import ctypes as C
    lib = O.lib()
    obs = (C.c_double * 1)()

    def one(t):
        lib.go_pf_maybe_resample(pf._p, float(n + 1), O.RESAMPLE_SYSTEMATIC, 0, 0, None, None)
        obs[0] = zs[t % len(zs)]
        lib.go_pf_step(pf._p, obs, 0, None, None, None, None)

    for t in range(1, 1 + args.warmup):
        one(t)
    t0 = time.perf_counter()
    for t in range(1 + args.warmup, 1 + args.warmup + args.steps):
        one(t)
    dt = time.perf_counter() - t0
    val = n * args.steps / dt
    sample = "%d particles/step (bounded sample of the 1e8-particle workload), %d OpenMP threads" % (n, threads)
    out = {
        "impl": "reference", "metric": "particle-steps/sec", "value": val, "unit": "particle-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": args.scaling or ("strong" if args.gpus > 1 else "weak"), "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": dict(workload_config(args, n_total=n), particles_note="the CPU arm times a bounded sample of the 1e8-particle "
                       "workload (rate is per particle-step, i.e. size-normalised)", workload_particles=int(args.particles)),
        "cpu_baseline": {"value": val, "unit": "particle-steps/s", "cores": threads, "kind": "port", "sample": sample,
                         "note": "Gen.jl is pure Julia and Julia is absent from this image; this is the C oracle "
                                 "restating Gen's particle filter (optimistic stand-in: no trace allocation/dispatch)"},
        "e2e": {"value": val, "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(out)


def workload_config(args, n_total):
    return {"workload": "bearings-only bootstrap PF (BASELINE.json configs[2]): 4-D state, nonlinear obs, "
                        "systematic resampling every step, history off",
            "particles": int(n_total), "state_dim": S_BEARINGS, "resampler": "systematic", "resample": "every step",
            "l2": "working set (%.1f GB) >> 126 MB L2" % (n_total / max(1, args.gpus) * 84e-9),
            "parallelism": "particles sharded over %d GPU(s)" % args.gpus}


def log_ml_error_check(g):
    """BASELINE.json's metric ends with "log-ML error": configs[0] (1-D linear-Gaussian SSM, N = 1000, T = 100,
    systematic resampling at ESS < N/2) has an exact answer, the Kalman-filter log marginal likelihood computed here."""
    import math
    ys = F.lgssm_series(101)
    m0, s0, a, q, r = F.LGSSM["m0"], F.LGSSM["s0"], F.LGSSM["a"], F.LGSSM["q"], F.LGSSM["r"]
    m, P, exact = m0, s0 * s0, 0.0
    for t, y in enumerate(ys):
        if t > 0:
            m, P = a * m, a * a * P + q * q
        S = P + r * r
        exact += -0.5 * (math.log(2 * math.pi * S) + (y - m) ** 2 / S)
        m, P = m + P / S * (y - m), P * (1 - P / S)
    model = g.lgssm_model(**F.LGSSM)
    out = {}
    for dtype in ("f64", "f32"):
        ests = []
        for seed in range(32):
            st = g.initialize_particle_filter(model, (0,), ys[0], 1000, dtype=dtype, seed=seed)
            for t in range(1, len(ys)):
                g.maybe_resample_(st, ess_threshold=500, resampler="systematic")
                g.particle_filter_step_(st, (t,), (g.UnknownChange,), ys[t], return_incr=False)
            ests.append(g.log_ml_estimate(st))
        mean = sum(ests) / len(ests)
        sd = (sum((e - mean) ** 2 for e in ests) / (len(ests) - 1)) ** 0.5
        out[dtype] = {"mean_error": mean - exact, "sd_over_seeds": sd, "standard_error": sd / len(ests) ** 0.5}
    return {"config": "BASELINE.json configs[0]: 1-D linear-Gaussian SSM, N=1000, T=100, 32 seeds, systematic @ ESS<N/2",
            "exact": "Kalman filter log marginal likelihood", "exact_value": exact, **out,
            "note": "E[log Z_hat] <= log Z (Jensen): the expected bias is about -var/2"}


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def ncu_traffic(dtype, S):
    """Measured DRAM bytes per particle of the dominant kernel from THIS round's committed `ncu --set full` capture
    (profiles/r2_step_kernel_traffic.json, written by tools/ncu_traffic.py from the .ncu-rep): applies to the captured
    configuration only (same dtype / state size); anything else reports null."""
    p = os.path.join(ROOT, "profiles", "r2_step_kernel_traffic.json")
    try:
        d = json.load(open(p))
        if d.get("dtype") == dtype and int(d.get("state_dim", -1)) == S:
            return d
    except Exception:
        pass
    return None


def time_bearings_filter(g, torch, dist, args, n_total, world, zs, steps, warmup, sampler=None):
    """Device-timed throughput + the end-to-end loop of the bearings filter at n_total particles over `world` ranks."""
    from genb200 import _lib as L
    thr = n_total + 1   # ESS <= N < N + 1: resample every step (the accounting of SURVEY.md section 8d)
    model = g.bearings_model()
    if world == 1:
        pf = g.initialize_particle_filter(model, (0,), zs[0], n_total, dtype=args.dtype, seed=20242)

        def step(t, sync_scalar=False):
            g.maybe_resample_(pf, ess_threshold=thr, resampler="systematic")
            g.particle_filter_step_(pf, (t,), (g.UnknownChange,), zs[t], return_incr=False)
            if sync_scalar:
                return g.log_ml_estimate(pf)

        def run_many(t0, k):
            for t in range(t0, t0 + k):
                step(t)
        state, spf = pf, None
    else:
        from genb200 import sharded
        spf = sharded.ShardedParticleFilter(model, n_total, dtype=args.dtype, seed=20242)
        spf.initialize(zs[0])

        def step(t, sync_scalar=False):
            spf.maybe_resample_(ess_threshold=thr)
            spf.particle_filter_step_((t,), zs[t])
            if sync_scalar:
                return spf.log_ml_estimate()

        def run_many(t0, k):
            # device-resident path: statistics exchange, plan, ESS test, log-ML and offspring exchange on the devices
            spf.run_([zs[t] for t in range(t0, t0 + k)], ess_threshold=thr)
        state = spf.local

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # NCCL calls inside the timed regions, counted (the device-controlled run exchanges through peer memory: zero expected)
    coll = {"n": 0}
    coll_names = ("all_reduce", "all_gather", "all_gather_into_tensor", "all_to_all", "all_to_all_single", "broadcast", "reduce",
                  "reduce_scatter", "reduce_scatter_tensor", "send", "recv", "isend", "irecv", "batch_isend_irecv", "barrier")
    coll_orig = {}
    if dist is not None:
        def counted(fn):
            def w(*a, **k):
                coll["n"] += 1
                return fn(*a, **k)
            return w
        for nm in coll_names:
            if hasattr(dist, nm):
                coll_orig[nm] = getattr(dist, nm)
                setattr(dist, nm, counted(coll_orig[nm]))

    t = 1
    run_many(t, warmup)
    t += warmup
    state.sync()
    # ---- device-timed region: value (state resident in HBM; per step the host passes the 8-byte observation and, on one
    #      GPU, reads the 48-byte weight statistics maybe_resample!'s Bool needs) ----
    state.enable_timing(True)
    state.get_timing()
    if sampler is not None:
        sampler.wait_ready()
    barrier()
    w0 = time.perf_counter()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    run_many(t, steps)
    t += steps
    ev1.record()
    barrier()
    clocks = None
    if sampler is not None:
        sampler.window(w0, time.perf_counter())
        clocks = sampler.stop()          # (before the end-to-end loop: polling nvidia-smi perturbs its host round trips)
    ms = ev0.elapsed_time(ev1)
    timing = state.get_timing()
    state.enable_timing(False)
    # per-kernel events serialise the launches (no programmatic-dependent-launch overlap): a second, uninstrumented pass
    # of the same K steps gives the production number; `value` is the slower (instrumented) of the two only if it IS slower
    barrier()
    u0, u1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    u0.record()
    coll["n"] = 0
    run_many(t, steps)
    coll_run = coll["n"]
    t += steps
    u1.record()
    barrier()
    ms_plain = u0.elapsed_time(u1)
    # ---- end to end through the public API with host buffers: every step passes the observation from the host and reads the
    #      step's result (log-ML estimate: the weight statistics, device -> host) back ----
    if world > 1:
        for _ in range(2):   # first use of the synchronous path (peer mappings) stays untimed
            step(t, sync_scalar=True)
            t += 1
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    lml = None
    coll["n"] = 0
    for _ in range(steps):
        lml = step(t, sync_scalar=True)
        t += 1
    coll_e2e = coll["n"]
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    for nm, fn in coll_orig.items():
        setattr(dist, nm, fn)
    # the variant in which the caller also LOOKS at what particle_filter_step! returns (the N log incremental weights)
    ms_incr = None
    if world == 1 and n_total <= 2 * 10 ** 8:
        barrier()
        i0, i1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        i0.record()
        k_incr = max(2, steps // 5)
        for _ in range(k_incr):
            g.maybe_resample_(pf, ess_threshold=thr, resampler="systematic")
            incr, = g.particle_filter_step_(pf, (t,), (g.UnknownChange,), zs[t], return_incr=True)
            t += 1
        i1.record()
        barrier()
        ms_incr = i0.elapsed_time(i1) / k_incr
    if dist is not None:
        tt = torch.tensor([ms, ms_plain, ms_e2e], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms, ms_plain, ms_e2e = (float(x) for x in tt.tolist())
    out = {"n_total": n_total, "ms": ms, "ms_plain": ms_plain, "ms_e2e": ms_e2e, "ms_e2e_incr_per_step": ms_incr, "timing": timing,
           "clocks": clocks, "lml": lml,
           "collectives": None if dist is None else {
               "torch_distributed_calls_in_timed_region": coll_run, "per_step_in_e2e_region": coll_e2e / max(1, steps),
               "how": "every torch.distributed collective / point-to-point entry wrapped with a counter; the timed region is K steps of "
                      "the device-controlled run (max, weight sums and offspring rows travel through peer memory written by the "
                      "filter's own kernels); the barriers that bracket the region are outside the count"}}
    del state
    if world == 1:
        del pf
    else:
        del spf
    g.cache_trim()
    return out


def sharded_parity_check(g, torch, dist, rank):
    """SCALE lines carry their own proof: an untimed strong-sharded filter (synchronous calls AND the device-controlled run)
    against the same filter on ONE GPU (rank 0), compared bit for bit."""
    from genb200 import sharded
    n, T, seed = 400003, 10, 77
    ys = F.bearings_series(T + 1)
    model = g.bearings_model()
    res = {}
    for mode in ("sync", "run"):
        spf = sharded.ShardedParticleFilter(model, n, seed=seed).initialize(ys[0])
        if mode == "sync":
            nres = 0
            for t in range(1, T + 1):
                nres += spf.maybe_resample_(ess_threshold=(n + 1 if t % 2 else n / 2))
                spf.particle_filter_step_((t,), ys[t])
        else:
            nres = spf.run_(ys[1:], ess_threshold=n / 2)
        cols = [spf.get_state(f) for f in range(4)]
        par, lw, lml = spf.get_parents(), spf.get_log_weights(), spf.log_ml_estimate()
        if rank == 0:
            st = g.initialize_particle_filter(model, (0,), ys[0], n, seed=seed)
            if mode == "sync":
                nres1 = 0
                for t in range(1, T + 1):
                    nres1 += g.maybe_resample_(st, ess_threshold=(n + 1 if t % 2 else n / 2), resampler="systematic")
                    g.particle_filter_step_(st, (t,), (g.UnknownChange,), ys[t], return_incr=False)
            else:
                nres1 = g.particle_filter_run_(st, ys[1:], ess_threshold=n / 2)
            lml1 = g.log_ml_estimate(st)
            res[mode] = {"traces_bit_exact": bool(all(np.array_equal(cols[f], st.get_state(f)) for f in range(4))),
                         "parents_bit_exact": bool(np.array_equal(par, st.parents)),
                         "log_weights_bit_exact": bool(np.array_equal(lw, st.log_weights)),
                         "resamples": [int(nres), int(nres1)], "log_ml_rel_err": abs(lml - lml1) / max(1.0, abs(lml1))}
            del st
        del spf
    g.cache_trim()
    if rank != 0:
        return None
    ok = all(v["traces_bit_exact"] and v["parents_bit_exact"] and v["log_weights_bit_exact"] and v["resamples"][0] == v["resamples"][1]
             and v["log_ml_rel_err"] <= 1e-10 for v in res.values())
    return {"bit_exact": ok, "particles": n, "steps": T, "against": "the same filter on one GPU (rank 0), ordinary API",
            "paths": res}


def other_configs(g, torch, peak):
    """BASELINE.json configs[0], [1], [3], [4] on one GPU, bounded (a few seconds in total): driver-timed companions of the
    headline config.  `frac` = algorithmic bytes (SURVEY.md 8d per-unit figures) / time / measured HBM peak."""
    import ctypes as C
    from genb200 import _lib as L
    out = []

    def ev_time(fn, reps=1):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    # cfg 0: LGSSM N = 1000, T = 100 (latency regime: single-CTA fused run)
    ys = F.lgssm_series(101)
    model = g.lgssm_model(**F.LGSSM)
    st = g.initialize_particle_filter(model, (0,), ys[0], 1000, seed=1)
    g.particle_filter_run_(st, ys[1:], ess_threshold=500)
    st = g.initialize_particle_filter(model, (0,), ys[0], 1000, seed=2)
    ms = ev_time(lambda: g.particle_filter_run_(st, ys[1:], ess_threshold=500))
    out.append({"config": "configs[0]: LGSSM N=1e3 T=100 fp64, systematic @ ESS<N/2, gb_pf_run (one single-CTA kernel)",
                "us_per_step": 1e3 * ms / 100, "value": 1000 * 100 / (ms * 1e-3), "unit": "particle-steps/s", "frac": None,
                "bound": "launch latency"})
    # cfg 1: SV N = 1e6, T = 1000, fp64 and fp32, resampling every step, fused run
    ys = F.sv_series(1001)
    model = g.sv_model(**F.SV)
    for dtype, E in (("f64", 8), ("f32", 4)):
        n = 10 ** 6
        st = g.initialize_particle_filter(model, (0,), ys[0], n, dtype=dtype, seed=3)
        g.particle_filter_run_(st, ys[1:51], ess_threshold=n + 1)
        ms = ev_time(lambda: g.particle_filter_run_(st, ys[51:1001], ess_threshold=n + 1))
        b = algorithmic_bytes_per_particle_step(1, E)
        out.append({"config": "configs[1]: stochastic volatility N=1e6 T=1000 %s, systematic every step, gb_pf_run" % dtype,
                    "us_per_step": 1e3 * ms / 950, "value": n * 950 / (ms * 1e-3), "unit": "particle-steps/s",
                    "algorithmic_bytes_per_unit": b, "frac": b * n * 950 / (ms * 1e-3) / 1e9 / peak, "bound": "launch latency / hbm",
                    "log_ml_estimate": g.log_ml_estimate(st)})
        del st
    # cfg 3: BLR importance sampling D = 64, 1e7 samples
    X, y, sigma = F.blr_problem(D=64, n=64)
    model = g.blr_model(X, sigma)
    n = 10 ** 7
    for dtype, E in (("f64", 8), ("f32", 4)):
        st = g.create_particle_filter(model, n, dtype=dtype, seed=4)
        g.generate_(st, y)
        st.ess()
        ms = ev_time(lambda: (g.generate_(st, y), st.ess()), reps=3)
        del st
        t0 = time.perf_counter()
        _, lnw, lml = g.importance_sampling(model, (), y, n, dtype=dtype, seed=4, return_traces=False)
        wall = time.perf_counter() - t0
        b = (64 + 1) * E
        out.append({"config": "configs[3]: Bayesian linear regression importance_sampling D=64 n=64, 1e7 samples %s" % dtype,
                    "ms": ms, "value": n / (ms * 1e-3), "unit": "samples/s", "algorithmic_bytes_per_unit": b,
                    "frac": b * n / (ms * 1e-3) / 1e9 / peak, "bound": "fp64 ALU (8192 likelihood flop + 32 Box-Muller pairs per sample)",
                    "likelihood_tflops": 2.0 * 64 * 64 * n / (ms * 1e-3) / 1e12,
                    "e2e_ms_with_80MB_weights_to_host": 1e3 * wall, "log_ml_estimate": lml})
        del lnw
    # cfg 4: resampling microbenchmark, device-resident log weights ~ N(0, 3^2)
    for n in (10 ** 6, 10 ** 8):
        ld = (n + 2047) // 2048 * 2048
        gen = torch.Generator(device="cuda").manual_seed(20244)
        lw = torch.zeros(ld, dtype=torch.float64, device="cuda")
        lw[:n] = 3.0 * torch.randn(n, dtype=torch.float64, device="cuda", generator=gen)
        anc = torch.empty(ld, dtype=torch.int32, device="cuda")
        for kind, name in ((L.RESAMPLE_SYSTEMATIC, "systematic"), (L.RESAMPLE_RESIDUAL, "residual"),
                           (L.RESAMPLE_MULTINOMIAL, "multinomial (reference semantics: iid draws in slot order)"),
                           (L.RESAMPLE_MULTINOMIAL_SORTED, "multinomial_sorted (the same draws' distribution, sorted by ancestor)")):
            def call():
                L.check(L.lib().gb_resample_device(C.c_void_p(lw.data_ptr()), n, L.F64, kind, 12345, None, 7, 3, C.c_void_p(anc.data_ptr()), None))
            call()
            ms = ev_time(call, reps=3)
            out.append({"config": "configs[4]: resample %s, N=%.0e fp64 log weights resident in HBM" % (name, n), "ms": ms,
                        "value": n / (ms * 1e-3), "unit": "weights/s", "algorithmic_bytes_per_unit": 12,
                        "frac": 12.0 * n / (ms * 1e-3) / 1e9 / peak, "bound": "hbm",
                        "note": "includes the max + statistics passes the raw-array entry needs (inside a filter the step kernel "
                                "provides the max)"})
        del lw, anc
    g.cache_trim()
    return out


def run_genb200(args):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    sampler = ClockSampler(local_rank)
    sampler.start()                      # nvidia-smi needs a few 100 ms to come up: start it before the filter is built
    g = entry.load_package()
    from genb200 import _lib as L
    L.check(L.lib().gb_set_device(local_rank))
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))   # plumbing: IPC handles, barriers, max of the timings

    E = 8 if args.dtype == "f64" else 4
    scaling = args.scaling or ("strong" if world > 1 else "weak")
    n_total = int(args.particles) * (world if scaling == "weak" else 1)
    zs = F.bearings_series(args.warmup + 4 * args.steps + 16)
    parity = sharded_parity_check(g, torch, dist, rank) if world > 1 else None
    main = time_bearings_filter(g, torch, dist, args, n_total, world, zs, args.steps, args.warmup, sampler)
    other = None
    if world > 1 and not args.no_second_scaling:
        n_other = int(args.particles) * (1 if scaling == "weak" else world)
        other = time_bearings_filter(g, torch, dist, args, n_other, world, zs, args.steps, args.warmup)
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peak()
    S = S_BEARINGS
    n_local = n_total // world
    ms = min(main["ms"], main["ms_plain"])
    ms_per_step = ms / args.steps
    value = n_total * args.steps / (ms * 1e-3)
    timing = main["timing"]
    step_name = L.lib().gb_kernel_name(L.K_STEP).decode()
    k_ms, k_launches = timing[step_name]
    avg_ms = k_ms / max(1, k_launches)
    # bytes the fused step kernel has to move per particle: ancestor index + gathered row in, new row + log weight out
    # = 4 + 2 S E + E (76 for S = 4, fp64).  After a resample the log weights restart from exactly 0.0, so the step's log
    # incremental weights ARE its new log weights: they are stored once and gb_pf_get_log_increments reads that column (up
    # to r2t the kernel wrote both, 84 B).  This is what `frac` is a fraction OF.
    fused_bytes = (4 + 2 * S * E + E) * n_local
    achieved = fused_bytes / (avg_ms * 1e-3) / 1e9 if avg_ms > 0 else 0.0
    b_step = algorithmic_bytes_per_particle_step(S, E)                    # SURVEY.md 8d accounting of the UNFUSED API calls: 176
    # fused step (76) + statistics: log weights in, quantised weights out (2 E) + ancestors: quantised weights in, indices out
    # (E + 4): 104 for S = 4, fp64
    b_necessary = (4 + 2 * S * E + E) + 2 * E + (E + 4)
    kernels = {k: {"ms_per_step": v[0] / args.steps, "launches_per_step": v[1] / args.steps} for k, v in timing.items() if v[1]}
    launches = int(sum(v[1] for v in timing.values()))
    tr = ncu_traffic(args.dtype, S)

    def whole(nbytes):
        gbs = nbytes * n_local / (ms_per_step * 1e-3) / 1e9
        return {"bytes_per_particle_step": nbytes, "achieved": gbs, "frac": gbs / peak, "frac_of_8TBs_spec": gbs / 8000.0}

    # ---- CPU baseline (rank 0, N = 1 only): bounded sample of the same workload ----
    cpu = None
    if world == 1 and not args.no_cpu:
        O = entry.load_oracle()
        threads = O.lib().go_get_max_threads()
        rate, secs = cpu_pf_rate(args.cpu_particles, 10, threads, zs)
        rate1, secs1 = cpu_pf_rate(args.cpu_particles // 4, 4, 1, zs)
        cpu = {"value": rate, "unit": "particle-steps/s", "cores": threads, "kind": "port",
               "sample": "%d particles x 10 steps of the same bearings PF (%.1f s); C oracle restating Gen's PF"
                         % (args.cpu_particles, secs),
               "single_thread_value": rate1,
               "note": "Gen.jl (pure Julia) cannot run in this image; tutorial-derived Gen.jl figure is "
                       "4e3-8e3 particle-steps/s (docs/src/tutorials/smc.md:298-300)"}

    def second(o, label):
        if o is None:
            return None
        m2 = min(o["ms"], o["ms_plain"])
        return {"scaling": label, "particles": o["n_total"], "value": o["n_total"] * args.steps / (m2 * 1e-3), "ms_per_step": m2 / args.steps,
                "e2e_value": o["n_total"] * args.steps / (o["ms_e2e"] * 1e-3), "e2e_ms_per_step": o["ms_e2e"] / args.steps}

    out = {
        "metric": "particle-steps/sec", "value": value, "unit": "particle-steps/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": scaling, "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
        "config": workload_config(args, n_total),
        "clocks": main["clocks"],
        "e2e": {"value": n_total * args.steps / (main["ms_e2e"] * 1e-3), "unit": "particle-steps/s", "h2d_bytes_per_step": 8,
                "d2h_bytes_per_step": 48 if world == 1 else 628,
                "ms_per_step": main["ms_e2e"] / args.steps, "log_ml_estimate": main["lml"],
                "note": "maybe_resample_ + particle_filter_step_ + log_ml_estimate through the host API each step: observation "
                        "passed from host memory, weight statistics read back; the N log incremental weights "
                        "particle_filter_step! returns stay on the device as a lazy vector (gb_pf_get_log_increments)",
                "reading_log_increments": None if main["ms_e2e_incr_per_step"] is None else {
                    "ms_per_step": main["ms_e2e_incr_per_step"], "value": n_total / (main["ms_e2e_incr_per_step"] * 1e-3),
                    "d2h_bytes_per_step": 8 * n_total + 48,
                    "note": "the same loop when the caller materialises the returned vector every step (N x 8 bytes device -> host)"}},
        "gpu_launches": launches,
        "timing_note": "ms_per_step = the faster of two passes of K steps: with per-kernel CUDA events (which serialise the "
                       "launches) and without (programmatic dependent launch overlaps launch latencies)",
        "ms_per_step_instrumented": main["ms"] / args.steps, "ms_per_step_uninstrumented": main["ms_plain"] / args.steps,
        "roofline": {"bound": "hbm", "kernel": step_name, "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "peak_source": peak_src,
                     "bytes_per_launch": fused_bytes, "bytes_per_particle": fused_bytes / n_local, "avg_launch_ms": avg_ms,
                     "traffic": None if tr is None else tr["dram_bytes_per_particle"] * n_local,
                     "traffic_source": None if tr is None else {k: tr[k] for k in ("file", "kernel", "particles", "dram_bytes_per_particle")},
                     "note": "bytes = what the FUSED kernel must move per particle (ancestor index + gathered row in; row + log weight "
                             "out; the log incremental weights equal the new log weights after a resample and share their column); "
                             "frac = bytes / measured time / measured copy peak",
                     "whole_step": {"algorithmic": dict(whole(b_step), note="SURVEY.md 8d accounting of the unfused API calls "
                                                                           "(step! + maybe_resample! with a materialised gather): the deferred-gather fusion removes a "
                                                                           "write + read of the state, so this can exceed 1"),
                                    "necessary": dict(whole(b_necessary), note="bytes the fused pipeline really moves: step kernel + "
                                                                             "statistics pass (log weights in, quantised weights out) + "
                                                                             "ancestors pass (quantised weights in, indices out)")}},
        "kernels": kernels,
        "cpu_baseline": cpu,
        "sharded_parity": parity,
        "collectives": main.get("collectives"),
        "weak" if scaling == "strong" else "strong": second(other, "weak" if scaling == "strong" else "strong"),
        "configs": other_configs(g, torch, peak) if world == 1 and not args.no_configs else None,
        "log_ml_error": log_ml_error_check(g) if world == 1 else None,
    }
    emit(out)
    if dist is not None:
        dist.destroy_process_group()


_REAL_STDOUT = None


def emit(obj):
    """The ONE JSON line goes to the real stdout; everything else the libraries print (NCCL banner, ...)
    was diverted to stderr."""
    line = json.dumps(obj) + "\n"
    if _REAL_STDOUT is not None:
        os.write(_REAL_STDOUT, line.encode())
    else:
        sys.stdout.write(line)
        sys.stdout.flush()


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="genb200", choices=["genb200", "reference"])
    ap.add_argument("--particles", type=float, default=1e8, help="particles per GPU (weak scaling) / in total (strong scaling)")
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"])
    ap.add_argument("--scaling", default=None, choices=["strong", "weak"],
                    help="strong (default at --gpus > 1, BASELINE.json configs[2]): 1e8 particles in total, sharded; weak: 1e8 "
                         "particles PER GPU.  The other one is measured too and reported under its own key.")
    ap.add_argument("--no-second-scaling", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the bounded runs of the other BASELINE configs (N = 1)")
    ap.add_argument("--cpu-particles", type=int, default=20_000_000)
    ap.add_argument("--ref-particles", type=int, default=20_000_000)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_genb200(args)


if __name__ == "__main__":
    main()
