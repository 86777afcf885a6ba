#!/usr/bin/env python
"""bench.py -- particle-steps/sec of the SMC hot path (BASELINE.json metric) on N B200s.

Workload (config.workload): BASELINE.json configs[2], the 2-D bearings-only tracker
(docs/src/tutorials/smc.md:607-661; 4-D state, nonlinear observation) bootstrap particle filter with
N = 10^8 particles (fp64) and systematic resampling EVERY step.  One "step" = one particle-filter
step over all particles: maybe_resample! (ESS, ancestors, resampled-trace gather) followed by
particle_filter_step! (propose + weight + log-sum-exp partials).  At --gpus G > 1 the particles are
sharded G ways (one process per GPU, global-prefix systematic resampling, NCCL exchange).

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
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, n_total=args.particles),
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
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    E = 8 if args.dtype == "f64" else 4
    n_total = int(args.particles) * (world if args.scaling == "weak" else 1)
    zs = F.bearings_series(args.warmup + 2 * args.steps + 12)
    model = g.bearings_model()
    if world == 1:
        pf = g.initialize_particle_filter(model, (0,), zs[0], n_total, dtype=args.dtype, seed=20242)
        thr = n_total + 1   # ESS <= N < N + 1: resample every step (the accounting of SURVEY.md section 8d)

        def step(t, sync_scalar=False):
            g.maybe_resample_(pf, ess_threshold=thr, resampler="systematic")
            g.particle_filter_step_(pf, (t,), (g.UnknownChange,), zs[t], return_incr=False)
            if sync_scalar:
                return g.log_ml_estimate(pf)
        state = pf
    else:
        from genb200 import sharded
        spf = sharded.ShardedParticleFilter(model, n_total, dtype=args.dtype, seed=20242)
        spf.initialize(zs[0])
        thr = n_total + 1

        def step(t, sync_scalar=False):
            spf.maybe_resample_(ess_threshold=thr)
            spf.particle_filter_step_((t,), zs[t])
            if sync_scalar:
                return spf.log_ml_estimate()

        def run_many(t0, k):
            # device-resident path: plan / ESS test / log-ML on the device, no host round trip per step
            spf.run_([zs[t] for t in range(t0, t0 + k)], ess_threshold=thr)
        state = spf.local

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    t = 1
    if world == 1:
        for _ in range(args.warmup):
            step(t)
            t += 1
    else:
        run_many(t, args.warmup)
        t += args.warmup
    state.sync()

    # ---- device-timed region: value (state resident in HBM, no per-step host read beyond the API's own
    #      8-byte observation argument and the 40-byte weight-statistics read maybe_resample! needs) ----
    state.enable_timing(True)
    state.get_timing()
    sampler.wait_ready()
    barrier()
    w0 = time.perf_counter()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    if world == 1:
        for _ in range(args.steps):
            step(t)
            t += 1
    else:
        run_many(t, args.steps)
        t += args.steps
    ev1.record()
    barrier()
    sampler.window(w0, time.perf_counter())
    clocks = sampler.stop()              # (before the end-to-end loop: polling nvidia-smi perturbs its host round trips)
    ms = ev0.elapsed_time(ev1)
    timing = state.get_timing()
    state.enable_timing(False)
    if dist is not None:
        tt = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt.item())
    ms_per_step = ms / args.steps
    value = n_total * args.steps / (ms * 1e-3)

    # ---- end to end through the public API with host buffers: every step passes the observation from
    #      the host and reads the step's result (log-ML estimate: weight statistics D2H) back ----
    if world > 1:
        for _ in range(2):   # first use of the synchronous path's collectives (NCCL p2p channel set-up) stays untimed
            step(t, sync_scalar=True)
            t += 1
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    lml = None
    for _ in range(args.steps):
        if world > 1 and args.e2e_api == "run":
            spf.run_([zs[t]], ess_threshold=thr)       # one observation per call, passed from host memory
            lml = spf.log_ml_estimate()                # device -> host read of the step's result
        else:
            lml = step(t, sync_scalar=True)
        t += 1
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    if dist is not None:
        tt = torch.tensor([ms_e2e], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_e2e = float(tt.item())
    e2e_value = n_total * args.steps / (ms_e2e * 1e-3)

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peak()
    n_local = n_total // world
    b_step = algorithmic_bytes_per_particle_step(S_BEARINGS, E)
    # dominant kernel = fused (deferred gather +) propose + weight + LSE kernel.  Its algorithmic bytes
    # (DESIGN.md "Kernels"): particle_filter_step! (2S+3)E plus the resampled-trace gather it absorbs
    # (A + 2S E + E) per particle.
    step_name = L.lib().gb_kernel_name(L.K_STEP).decode()
    k_ms, k_launches = timing[step_name]
    alg_step_kernel = ((2 * S_BEARINGS + 3) * E + (4 + 2 * S_BEARINGS * E + E)) * n_local
    avg_ms = k_ms / max(1, k_launches)
    achieved = alg_step_kernel / (avg_ms * 1e-3) / 1e9 if avg_ms > 0 else 0.0
    kernels = {k: {"ms_per_step": v[0] / args.steps, "launches_per_step": v[1] / args.steps} for k, v in timing.items() if v[1]}
    launches = int(sum(v[1] for v in timing.values()))
    whole_step_gbs = b_step * n_local / (ms_per_step * 1e-3) / 1e9
    # DRAM traffic of the step kernel from the committed `ncu --set full` capture (profiles/r1e_kernels_final_ncu.txt:
    # dram__bytes_read.sum + dram__bytes_write.sum = 2.997 + 4.746 GB at N = 1e8, fp64); scales with N
    traffic = 77.43 * n_local if args.dtype == "f64" else None
    fused_io = (4 + 2 * S_BEARINGS * E + 2 * E) * n_local     # what the fused kernel itself must move: anc + rows in, rows + logw + incr out

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

    out = {
        "metric": "particle-steps/sec", "value": value, "unit": "particle-steps/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
        "config": workload_config(args, n_total),
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "particle-steps/s", "h2d_bytes_per_step": 8, "d2h_bytes_per_step": 80,
                "ms_per_step": ms_e2e / args.steps, "log_ml_estimate": lml,
                "note": "maybe_resample_ + particle_filter_step_ + log_ml_estimate through the host API each step: "
                        "observation passed from host memory, weight statistics read back twice (40 B each); the N log "
                        "incremental weights particle_filter_step! returns stay on the device as a lazy vector "
                        "(gb_pf_get_log_increments) and are not read"},
        "gpu_launches": launches,
        "roofline": {"bound": "hbm", "kernel": step_name, "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": alg_step_kernel, "avg_launch_ms": avg_ms,
                     "note": "algorithmic = particle_filter_step! (2S+3)E + the maybe_resample! gather it absorbs (A+2SE+E) per "
                             "particle (SURVEY.md 8d); the deferred-gather fusion removes an intermediate write+read of the state, "
                             "so real DRAM traffic (`traffic`, ncu) is below the algorithmic bytes and frac can exceed 1",
                     "fused_io": {"bytes_per_launch": fused_io, "achieved": fused_io / (avg_ms * 1e-3) / 1e9 if avg_ms > 0 else 0.0,
                                  "frac": (fused_io / (avg_ms * 1e-3) / 1e9 / peak) if avg_ms > 0 else 0.0},
                     "whole_step": {"algorithmic_bytes_per_particle_step": b_step, "achieved": whole_step_gbs,
                                    "frac": whole_step_gbs / peak, "frac_of_8TBs_spec": whole_step_gbs / 8000.0}},
        "kernels": kernels,
        "cpu_baseline": cpu,
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
    ap.add_argument("--scaling", default="weak", choices=["strong", "weak"],
                    help="weak (default): 1e8 particles PER GPU; strong: 1e8 particles in total, sharded")
    ap.add_argument("--cpu-particles", type=int, default=20_000_000)
    ap.add_argument("--ref-particles", type=int, default=20_000_000)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--e2e-api", default="sync", choices=["sync", "run"],
                    help="multi-GPU end-to-end loop: maybe_resample_/particle_filter_step_ per step (sync) or run_ with one "
                         "observation per call (run)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_genb200(args)


if __name__ == "__main__":
    main()
