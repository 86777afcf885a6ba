"""Bearings-only tracking (docs/src/tutorials/smc.md:607-689) through the Python mirror of Gen's particle-filter API.

    python examples/pf_bearings.py [num_particles] [num_samples]

The same program in Gen.jl:

    state = initialize_particle_filter(model, (0,), choicemap((:z0, zs[1])), N)
    for t = 1:T
        maybe_resample!(state, ess_threshold=N/2)
        particle_filter_step!(state, (t,), (UnknownChange(),), choicemap((:chain => t => :z, zs[t+1])))
    end
    sample_unweighted_traces(state, num_samples)
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as entry  # noqa: E402
import fixtures as F  # noqa: E402


def main():
    n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10 ** 6
    num_samples = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    g = entry.load_package()
    zs = F.bearings_series(51)
    model = g.bearings_model()
    g.log_ml_estimate(g.initialize_particle_filter(model, (0,), {"z0": zs[0]}, 1000))   # CUDA context + module load
    t0 = time.perf_counter()
    state = g.initialize_particle_filter(model, (0,), {"z0": zs[0]}, n, seed=1)
    resamples = 0
    for t in range(1, len(zs)):
        resamples += g.maybe_resample_(state, ess_threshold=n / 2)
        g.particle_filter_step_(state, (t,), (g.UnknownChange,), {("chain", t, "z"): zs[t]}, return_incr=False)
        g.rejuvenate_(state, zs[t], move="drift", drift_std=1e-3)        # random-walk MH on the newest velocities
    lml = g.log_ml_estimate(state)
    dt = time.perf_counter() - t0
    print("N = %d, T = %d: log marginal likelihood estimate %.6f, %d resamples, %.1f ms (%.3g particle-steps/s)"
          % (n, len(zs) - 1, lml, resamples, 1e3 * dt, n * (len(zs) - 1) / dt))
    idx, rows = g.sample_unweighted_traces(state, num_samples)
    for k in range(num_samples):
        print("  sample %d: particle %d, final (x, y, vx, vy) = (%.4f, %.4f, %.5f, %.5f)" % ((k, idx[k]) + tuple(rows[k])))


if __name__ == "__main__":
    main()
