"""Generates the committed golden vectors under tests/golden/ from the CPU oracle (oracle/gen_oracle.c).

Gen.jl is Julia and cannot run in this environment, so these are NOT outputs of the reference itself: they freeze the
oracle's outputs (which are pinned on the reference's own known-answer tests, tests/test_oracle_golden.py) so that
  * any later change of the oracle is caught on the CPU (tests/test_golden_fixtures.py::test_oracle_reproduces_*), and
  * the CUDA path is compared with committed numbers on the GPU box, where /root/reference does not exist.
Run:  python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as entry  # noqa: E402
import fixtures as F  # noqa: E402

O = entry.load_oracle()


def resampling():
    rng = np.random.default_rng(20244)
    out = {}
    for n in (7, 257, 4099):
        lw = 3.0 * rng.standard_normal(n)
        u0 = int(rng.integers(0, 2 ** 32))
        us = rng.integers(0, 2 ** 64, size=n, dtype=np.uint64)
        out[str(n)] = {
            "logw": [float(v) for v in lw], "u0": u0, "u64": [int(v) for v in us],
            "systematic": [int(v) for v in O.resample(O.RESAMPLE_SYSTEMATIC, lw, u0=u0)],
            "residual": [int(v) for v in O.resample(O.RESAMPLE_RESIDUAL, lw, u0=u0)],
            "multinomial": [int(v) for v in O.resample(O.RESAMPLE_MULTINOMIAL, lw, u64s=us)],
            "logsumexp": O.logsumexp(lw),
            "ess": O.effective_sample_size(O.normalize_weights(lw)[1]),
        }
    json.dump(out, open(os.path.join(HERE, "resampling.json"), "w"))


def bearings_pf():
    n, T = 512, 5
    rng = np.random.default_rng(20242)
    zs = F.bearings_series(T + 1)
    eps0 = rng.standard_normal((4, n))
    eps = rng.standard_normal((T, 2, n))
    u0s = [int(v) for v in rng.integers(0, 2 ** 32, size=T)]
    pf = O.OraclePF.init(O.MODEL_BEARINGS, F.BEARINGS_PARAMS, n, [zs[0]], normals=eps0)
    did = []
    for t in range(1, T + 1):
        did.append(bool(pf.maybe_resample(ess_threshold=n / 2 if t % 2 else n + 1, kind=O.RESAMPLE_SYSTEMATIC, u0=u0s[t - 1])))
        pf.step([zs[t]], normals=eps[t - 1])
    np.savez_compressed(os.path.join(HERE, "bearings_pf.npz"), zs=zs, eps0=eps0, eps=eps, u0s=np.array(u0s, dtype=np.int64),
                        did=np.array(did), state=np.stack([pf.state(f) for f in range(4)]), parents=pf.parents,
                        logw=pf.log_weights, log_ml_est=pf.log_ml_est, log_ml=pf.log_ml_estimate())


def scalars():
    ys = F.lgssm_series(101)
    X, y, sigma = F.blr_problem()
    json.dump({
        "hmm_log_ml": float(np.log(O.hmm_forward(F.HMM_PRIOR, F.HMM_EMISSION, F.HMM_TRANSITION, F.HMM_OBS))),
        "hmm_log_ml_reference_value": F.HMM_EXPECTED_LOG_ML,
        "kalman_log_ml_cfg1": O.kalman_logml(0.0, 1.0, 0.9, 0.5, 1.0, ys),
        "blr_log_evidence_cfg4": O.blr_log_evidence(X, y, sigma),
        "normal_logpdf_1e13": O.lib().go_logpdf_normal(1e13, 0.0, 1.0),
    }, open(os.path.join(HERE, "scalars.json"), "w"), indent=1)


if __name__ == "__main__":
    resampling()
    bearings_pf()
    scalars()
    print("golden vectors written to", HERE)
