"""ctypes binding of libgenb200.so (include/genb200.h).

The library is the product; this module only declares its signatures.  It raises ImportError-like
RuntimeError loudly when the shared object is missing -- there is no Python/CPU fallback path.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# GENB200_LIB selects an alternative build of the same library (kernel tuning experiments)
LIB_PATH = os.environ.get("GENB200_LIB") or os.path.join(_HERE, "libgenb200.so")

OK, EINVAL, ENOGPU, ECUDA, EUNSUPPORTED, ESTATE = range(6)
MODEL_LGSSM, MODEL_SV, MODEL_BEARINGS, MODEL_HMM, MODEL_BLR, MODEL_SPEC = 1, 2, 3, 4, 5, 6
PROPOSAL_DEFAULT, PROPOSAL_AFFINE_NORMAL, PROPOSAL_CATEGORICAL_TABLE = 0, 1, 2
RESAMPLE_MULTINOMIAL, RESAMPLE_SYSTEMATIC, RESAMPLE_RESIDUAL, RESAMPLE_MULTINOMIAL_SORTED = 0, 1, 2, 3
MH_RESIMULATE, MH_DRIFT = 0, 1
F64, F32 = 0, 1
DIST_NORMAL, DIST_GAMMA, DIST_BERNOULLI, DIST_CATEGORICAL, DIST_MVNORMAL = 1, 2, 3, 4, 5
K_STEP, K_FINALIZE, K_QTOTAL, K_TILESCAN, K_ANCESTORS, K_GATHER, K_MISC, K_COUNT = range(8)

dp = C.POINTER(C.c_double)
i64p = C.POINTER(C.c_int64)
u64p = C.POINTER(C.c_uint64)
u32p = C.POINTER(C.c_uint32)
ip = C.POINTER(C.c_int)
vp = C.c_void_p
vpp = C.POINTER(C.c_void_p)
d, i, i64, u64, u32 = C.c_double, C.c_int, C.c_int64, C.c_uint64, C.c_uint32

# name -> (restype, argtypes): one entry per symbol declared in include/genb200.h
SIGNATURES = {
    "gb_last_error": (C.c_char_p, []),
    "gb_version": (C.c_char_p, []),
    "gb_device_count": (i, [ip]),
    "gb_set_device": (i, [i]),
    "gb_model_create": (i, [i, dp, i, vpp]),
    "gb_model_free": (None, [vp]),
    "gb_model_state_dim": (i, [vp]),
    "gb_model_num_obs": (i, [vp]),
    "gb_model_num_normals": (i, [vp, i, i]),
    "gb_model_num_uniforms": (i, [vp, i, i]),
    "gb_pf_create": (i, [vp, i64, i64, i64, i, u64, vpp]),
    "gb_pf_set_noise": (i, [vp, dp, i, dp, i]),
    "gb_pf_generate": (i, [vp, dp, i, i, dp, i]),
    "gb_pf_initialize": (i, [vp, i64, i, u64, dp, i, i, dp, i, vpp]),
    "gb_pf_step": (i, [vp, dp, i, i, dp, i, dp]),
    "gb_pf_maybe_resample": (i, [vp, d, i, ip]),
    "gb_pf_enable_history": (i, [vp, i]),
    "gb_pf_history_length": (i, [vp, ip]),
    "gb_pf_get_trajectory": (i, [vp, i, dp]),
    "gb_pf_sample_unweighted": (i, [vp, i64, u64, i64p, dp]),
    "gb_pf_rejuvenate": (i, [vp, i, dp, i, d, i64p, dp]),
    "gb_pf_enable_window": (i, [vp, i]),
    "gb_pf_rejuvenate_window": (i, [vp, i, i, dp, i, d, i64p, dp]),
    "gb_pf_run": (i, [vp, i, dp, i, d, ip]),
    "gb_pf_run_mh": (i, [vp, i, dp, i, d, i, d, i, ip, i64p]),
    "gb_cache_trim": (i, []),
    "gb_cache_stats": (i, [i64p, i64p, i64p]),
    "gb_spec_jit_source": (i, [vp, C.c_char_p, i64, i64p]),
    "gb_spec_jit_compile": (i, [vp, i, i64p]),
    "gb_pf_spec_backend": (i, [vp, ip]),
    "gb_pf_set_resample_noise": (i, [vp, i, u32, u64p]),
    "gb_pf_set_resample_spacings": (i, [vp, u64p]),
    "gb_pf_log_ml_estimate": (i, [vp, dp]),
    "gb_pf_get_log_weights": (i, [vp, dp]),
    "gb_pf_get_log_increments": (i, [vp, dp]),
    "gb_pf_set_log_weights": (i, [vp, dp]),
    "gb_pf_get_parents": (i, [vp, i64p]),
    "gb_pf_get_state": (i, [vp, i, dp]),
    "gb_pf_get_log_ml_est": (i, [vp, dp]),
    "gb_pf_get_ess": (i, [vp, dp, dp]),
    "gb_pf_num_particles": (i, [vp, i64p, i64p]),
    "gb_pf_step_index": (i, [vp, ip]),
    "gb_pf_clone": (i, [vp, vpp]),
    "gb_pf_free": (None, [vp]),
    "gb_pf_set_lazy_gather": (i, [vp, i]),
    "gb_pf_set_stream": (i, [vp, vp]),
    "gb_pf_sync": (i, [vp]),
    "gb_pf_weight_max": (i, [vp, dp]),
    "gb_pf_weight_sums": (i, [vp, d, i, dp, dp, u64p]),
    "gb_pf_stats_device": (i, [vp, vpp, ip]),
    "gb_pf_weight_max_device": (i, [vp]),
    "gb_pf_weight_sums_device": (i, [vp, i]),
    "gb_pf_ipc_bundle_bytes": (i, []),
    "gb_pf_ipc_export": (i, [vp, vp, ip]),
    "gb_pf_async_attach_ipc": (i, [vp, i, vp]),
    "gb_pf_async_attach_local": (i, [vp, i, vp]),
    "gb_pf_async_setup": (i, [vp, i, i, i64p, d]),
    "gb_pf_async_arm": (i, [vp, d]),
    "gb_pf_async_cur": (i, [vp, ip]),
    "gb_pf_xchg_post": (i, [vp]),
    "gb_pf_xchg_stats": (i, [vp]),
    "gb_pf_xchg_stats_plan": (i, [vp, u32]),
    "gb_pf_async_plan": (i, [vp, u32, d, i]),
    "gb_pf_async_sweep": (i, [vp]),
    "gb_pf_async_export": (i, [vp]),
    "gb_pf_async_wait": (i, [vp]),
    "gb_pf_async_step": (i, [vp, dp, i]),
    "gb_pf_async_iterate": (i, [vp, dp, i]),
    "gb_pf_async_finish": (i, [vp, ip, ip]),
    "gb_pf_xchg_read": (i, [vp, dp, dp, u64p, ip]),
    "gb_pf_xchg_records": (i, [vp, dp, u64p]),
    "gb_pf_xchg_commit": (i, [vp, ip, dp]),
    "gb_spf_create": (i, [vp, i64, i, u64, i, ip, vpp]),
    "gb_spf_generate": (i, [vp, dp, i, i, dp, i]),
    "gb_spf_initialize": (i, [vp, i64, i, u64, i, ip, dp, i, i, dp, i, vpp]),
    "gb_spf_step": (i, [vp, dp, i, i, dp, i]),
    "gb_spf_maybe_resample": (i, [vp, d, ip]),
    "gb_spf_set_resample_noise": (i, [vp, i, u32]),
    "gb_spf_run": (i, [vp, i, dp, i, d, ip]),
    "gb_spf_log_ml_estimate": (i, [vp, dp]),
    "gb_spf_get_log_ml_est": (i, [vp, dp]),
    "gb_spf_get_ess": (i, [vp, dp, dp]),
    "gb_spf_get_log_weights": (i, [vp, dp]),
    "gb_spf_get_log_increments": (i, [vp, dp]),
    "gb_spf_get_state": (i, [vp, i, dp]),
    "gb_spf_get_parents": (i, [vp, i64p]),
    "gb_spf_num_shards": (i, [vp, ip]),
    "gb_spf_shard": (i, [vp, i, vpp, i64p]),
    "gb_spf_num_particles": (i, [vp, i64p]),
    "gb_spf_step_index": (i, [vp, ip]),
    "gb_spf_sync": (i, [vp]),
    "gb_spf_free": (None, [vp]),
    "gb_spf_importance_sampling": (i, [vp, i64, i, u64, i, ip, dp, i, i, dp, i, dp, dp, vpp]),
    "gb_spf_resample": (i, [dp, i64, i, i, ip, u32, i64p]),
    "gb_pf_systematic_offspring": (i, [vp, d, i, u64, u64, u64, u32, i64p, i64p]),
    "gb_pf_block_bytes": (i64, [vp, i64]),
    "gb_pf_ext_capacity": (i, [vp, i64p]),
    "gb_pf_offspring_local": (i, [vp, i64, i64]),
    "gb_pf_import_ext": (i, [vp, i64, i64, i64, vp]),
    "gb_pf_commit_resample_lazy": (i, [vp, d]),
    "gb_pf_async_commit": (i, [vp, d]),
    "gb_pf_export_offspring": (i, [vp, i64, i64, vp]),
    "gb_pf_import_rows": (i, [vp, i64, i64, vp]),
    "gb_pf_commit_resample": (i, [vp, d]),
    "gb_importance_sampling": (i, [vp, i64, i, u64, dp, i, i, dp, i, dp, i, dp, i, dp, dp, vpp]),
    "gb_importance_resampling": (i, [vp, i64, i, u64, dp, i, i, dp, i, i64, dp, dp]),
    "gb_enumerative_normalize": (i, [dp, dp, i64, i, dp, dp]),
    "gb_resample": (i, [dp, i64, i, i, u32, u64p, u64, u32, i64p]),
    "gb_resample_device": (i, [vp, i64, i, i, u32, vp, u64, u32, vp, vp]),
    "gb_logsumexp": (i, [dp, i64, i, dp, dp]),
    "gb_dist_logpdf": (i, [i, i, i64, dp, dp, i, dp, i, i, dp]),
    "gb_dist_random": (i, [i, i, i64, dp, i, dp, i, i, u64, u32, dp]),
    "gb_philox_normals": (i, [u64, u32, i64, i64, i, i, dp]),
    "gb_philox_uniforms": (i, [u64, u32, i64, i64, i, i, dp]),
    "gb_philox_resample_u0": (u32, [u64, u32]),
    "gb_philox_multinomial_u64": (i, [u64, u32, i64, i64, u64p]),
    "gb_philox_sorted_spacings": (i, [u64, u32, i64, i64, u64p]),
    "gb_quant_bits": (i, [i64]),
    "gb_pf_enable_timing": (i, [vp, i]),
    "gb_pf_get_timing": (i, [vp, dp, i64p]),
    "gb_kernel_name": (C.c_char_p, [i]),
    "gb_selftest_slots_below": (i64, [u64, i64, u64, u32]),
    "gb_selftest_slots_below_residual": (i64, [u64, u64, u64, u32]),
    "gb_selftest_div128": (u64, [u64, u64, u64]),
    "gb_selftest_quantize": (u64, [d, i]),
    "gb_selftest_math": (d, [i, d, d]),
    "gb_selftest_philox": (None, [u32p, u32p, u32p]),
}

_lib = None


class GenB200Error(RuntimeError):
    """Raised for every non-zero status of the C ABI (mirrors Julia's ErrorException)."""

    def __init__(self, code, msg):
        super().__init__("genb200 error %d: %s" % (code, msg))
        self.code = code


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "libgenb200.so is not built (%s). Run `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `python gen.jl_b200/build.py`. There is no CPU fallback." % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            f = getattr(L, name)
            f.restype = res
            f.argtypes = args
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise GenB200Error(rc, lib().gb_last_error().decode())
