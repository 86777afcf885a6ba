/*
 * gen_oracle.h -- CPU restatement of Gen.jl's SMC / importance-sampling hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under gen.jl_b200/ may include, link or call
 * this; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs use it, and only as the checker / CPU arm.
 *
 * PARITY PINNING: Gen.jl is 100% Julia and there is no Julia in this image, so the
 * reference itself cannot run here.  This restatement is pinned against the
 * reference's own known-answer tests (see tests/test_oracle_golden.py):
 *   - test/inference/particle_filter.jl:1-81  (HMM forward alg, log ML -4.87645083351704)
 *   - test/modeling_library/distributions.jl:50,73,109 (-Inf / -5e25 KATs)
 *   - test/inference/importance_sampling.jl:18-34 (logsumexp(lognormw) ~ 0 @1e-14)
 *   - test/inference/trace_translators.jl:58-69 (custom-proposal weight algebra)
 *   - test/modeling_library/unfold.jl:41-70 (generate weight = sum of constrained logpdfs)
 * The resampler's *index stream* is "parity unpinned" in the reference (third-party
 * Distributions.jl alias sampler on Julia's global RNG; no test pins ancestors), so
 * the integer-CDF resamplers below are this repo's spec.
 *
 * All reference citations are relative to /root/reference/.
 */
#ifndef GEN_ORACLE_H
#define GEN_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* model kinds / proposal kinds / resampler kinds: numeric values match include/genb200.h */
enum { GO_MODEL_LGSSM = 1, GO_MODEL_SV = 2, GO_MODEL_BEARINGS = 3, GO_MODEL_HMM = 4, GO_MODEL_BLR = 5 };
enum { GO_PROPOSAL_DEFAULT = 0, GO_PROPOSAL_AFFINE_NORMAL = 1, GO_PROPOSAL_CATEGORICAL_TABLE = 2 };
enum { GO_RESAMPLE_MULTINOMIAL = 0, GO_RESAMPLE_SYSTEMATIC = 1, GO_RESAMPLE_RESIDUAL = 2, GO_RESAMPLE_MULTINOMIAL_SORTED = 3 };

/* ---- distributions (src/modeling_library/distributions/{normal,gamma,...}.jl) ---- */
double go_logpdf_normal(double x, double mu, double std);
double go_logpdf_gamma(double x, double shape, double scale);
double go_logpdf_bernoulli(int x, double p);
double go_logpdf_categorical(int64_t x, const double *probs, int64_t len);
double go_logpdf_mvnormal(const double *x, const double *mu, const double *cov, int k);
double go_random_normal(double mu, double std, double eps);
int go_random_bernoulli(double p, double u);
int64_t go_random_categorical(const double *probs, int64_t len, double u);
/* Marsaglia-Tsang gamma sampler driven by the engine's Philox stream (spec of this repo) */
double go_random_gamma_philox(double shape, double scale, uint64_t seed, uint64_t pid, uint32_t step);

/* ---- logsumexp / ESS (src/inference/inference.jl:3-11, particle_filter.jl:3-12) ---- */
double go_logsumexp(const double *a, int64_t n);
double go_logsumexp2(double x1, double x2);
double go_effective_sample_size(const double *lognormw, int64_t n);
double go_normalize_weights(const double *logw, int64_t n, double *lognormw_out);

/* ---- Philox4x32-10 noise spec shared with the engine (integer, bit exact) ---- */
void go_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
void go_philox_normals(uint64_t seed, uint32_t step, int64_t pid0, int64_t n, int ndraw, double *out /*[ndraw][n]*/);
void go_philox_uniforms(uint64_t seed, uint32_t step, int64_t pid0, int64_t n, int ndraw, double *out /*[ndraw][n]*/);
uint32_t go_philox_resample_u0(uint64_t seed, uint32_t step);
void go_philox_multinomial_u64(uint64_t seed, uint32_t step, int64_t k0, int64_t n, uint64_t *out);

/* ---- integer-CDF resamplers (this repo's spec) ---- */
double go_dexp(double x);                         /* deterministic exp for x <= 0 */
int go_quant_bits(int64_t n_global);
uint64_t go_quantize(double lw_minus_max, int b); /* floor(exp(x) * 2^b) via go_dexp */
/* returns Q_N; q_out may be NULL */
uint64_t go_quantized_weights(const double *logw, int64_t n, double m, int b, uint64_t *q_out);
/* ancestors are 0-based here; callers add 1 for Gen's `parents` */
int go_resample_systematic(const double *logw, int64_t n, uint32_t u0, int64_t *anc);
int go_resample_multinomial(const double *logw, int64_t n, const uint64_t *u, int64_t *anc);
int go_resample_multinomial_sorted(const double *logw, int64_t n, const uint64_t *e /* n + 1 spacings */, int64_t *anc);
int go_resample_residual(const double *logw, int64_t n, uint32_t u0, int64_t *anc);

/* ---- exact log-ML oracles ---- */
double go_kalman_logml(double m0, double s0, double a, double q, double r, const double *y, int T1);
double go_hmm_forward(const double *prior, const double *emission /*M x K col-major*/,
                      const double *transition /*K x K col-major*/, int K, int M,
                      const int64_t *obs, int T);
double go_blr_log_evidence(const double *X /*n x D row-major*/, const double *y, int n, int D, double sigma);

/* ---- particle filter / importance sampling restatement ---- */
typedef struct go_pf go_pf_t;
int go_model_state_dim(int kind, const double *params);
int go_model_num_normals(int kind, const double *params, int proposal_kind, int is_init);
int go_model_num_uniforms(int kind, const double *params, int proposal_kind, int is_init);
int go_model_num_obs(int kind, const double *params);

void go_set_threads(int nthreads);
/* GB_F32 storage model (this repo's extension): filters created after go_set_storage_f32(1) round every stored trace
 * field, log weight and log incremental weight to single precision; the arithmetic stays fp64. */
void go_set_storage_f32(int on);
int go_get_max_threads(void);

/* particle_filter.jl:118-154.  normals/uniforms: pre-drawn [ndraw][n] (NULL => Philox(seed)). */
go_pf_t *go_pf_init(int kind, const double *params, int nparams, int64_t n, uint64_t seed,
                    const double *obs, int proposal_kind, const double *pargs,
                    const double *normals, const double *uniforms);
/* particle_filter.jl:186-240 (+ trace_translators.jl:836-856); incr_out may be NULL */
int go_pf_step(go_pf_t *pf, const double *obs, int proposal_kind, const double *pargs,
               const double *normals, const double *uniforms, double *incr_out);
/* particle_filter.jl:249-275 with the integer-CDF resamplers; u64s only for multinomial validation */
int go_pf_maybe_resample(go_pf_t *pf, double ess_threshold, int kind, int use_given_u0, uint32_t u0,
                         const uint64_t *u64s, double *ess_out);
/* the resampling branch of maybe_resample! (:261-272) with GIVEN 1-based parents (validation: shared genealogy) */
int go_pf_resample_given(go_pf_t *pf, const int64_t *parents);
/* MH rejuvenation of every particle's latest latent choices: move 0 = mh(trace, selection) (mh.jl:16-27),
 * move 1 = mh(trace, gaussian-drift proposal, (drift_std,)) (mh.jl:37-58).  Returns #accepted, -1 if unsupported. */
int64_t go_pf_rejuvenate(go_pf_t *pf, int move, const double *obs, double drift_std, uint32_t move_index,
                         const double *normals, const double *uniforms, double *alpha_out, uint8_t *accepted_out);
/* sliding-window rejuvenation (docs/src/tutorials/smc.md:492-541): keep the last W states of every particle's lineage
 * (call right after go_pf_init), then one block drift move mh(trace, perturbation_proposal, (a, b)) per particle */
void go_pf_enable_window(go_pf_t *pf, int W);
int64_t go_pf_rejuvenate_window(go_pf_t *pf, int a, int b, const double *obs_window, double drift_std, uint32_t move_index,
                                const double *normals, const double *uniforms, double *alpha_out, uint8_t *accepted_out);
double go_pf_log_ml_estimate(const go_pf_t *pf);        /* particle_filter.jl:91-94 */
const double *go_pf_log_weights(const go_pf_t *pf);     /* :82-84 */
const int64_t *go_pf_parents(const go_pf_t *pf);        /* 1-based */
const double *go_pf_state(const go_pf_t *pf, int field);
double go_pf_log_ml_est_field(const go_pf_t *pf);
int64_t go_pf_num_particles(const go_pf_t *pf);
int go_pf_step_index(const go_pf_t *pf);
go_pf_t *go_pf_copy(const go_pf_t *pf);                 /* :43-51 */
void go_pf_set_log_weights(go_pf_t *pf, const double *lw);
void go_pf_free(go_pf_t *pf);
/* importance.jl:20-59: returns lml estimate, writes log-normalised weights; state kept in *pf_out if non-NULL */
double go_importance_sampling(int kind, const double *params, int nparams, int64_t n, uint64_t seed,
                              const double *obs, int proposal_kind, const double *pargs,
                              const double *normals, const double *uniforms,
                              double *lognormw_out, go_pf_t **pf_out);

#ifdef __cplusplus
}
#endif
#endif
