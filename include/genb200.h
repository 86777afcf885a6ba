/*
 * genb200.h -- C ABI of libgenb200.so: the B200 (sm_100a) particle inference engine that sits
 * behind Gen.jl's particle-filter / importance-sampling API.
 *
 * This is the `ccall` target of the Julia glue and of the Python ctypes host
 * mirror in gen.jl_b200/.  Plain pointers and sizes only; every entry returns 0 on success or a
 * non-zero GB_E* code with the message available from gb_last_error() (thread local), which
 * replaces Julia's `error(...)` (src/inference/particle_filter.jl:227-229).  There is NO CPU
 * fallback: every compute entry fails with GB_ENOGPU when no sm_100 device is usable.
 *
 * All file:line citations are relative to the reference checkout (probcomp/Gen.jl v0.4.8).
 * Unless a parameter is documented as a DEVICE pointer it is a HOST pointer owned by the caller;
 * the engine copies in/out.  Handles are not thread safe (one caller thread per handle).
 */
#ifndef GENB200_H
#define GENB200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

enum { GB_OK = 0, GB_EINVAL = 1, GB_ENOGPU = 2, GB_ECUDA = 3, GB_EUNSUPPORTED = 4, GB_ESTATE = 5 };

/* Supported model class: Unfold / Map / static-IR programs whose random
 * choices are built-in distributions.  `params` layouts:
 *   GB_MODEL_LGSSM    [m0, s0, a, q, r]            x0~normal(m0,s0); x_t~normal(a*x,q); y_t~normal(x_t,r)
 *   GB_MODEL_SV       [mu, phi, sigma, sd0]        h0~normal(mu,sd0); h_t~normal(mu+phi*(h-mu),sigma); y~normal(0,exp(h/2))
 *   GB_MODEL_BEARINGS [x0mu,x0sd,y0mu,y0sd,vx0mu,vx0sd,vy0mu,vy0sd, velocity_var, measurement_noise]
 *                     (docs/src/tutorials/smc.md:607-661)
 *   GB_MODEL_HMM      [K, M, prior[K], transition[K*K] (column = previous z), emission[M*K] (column = z)]
 *                     (test/inference/particle_filter.jl:52-78; categorical choices, 1-based)
 *   GB_MODEL_BLR      [D, n, sigma, X[n*D] row-major]   w~Map(normal)(0,1); y_j~normal(x_j.w, sigma)
 */
enum { GB_MODEL_LGSSM = 1, GB_MODEL_SV = 2, GB_MODEL_BEARINGS = 3, GB_MODEL_HMM = 4, GB_MODEL_BLR = 5, GB_MODEL_SPEC = 6 };
/* GB_MODEL_SPEC: a declarative specification of any model of the supported class -- a static-IR `generate` program for
 * t = 0 and the Unfold kernel's extend-by-one program, written out as SSA ops over at most 32 registers
 * (Gen's own DAG cannot be lowered automatically: its Julia nodes are opaque closures, src/static_ir/dag.jl:14-19).
 *   params = [S, n_params, n_obs, n_init_ops, n_step_ops, params[n_params], ops[(n_init+n_step) x (code, dst, a, b, imm)]]
 *   ops:  0 CONST dst<-imm   1 PREV dst<-previous state field a   2 PARAM dst<-params[a]   3 OBS dst<-obs[a]   4 TIME dst<-t
 *        10 ADD 11 SUB 12 MUL 13 DIV (dst<-a op b)  14 NEG 15 SQRT 16 EXP 17 LOG 19 ABS (dst<-f(a))  18 ATAN2 dst<-atan(a, b)
 *        20 SAMPLE_NORMAL dst ~ normal(a, b)          21 OBSERVE_NORMAL  weight += logpdf(normal, reg dst, a, b)
 *        22 SAMPLE_BERNOULLI dst ~ bernoulli(a)       23 OBSERVE_BERNOULLI weight += logpdf(bernoulli, reg dst, a)
 *        24 SAMPLE_GAMMA dst ~ gamma(a, b)            25 OBSERVE_GAMMA   weight += logpdf(gamma, reg dst, a, b)
 *        26 SAMPLE_CATEGORICAL dst ~ categorical(params[reg a : reg a + b])   (b = length, immediate)
 *        27 OBSERVE_CATEGORICAL weight += logpdf(categorical, reg dst, params[reg a : reg a + b])
 *        30 STORE state field dst <- reg a   (every field exactly once per program)
 *        31 WEIGHT weight += reg a              32 LOGPDF_NORMAL dst <- logpdf(normal, reg (int)imm, a, b)
 *        40 SAMPLE_DIST dst ~ D(a, b), D = GB_DIST_* (int)imm, for the built-ins with a closed-form quantile function
 *           (uniform, exponential, laplace, cauchy, geometric, uniform_discrete; one draw of the uniform stream)
 *        41 OBSERVE_DIST weight += logpdf(D, reg dst, a, b), D = GB_DIST_UNIFORM .. GB_DIST_UNIFORM_DISCRETE
 *        42 LOGPDF_DIST dst <- logpdf(D, reg x, a, b) with imm = D + 100 * x
 *        43 SAMPLE_MVNORMAL  regs dst .. dst+k-1 ~ mvnormal(mu = regs a .. a+k-1, cov = params[b : b + k*k] row-major), k = (int)imm
 *        44 OBSERVE_MVNORMAL weight += logpdf(mvnormal, x = regs dst .. dst+k-1, mu = regs a .. a+k-1, cov = params[b : b + k*k])
 *           (mvnormal.jl:12-33; 1 <= k <= 8.  The covariance is a model parameter shared by all particles, so it is factored ONCE, at
 *           gb_model_create -- the reference re-factors it at every logpdf / random call -- and the kernels use L and log det)
 * `sample` ops are unconstrained choices (weight unchanged: static_ir/generate.jl:31-38), `observe` ops constrained ones
 * (weight += logpdf: unfold/update.jl:64-68).  A custom proposal (particle_filter.jl:186-208, trace_translators.jl:854) is
 * written inside the program: sample from the proposal, OBSERVE the model's density of the proposed value, and subtract
 * the proposal's score (LOGPDF_NORMAL, NEG, WEIGHT). */
/* Custom proposals (particle_filter.jl:118-134, :186-208; trace_translators.jl:836-856):
 *   GB_PROPOSAL_AFFINE_NORMAL     pargs = [c0, c1, std]: x ~ normal(c0 + c1*x_prev, std)   (c1 unused at t=0)
 *   GB_PROPOSAL_CATEGORICAL_TABLE pargs = K probs at t=0, K*K (column = previous z) afterwards */
enum { GB_PROPOSAL_DEFAULT = 0, GB_PROPOSAL_AFFINE_NORMAL = 1, GB_PROPOSAL_CATEGORICAL_TABLE = 2 };
/* particle_filter.jl:261-262 is multinomial; systematic / residual are the north-star extensions
 * (exact integer-CDF definitions shared with the oracle). */
enum { GB_RESAMPLE_MULTINOMIAL = 0, GB_RESAMPLE_SYSTEMATIC = 1, GB_RESAMPLE_RESIDUAL = 2,
       /* multinomial resampling with the draws SORTED by ancestor (production mode; not in the reference): the order statistics
        * of N iid uniforms through exponential spacings, merged with the integer CDF in one streaming pass.  Same multiset
        * distribution as GB_RESAMPLE_MULTINOMIAL (particle_filter.jl:261-262), different slot order.  Single shard. */
       GB_RESAMPLE_MULTINOMIAL_SORTED = 3 };
enum { GB_F64 = 0, GB_F32 = 1 };
/* distributions for the batched logpdf/random entries (modeling_library/distributions/<name>.jl) */
enum { GB_DIST_NORMAL = 1, GB_DIST_GAMMA = 2, GB_DIST_BERNOULLI = 3, GB_DIST_CATEGORICAL = 4, GB_DIST_MVNORMAL = 5,
       /* the other scalar built-ins (batched entries only): (a, b) = */
       GB_DIST_UNIFORM = 6 /* low, high */, GB_DIST_EXPONENTIAL = 7 /* rate */, GB_DIST_LAPLACE = 8 /* loc, scale */,
       GB_DIST_CAUCHY = 9 /* x0, gamma */, GB_DIST_INV_GAMMA = 10 /* shape, scale */, GB_DIST_BETA = 11 /* alpha, beta */,
       GB_DIST_POISSON = 12 /* lambda */, GB_DIST_GEOMETRIC = 13 /* p */, GB_DIST_BINOM = 14 /* n, p */,
       GB_DIST_NEG_BINOM = 15 /* r, p */, GB_DIST_UNIFORM_DISCRETE = 16 /* low, high */,
       /* normal.jl:61-120: x, mu, std broadcast against each other (a scalar or n values each); logpdf is ONE number, the sum */
       GB_DIST_BROADCASTED_NORMAL = 17,
       GB_DIST_DIRICHLET = 18 /* a = alpha[len]; x is [n][len] row-major (dirichlet.jl:12-21, :33-35) */,
       GB_DIST_PIECEWISE_UNIFORM = 19 /* a = bounds[len + 1], b = probs[len] (piecewise_uniform.jl:31-52) */,
       GB_DIST_BETA_UNIFORM = 20 /* a = theta, b = (alpha, beta) pairs: 2 values, or 2 n (beta_uniform.jl:10-18, :36-42) */ };

typedef struct gb_model gb_model_t;
typedef struct gb_pf gb_pf_t;

const char *gb_last_error(void);
const char *gb_version(void);
int gb_device_count(int *count);              /* 0 devices => *count = 0, returns GB_OK */
/* Device and pinned-host buffers released by gb_pf_free / gb_model_free are kept on a per-device free list and reused by
 * the next filter (a raw cudaMalloc/cudaFree cycle costs more than running a small filter).  GENB200_CACHE_MB caps the
 * cached bytes (default 8192; 0 = no caching).  gb_cache_trim returns the cached blocks to the driver. */
int gb_cache_trim(void);
int gb_cache_stats(int64_t *cached_bytes, int64_t *cached_blocks, int64_t *live_blocks);
int gb_set_device(int device);

/* ---- model handles (replace GenerativeFunction objects of the supported class) ---- */
int gb_model_create(int kind, const double *params, int nparams, gb_model_t **out);
void gb_model_free(gb_model_t *m);
int gb_model_state_dim(const gb_model_t *m);
int gb_model_num_obs(const gb_model_t *m);
int gb_model_num_normals(const gb_model_t *m, int proposal_kind, int is_init);
int gb_model_num_uniforms(const gb_model_t *m, int proposal_kind, int is_init);

/* ---- ParticleFilterState (particle_filter.jl:35-41) as a device-resident SoA store ---- */
/* Allocate a shard of n_local particles with global ids [pid_offset, pid_offset+n_local) of an
 * n_global-particle filter (single GPU: n_local == n_global, pid_offset == 0). */
int gb_pf_create(const gb_model_t *m, int64_t n_local, int64_t n_global, int64_t pid_offset, int dtype,
                 uint64_t seed, gb_pf_t **out);
/* Validation mode: pre-drawn noise [nd][n_local] consumed by the NEXT generate/step call in place
 * of the Philox stream (Gen has no RNG hook).  NULL / 0 clears. */
int gb_pf_set_noise(gb_pf_t *pf, const double *normals, int n_normals, const double *uniforms, int n_uniforms);
/* initialize_particle_filter (particle_filter.jl:118-134, :142-154): generate at t = 0. */
int gb_pf_generate(gb_pf_t *pf, const double *obs, int nobs, int proposal_kind, const double *pargs, int npargs);
/* create + generate in one call (the usual entry) */
int gb_pf_initialize(const gb_model_t *m, int64_t n, int dtype, uint64_t seed, const double *obs, int nobs,
                     int proposal_kind, const double *pargs, int npargs, gb_pf_t **out);
/* particle_filter_step! (particle_filter.jl:186-208, :217-240), extend-by-one only (the reference's
 * non-empty-discard error, :227-229, is unreachable by construction).  incr_out (n_local doubles,
 * may be NULL) receives the log incremental weights the reference returns. */
int gb_pf_step(gb_pf_t *pf, const double *obs, int nobs, int proposal_kind, const double *pargs, int npargs,
               double *incr_out);
/* maybe_resample! (particle_filter.jl:249-275).  Single shard only (n_local == n_global).
 * Resampling randomness: Philox(seed, step) unless gb_pf_set_resample_noise was called. */
int gb_pf_maybe_resample(gb_pf_t *pf, double ess_threshold, int kind, int *did_resample);
/* Optional trace history (off by default: the throughput configurations keep only the current state).  With history on
 * (enable BEFORE gb_pf_generate) every generate/step keeps a snapshot of the trace columns and the ancestors of the
 * resample that preceded it, so that gb_pf_get_trajectory can return, for every CURRENT particle i and every time
 * t = 0..step, the value the particle's own lineage had: out[t * n + i] -- Gen's traces hold all choices at all times
 * (get_traces, particle_filter.jl:70-73; `trace[:chain => t => :z]`). */
int gb_pf_enable_history(gb_pf_t *pf, int on);
int gb_pf_history_length(const gb_pf_t *pf, int *steps);
int gb_pf_get_trajectory(gb_pf_t *pf, int field, double *out);
/* sample_unweighted_traces (particle_filter.jl:101-109): n_samples iid draws from Categorical(normalised weights).
 * idx_out: n_samples 1-based particle indices; rows_out (may be NULL): S columns x n_samples values of the sampled
 * traces.  `draw` distinguishes successive calls on the same state (Philox stream (seed, draw)). */
int gb_pf_sample_unweighted(gb_pf_t *pf, int64_t n_samples, uint64_t draw, int64_t *idx_out, double *rows_out);
/* MH rejuvenation of EVERY particle: the tutorial loop
 *   for i in 1:N  state.traces[i], _ = mh(state.traces[i], ...)  end      (docs/src/tutorials/smc.md:459-463, :518-520)
 * for moves on the latent choices of the LATEST kernel application (the only ones whose parent state the Markov store
 * still holds).  GB_MH_RESIMULATE = mh(trace, select(<latest latents>)) (src/inference/mh.jl:16-27; weight =
 * regenerate's, src/dynamic/regenerate.jl:36-50); GB_MH_DRIFT = mh(trace, proposal, args) with c' ~ normal(c, drift_std)
 * on every latest latent choice (mh.jl:37-58).  Accept iff log(rand()) < alpha.  Weights are untouched.  Valid after
 * gb_pf_generate / gb_pf_step and before gb_pf_maybe_resample (GB_ESTATE otherwise); obs = the latest observations.
 * n_accepted, alpha_out (n log acceptance ratios) may be NULL.  Validation mode (gb_pf_set_noise): normals = the
 * transition's rows (drift: one per latent choice), uniforms = the transition's rows followed by ONE accept row.
 * Models: LGSSM, SV, BEARINGS, HMM (resimulate only), SPEC (resimulate only; needs the NVRTC-specialised kernels and a program
 * without inline proposal terms -- WEIGHT / LOGPDF ops). */
#define GB_MH_RESIMULATE 0
#define GB_MH_DRIFT 1
int gb_pf_rejuvenate(gb_pf_t *pf, int move, const double *obs, int nobs, double drift_std, int64_t *n_accepted,
                     double *alpha_out);
/* Sliding-window rejuvenation (docs/src/tutorials/smc.md:492-541): ONE block MH move per particle,
 *     mh(trace, perturbation_proposal, (a, b)),  {latent choices of step t} ~ normal(current value, drift_std) for a <= t <= b,
 * accepted with the update weight of src/dynamic/update.jl:46-62 -- the latent choices of steps a..b+1 under their (moved)
 * parents and the observations of ALL steps a..T -- minus the forward and plus the backward proposal score (mh.jl:37-58).
 * gb_pf_enable_window(W) (before gb_pf_generate; 1 <= W <= 8; single shard; not with gb_pf_enable_history) makes every
 * particle carry the last W states of its own lineage (window columns gathered with it at every resample); a move may
 * reach back to step a >= T - W + 1.  obs_window: the observations of steps a..T (T - a + 1 scalars).  Call between
 * particle_filter_step! and maybe_resample!, like gb_pf_rejuvenate (which it disables until the next step).  Models: LGSSM,
 * SV, BEARINGS.  Validation noise (gb_pf_set_noise): (b - a + 1) x (choices per step) normal rows, step-major, and ONE
 * uniform row (the accept draw).  The fused gb_pf_run* entries do not maintain the window. */
int gb_pf_enable_window(gb_pf_t *pf, int W);
int gb_pf_rejuvenate_window(gb_pf_t *pf, int a, int b, const double *obs_window, int nobs, double drift_std,
                            int64_t *n_accepted, double *alpha_out);
/* Run-time specialisation of GB_MODEL_SPEC models.  Gen turns a static-IR function into straight-line code per trace
 * type (`@generated` functions, src/static_ir/generate.jl:68-109, update.jl:467-513); here the program that crossed
 * this boundary is lowered to CUDA C++ (one statement per node, parameters as literals) and compiled for sm_100a
 * with NVRTC on first use, per (model, dtype, device).  If libnvrtc / libcuda cannot be opened, or with
 * GENB200_SPEC_JIT=0 in the environment, the generic interpreter kernel runs the same program instead (same device
 * functions: identical states, weights equal to rounding of constant-folded logs).
 * GENB200_JIT_CACHE=<dir> keeps the compiled programs on disk (file name = hash of source + headers + dtype).
 *   gb_spec_jit_source   the generated translation unit (buf may be NULL to query *needed)
 *   gb_spec_jit_compile  compile only (no GPU needed); GB_EUNSUPPORTED + the NVRTC log in gb_last_error() on failure
 *   gb_pf_spec_backend   which kernel the filter's last generate/step ran: 1 specialised, 0 interpreter, -1 n/a */
int gb_spec_jit_source(const gb_model_t *m, char *buf, int64_t cap, int64_t *needed);
int gb_spec_jit_compile(const gb_model_t *m, int dtype, int64_t *cubin_bytes);
int gb_pf_spec_backend(const gb_pf_t *pf, int *backend);
/* Throughput entry: T iterations of { maybe_resample!(ess_threshold, systematic); particle_filter_step!(obs[t]) }
 * (default proposal) with the ESS test (particle_filter.jl:256), the log_ml_est update (:263) and the
 * gathered/direct choice of the step made ON THE DEVICE, i.e. without the host round trip the Bool return of
 * maybe_resample! forces per step.  Bit-identical to calling the two entries T times.  obs: T * nobs doubles. */
int gb_pf_run(gb_pf_t *pf, int T, const double *obs, int nobs, double ess_threshold, int *n_resampled);
/* The same loop with `sweeps` MH rejuvenation moves (gb_pf_rejuvenate's, GB_MH_*) on the traces each step just extended
 * -- the tutorial's particle_filter_rejuv loop (docs/src/tutorials/smc.md:513-531) in one call; identical to calling
 * gb_pf_maybe_resample / gb_pf_step / gb_pf_rejuvenate x sweeps per step.  n_accepted: total accepted moves. */
int gb_pf_run_mh(gb_pf_t *pf, int T, const double *obs, int nobs, double ess_threshold, int move, double drift_std, int sweeps,
                 int *n_resampled, int64_t *n_accepted);
/* validation mode for the resampler: u0 for systematic/residual, n u64 for multinomial (NULL clears) */
int gb_pf_set_resample_noise(gb_pf_t *pf, int use_u0, uint32_t u0, const uint64_t *u64s);
/* validation mode of GB_RESAMPLE_MULTINOMIAL_SORTED: n + 1 integer spacings (each >= 1, sum < 2^63) in place of the Philox
 * ones, consumed by the next resample (NULL clears).  gb_resample / gb_resample_device take them through their u64 argument. */
int gb_pf_set_resample_spacings(gb_pf_t *pf, const uint64_t *spacings);
int gb_pf_log_ml_estimate(gb_pf_t *pf, double *out);                 /* particle_filter.jl:91-94 */
int gb_pf_get_log_weights(gb_pf_t *pf, double *out_n);               /* :82-84 (unnormalised) */
/* The return value of the last particle_filter_step! (:199-207, :231-239), read on demand: gb_pf_step(..., NULL) leaves
 * the N increments on the device and the glue returns a lazy vector that calls this on first access. GB_ESTATE when
 * the filter has moved on (or has not stepped yet). */
int gb_pf_get_log_increments(gb_pf_t *pf, double *out_n);
int gb_pf_set_log_weights(gb_pf_t *pf, const double *in_n);          /* tests poke .log_weights (:176-197) */
int gb_pf_get_parents(gb_pf_t *pf, int64_t *out_n);                  /* :40, 1-based, global ids */
int gb_pf_get_state(gb_pf_t *pf, int field, double *out_n);          /* column `field` of the traces */
int gb_pf_get_log_ml_est(gb_pf_t *pf, double *out);                  /* the .log_ml_est field (:39) */
int gb_pf_get_ess(gb_pf_t *pf, double *ess_out, double *lse_out);    /* particle_filter.jl:3-12 */
int gb_pf_num_particles(const gb_pf_t *pf, int64_t *n_local, int64_t *n_global);
int gb_pf_step_index(const gb_pf_t *pf, int *step);
int gb_pf_clone(gb_pf_t *pf, gb_pf_t **out);                         /* Base.copy (:43-51) */
void gb_pf_free(gb_pf_t *pf);
/* 0 = eager gather inside maybe_resample!, 1 (default) = the resampled-trace gather is deferred and
 * fused into the next particle_filter_step! (materialised on any read of the state). */
int gb_pf_set_lazy_gather(gb_pf_t *pf, int lazy);
/* All launches of this handle go to `cuda_stream` (a cudaStream_t; NULL = legacy default stream). */
int gb_pf_set_stream(gb_pf_t *pf, void *cuda_stream);
int gb_pf_sync(gb_pf_t *pf);

/* ---- sharded (one process per GPU) pieces of maybe_resample!; the host glue runs the collectives
 *      between them (gen.jl_b200/sharded.py).  All scalars are host values. ---- */
/* max of the shard's log weights (NaN propagates like Julia's maximum; all-zero weights: 0.0) */
int gb_pf_weight_max(gb_pf_t *pf, double *m_local);
/* One pass over the shard's log weights w.r.t. the GLOBAL max m_ref (after the host's all-reduce):
 * s1 = sum exp(lw - m_ref), s2 = sum exp(2 (lw - m_ref)) (inference.jl:3-6, particle_filter.jl:3-12) and
 * q_local = sum floor(exp(lw - m_ref) * 2^bits), the shard's integer CDF mass; bits = gb_quant_bits(n_global). */
int gb_pf_weight_sums(gb_pf_t *pf, double m_ref, int bits, double *s1, double *s2, uint64_t *q_local);
/* Device-resident flavour of the two calls above, so that the host can run the collectives directly on
 * device memory (no host round trip): gb_pf_stats_device returns the DEVICE address of the statistics
 * record {m, s1, s2, lse, ess, q_total} (n_words 8-byte words); gb_pf_weight_max_device makes its m current;
 * the host all-reduces m in place (max); gb_pf_weight_sums_device fills s1, s2, q_total w.r.t. that m. */
int gb_pf_stats_device(gb_pf_t *pf, void **dev_stats, int *n_words);
int gb_pf_weight_max_device(gb_pf_t *pf);
int gb_pf_weight_sums_device(gb_pf_t *pf, int bits);
/* Sharded maybe_resample! WITHOUT collectives and without a host round trip: everything the shards tell each
 * other travels through peer memory over NVLink / NVSwitch, written and awaited by the kernels themselves.
 *   - every shard owns a small mailbox; the last CTA of the kernel that produced the weights peer-stores the shard's max
 *     into EVERY mailbox (phase A), the statistics kernel waits for the G maxima, reduces w.r.t. the global max and
 *     peer-stores (sum e, sum e^2, integer CDF mass) (phase B); a one-CTA kernel waits for the G records and computes the
 *     ESS test (particle_filter.jl:256), the log-ML update (:263) and the exchange plan -- the same arithmetic on the
 *     same records on every shard, so all agree without talking;
 *   - offspring owned by another shard are written by a kernel STRAIGHT INTO THAT SHARD'S MEMORY: rows into the
 *     extension rows of its current trace buffer, plus the slot's ancestor index and global parent id; its last CTA
 *     posts a "done" flag (phase C) that the peers' next step kernel waits for before it reads the rows.
 *   once   : gb_pf_ipc_export -> exchange of the bundles (any transport) -> gb_pf_async_attach_ipc for every peer in
 *            another process / gb_pf_async_attach_local for peers in this process -> gb_pf_async_setup -> a barrier
 *            across all shards (the set-up resets the mailbox)
 *   per run: gb_pf_async_arm(threshold), then per step gb_pf_async_iterate(obs) = gb_pf_xchg_stats_plan (or gb_pf_xchg_stats ->
 *            gb_pf_async_plan) -> gb_pf_async_sweep -> gb_pf_async_step: four kernels per shard and step, enqueues only.
 *            (gb_pf_async_export / gb_pf_async_wait are no-ops kept for older drivers: the export rides in the sweep's second
 *            kernel, the wait sits at the top of the step kernel.)
 *   finish : gb_pf_async_finish, one device->host read (log_ml_est, #resamples, overflow / time-out flags).  Overflow (a
 *            shard's extension rows or spill buffer too small for the weight skew) or a peer that did not answer within
 *            GENB200_XCHG_TIMEOUT_MS (default 20000) invalidates the run: GB_ESTATE.
 * Synchronous maybe_resample! (the Bool return needs one host read): gb_pf_xchg_stats -> gb_pf_async_plan(u0, -inf, 0) ->
 * gb_pf_xchg_read gives the global lse / ESS; the caller decides; gb_pf_async_plan(u0, +inf, 1) -> sweep -> export -> wait ->
 * gb_pf_async_commit.  Several shards on ONE device (tests): give them one stream and enqueue phase by phase over the shards. */
int gb_pf_ipc_bundle_bytes(void);
int gb_pf_ipc_export(gb_pf_t *pf, void *bundle_out, int *nbytes);
int gb_pf_async_attach_ipc(gb_pf_t *pf, int peer_rank, const void *bundle);
int gb_pf_async_attach_local(gb_pf_t *pf, int peer_rank, gb_pf_t *peer);
int gb_pf_async_setup(gb_pf_t *pf, int world, int rank, const int64_t *offs, double ess_threshold);
int gb_pf_async_arm(gb_pf_t *pf, double ess_threshold);
int gb_pf_async_cur(const gb_pf_t *pf, int *cur);
int gb_pf_xchg_post(gb_pf_t *pf);    /* phase A only (no-op when the producing kernel already posted); implied by gb_pf_xchg_stats */
int gb_pf_xchg_stats(gb_pf_t *pf);
/* gb_pf_xchg_stats + gb_pf_async_plan(u0, NaN, 1) fused into one kernel (the plan rides in the statistics kernel's tail);
 * only when no peer shard shares this shard's device */
int gb_pf_xchg_stats_plan(gb_pf_t *pf, uint32_t u0);
/* ess_threshold NaN: the armed one; do_scan != 0 when a sweep follows */
int gb_pf_async_plan(gb_pf_t *pf, uint32_t u0, double ess_threshold, int do_scan);
int gb_pf_async_sweep(gb_pf_t *pf);
int gb_pf_async_export(gb_pf_t *pf);
int gb_pf_async_wait(gb_pf_t *pf);
int gb_pf_async_step(gb_pf_t *pf, const double *obs, int nobs);
int gb_pf_async_iterate(gb_pf_t *pf, const double *obs, int nobs);
int gb_pf_async_finish(gb_pf_t *pf, int *n_resampled, int *overflow);
/* global statistics over ALL shards and the decision of the last gb_pf_async_plan (one device->host read) */
int gb_pf_xchg_read(gb_pf_t *pf, double *lse, double *ess, uint64_t *q_total, int *do_resample);
/* (no device access) the global max and every shard's integer CDF mass (world entries) as of the last gb_pf_xchg_read */
int gb_pf_xchg_records(gb_pf_t *pf, double *m_global, uint64_t *q_shard);
/* the same read + the host-side bookkeeping of a resample the DEVICE decided (gb_pf_async_plan with a finite threshold) */
int gb_pf_xchg_commit(gb_pf_t *pf, int *did_resample, double *ess);
/* Offspring of this shard's particles under global-prefix systematic resampling: the shard's particles
 * (integer CDF mass q_local, preceded by q_offset of a global q_total) fill the global slot range
 * [*slot_lo, *slot_hi).  Ancestors of the slots this shard owns itself go straight into its ancestor
 * column, the others are kept for gb_pf_export_offspring. */
int gb_pf_systematic_offspring(gb_pf_t *pf, double m_global, int bits, uint64_t q_offset, uint64_t q_local,
                               uint64_t q_total, uint32_t u0, int64_t *slot_lo, int64_t *slot_hi);
/* Exchange block of `count` offspring rows (DEVICE memory, gb_pf_block_bytes(count) bytes, 16-byte aligned):
 * [parents: count x int64, global 1-based][rows: S columns x count elements of the filter dtype]. */
int64_t gb_pf_block_bytes(const gb_pf_t *pf, int64_t count);
/* Fast path of the exchange (used when every shard receives at most gb_pf_ext_capacity rows): slots
 * [k_lo, k_hi) that stay on this shard just keep their ancestors (the gather is deferred and fused into
 * the next particle_filter_step!), rows received from other shards are parked in extension rows of the
 * current trace buffer (ext_offset = running count of received rows) and their slots point there;
 * gb_pf_commit_resample_lazy then finishes maybe_resample! (particle_filter.jl:263-272). */
int gb_pf_ext_capacity(const gb_pf_t *pf, int64_t *cap);
int gb_pf_offspring_local(gb_pf_t *pf, int64_t k_lo, int64_t k_hi);
int gb_pf_import_ext(gb_pf_t *pf, int64_t dst_lo, int64_t count, int64_t ext_offset, const void *dev_block);
int gb_pf_commit_resample_lazy(gb_pf_t *pf, double log_ml_increment);
/* The same bookkeeping after an exchange done by the device-planned peer stores (gb_pf_async_plan / _sweep / _export and
 * the caller's barrier) when the CALLER took the resampling decision: the synchronous maybe_resample! over NVLink peers. */
int gb_pf_async_commit(gb_pf_t *pf, double log_ml_increment);
/* Staged exchange (arbitrarily skewed weights).  gb_pf_export_offspring gathers the rows of global slots
 * [k_lo, k_hi) (inside the shard's offspring range, not straddling its own slot range) into a block;
 * gb_pf_import_rows installs a block at local slots [dst_lo, dst_lo+count) of the NEW trace buffer;
 * gb_pf_commit_resample then swaps buffers, zeroes the log weights and adds log_ml_increment to
 * log_ml_est (particle_filter.jl:263-272). */
int gb_pf_export_offspring(gb_pf_t *pf, int64_t k_lo, int64_t k_hi, void *dev_block);
int gb_pf_import_rows(gb_pf_t *pf, int64_t dst_lo, int64_t count, const void *dev_block);
int gb_pf_commit_resample(gb_pf_t *pf, double log_ml_increment);

/* ---- ONE filter sharded over several GPUs, driven from the one calling thread ----
 * gb_spf_* mirror the gb_pf_* entries above over G shards with contiguous global particle ids (the first n % G shards hold
 * one extra particle).  devices: G device indices (NULL: shard r on device r; indices may repeat -- several shards on one
 * GPU).  Results are bit-identical to the single-GPU filter for every G: Philox noise is keyed by the global particle id
 * and maybe_resample! is global-prefix systematic resampling over the exact integer CDF.  No NCCL, no MPI: the shards
 * exchange through peer memory (see "Sharded maybe_resample! WITHOUT collectives" above); needs peer access between the
 * devices (NVLink / NVSwitch or PCIe P2P).  Resampler: systematic. */
typedef struct gb_spf gb_spf_t;
int gb_spf_create(const gb_model_t *m, int64_t n_global, int dtype, uint64_t seed, int nshards, const int *devices, gb_spf_t **out);
int gb_spf_generate(gb_spf_t *spf, const double *obs, int nobs, int proposal_kind, const double *pargs, int npargs);
int gb_spf_initialize(const gb_model_t *m, int64_t n, int dtype, uint64_t seed, int nshards, const int *devices, const double *obs,
                      int nobs, int proposal_kind, const double *pargs, int npargs, gb_spf_t **out);
int gb_spf_step(gb_spf_t *spf, const double *obs, int nobs, int proposal_kind, const double *pargs, int npargs);
int gb_spf_maybe_resample(gb_spf_t *spf, double ess_threshold, int *did_resample);
int gb_spf_set_resample_noise(gb_spf_t *spf, int use_u0, uint32_t u0);
int gb_spf_run(gb_spf_t *spf, int T, const double *obs, int nobs, double ess_threshold, int *n_resampled);
int gb_spf_log_ml_estimate(gb_spf_t *spf, double *out);
int gb_spf_get_log_ml_est(gb_spf_t *spf, double *out);
int gb_spf_get_ess(gb_spf_t *spf, double *ess_out, double *lse_out);
int gb_spf_get_log_weights(gb_spf_t *spf, double *out_n);
int gb_spf_get_log_increments(gb_spf_t *spf, double *out_n);
int gb_spf_get_state(gb_spf_t *spf, int field, double *out_n);
int gb_spf_get_parents(gb_spf_t *spf, int64_t *out_n);
int gb_spf_num_shards(const gb_spf_t *spf, int *nshards);
int gb_spf_shard(gb_spf_t *spf, int r, gb_pf_t **shard /* borrowed */, int64_t *pid_offset);
int gb_spf_num_particles(const gb_spf_t *spf, int64_t *n);
int gb_spf_step_index(const gb_spf_t *spf, int *step);
int gb_spf_sync(gb_spf_t *spf);
void gb_spf_free(gb_spf_t *spf);
/* importance_sampling (importance.jl:20-59) with the samples sharded: only the log-sum-exp crosses devices (cfg 4) */
int gb_spf_importance_sampling(const gb_model_t *m, int64_t n, int dtype, uint64_t seed, int nshards, const int *devices,
                               const double *obs, int nobs, int proposal_kind, const double *pargs, int npargs,
                               double *lognormw_out, double *lml_out, gb_spf_t **spf_out);
/* global-prefix systematic resampling of a raw log-weight array sharded over the devices (cfg 5); ancestors 0-based */
int gb_spf_resample(const double *logw, int64_t n, int dtype, int nshards, const int *devices, uint32_t u0, int64_t *anc_out);

/* ---- importance_sampling (importance.jl:20-59) ---- */
/* lognormw_out: n doubles (may be NULL); lml_out: log ML estimate; pf_out (may be NULL) receives the
 * particle store (traces + unnormalised log weights).  normals/uniforms: validation noise or NULL. */
int gb_importance_sampling(const gb_model_t *m, int64_t n, int dtype, uint64_t seed, const double *obs, int nobs,
                           int proposal_kind, const double *pargs, int npargs, const double *normals,
                           int n_normals, const double *uniforms, int n_uniforms, double *lognormw_out,
                           double *lml_out, gb_pf_t **pf_out);

/* ---- resampling on raw log-weight arrays (config 5 microbenchmark; same kernels as maybe_resample!) ----
 * Host-buffer form: ancestors are 0-based int64. u64s only for multinomial (NULL => Philox(seed, step)). */
int gb_resample(const double *logw, int64_t n, int dtype, int kind, uint32_t u0, const uint64_t *u64s,
                uint64_t seed, uint32_t step, int64_t *anc_out);
/* Device-resident form for benchmarking: dev_logw (n elements of dtype), dev_anc (n int32), dev_u64
 * (n uint64 or NULL).  Workspace is cached per thread.  Asynchronous on cuda_stream. */
int gb_resample_device(const void *dev_logw, int64_t n, int dtype, int kind, uint32_t u0, const void *dev_u64,
                       uint64_t seed, uint32_t step, void *dev_anc, void *cuda_stream);
/* logsumexp / ESS of a raw array (inference.jl:3-6, particle_filter.jl:3-12) */
int gb_logsumexp(const double *a, int64_t n, int dtype, double *lse_out, double *ess_out);

/* ---- batched distribution kernels (modeling_library/distributions/{normal,gamma,bernoulli,
 *      categorical,mvnormal}.jl logpdf + random) ----
 * x, a, b: n elements each (a/b broadcast when a_stride/b_stride == 0).  categorical: a = probs[len],
 * x integer-valued; mvnormal: x is [n][k] row-major, a = mu[k], b = cov[k*k], k = len. */
int gb_dist_logpdf(int dist, int dtype, int64_t n, const double *x, const double *a, int a_stride, const double *b,
                   int b_stride, int len, double *out);
/* draws n values with the engine's Philox stream (seed, step, pid = 0..n-1) */
int gb_dist_random(int dist, int dtype, int64_t n, const double *a, int a_stride, const double *b, int b_stride,
                   int len, uint64_t seed, uint32_t step, double *out);

/* ---- importance_resampling (importance.jl:77-122) and the tail of enumerative_inference (enumerative.jl:39-46) ---- */
/* Streaming form, O(chunk) memory for any n: the reference keeps ONE trace while it runs over the n proposals, folding each
 * log weight into a running logsumexp(x1, x2) (inference.jl:8-11) and swapping the kept trace with probability
 * exp(w - running total) (importance.jl:95-103, :113-120).  Here the proposals are generated chunk by chunk (chunk <= 2^22
 * particles of device memory), each chunk is reduced on the device (log-sum-exp + one categorical draw from its normalised
 * weights), and the chunks are folded on the host by the same pairwise rule -- the same distribution over the kept trace.
 * row_out: S doubles, the kept trace; lml_out: logsumexp of all n log weights - log n (:104, :121). */
int gb_importance_resampling(const gb_model_t *m, int64_t n, int dtype, uint64_t seed, const double *obs, int nobs,
                             int proposal_kind, const double *pargs, int npargs, int64_t chunk, double *row_out, double *lml_out);
/* log_total = logsumexp(log_weights [+ log_vols]); lognormw_out = log_weights [+ log_vols] - log_total (may be NULL) */
int gb_enumerative_normalize(const double *log_weights, const double *log_vols, int64_t n, int dtype, double *lognormw_out,
                             double *log_total_out);

/* ---- the engine's counter-based noise, exported so the oracle can consume identical draws ---- */
int gb_philox_normals(uint64_t seed, uint32_t step, int64_t pid0, int64_t n, int ndraw, int dtype, double *out);
int gb_philox_uniforms(uint64_t seed, uint32_t step, int64_t pid0, int64_t n, int ndraw, int dtype, double *out);
uint32_t gb_philox_resample_u0(uint64_t seed, uint32_t step);         /* host-side integer, no GPU */
int gb_philox_multinomial_u64(uint64_t seed, uint32_t step, int64_t k0, int64_t n, uint64_t *out);
int gb_philox_sorted_spacings(uint64_t seed, uint32_t step, int64_t k0, int64_t n, uint64_t *out);   /* needs a GPU */
int gb_quant_bits(int64_t n_global);                                   /* host-side, no GPU */

/* ---- instrumentation for bench.py: per-kernel CUDA-event timings on the handle's stream ---- */
enum { GB_K_STEP = 0, GB_K_FINALIZE = 1 /* weight statistics */, GB_K_QTOTAL = 2, GB_K_TILESCAN = 3, GB_K_ANCESTORS = 4, GB_K_GATHER = 5,
       GB_K_MISC = 6, GB_K_COUNT = 7 };
int gb_pf_enable_timing(gb_pf_t *pf, int on);
/* ms[GB_K_COUNT], launches[GB_K_COUNT]; resets the accumulators */
int gb_pf_get_timing(gb_pf_t *pf, double *ms, int64_t *launches);
const char *gb_kernel_name(int k);

/* ---- internal test hooks: HOST-side evaluation of the exact-integer helpers the kernels use, so the
 *      CPU-only test suite can check them against Python big integers (not part of the product path) ---- */
int64_t gb_selftest_slots_below(uint64_t C, int64_t nslots, uint64_t q_total, uint32_t u0);
int64_t gb_selftest_slots_below_residual(uint64_t chi, uint64_t clo, uint64_t q_total, uint32_t u0);
uint64_t gb_selftest_div128(uint64_t ahi, uint64_t alo, uint64_t d);
uint64_t gb_selftest_quantize(double x, int bits);
/* fn: 0 log(a), 1 sin(2 pi a), 2 cos(2 pi a), 3 atan2(a, b), 4 exp(a), 5 u52(hi = a, lo = b): the engine's own
 * fp64 elementary functions (csrc/gbmath.cuh), host instance (bit-identical to the device one) */
double gb_selftest_math(int fn, double a, double b);
void gb_selftest_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);

#ifdef __cplusplus
}
#endif
#endif
