/*
 * gen_oracle.c -- CPU (fp64, sequential semantics) restatement of Gen.jl's particle
 * filter / importance sampling path.  TEST INFRASTRUCTURE ONLY (see gen_oracle.h).
 *
 * Build: gcc -O2 -ffp-contract=off -fopenmp -fPIC -shared (oracle/build.py).
 * -ffp-contract=off: Julia never fuses `mu + std*randn()` (normal.jl:128), neither do we.
 *
 * Every RNG call site of the reference (Base randn()/rand(), Distributions.jl samplers)
 * is replaced by a read of a pre-drawn array, or by this repo's Philox4x32-10 stream
 * (Gen has no RNG hook: SURVEY.md section 2a).  Every deterministic formula follows the
 * cited Julia.  Citations are relative to /root/reference/.
 */
#include "gen_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef unsigned __int128 u128;

static int g_threads = 1;
void go_set_threads(int n) {
  g_threads = n < 1 ? 1 : n;
#ifdef _OPENMP
  omp_set_num_threads(g_threads);
#endif
}
int go_get_max_threads(void) {
#ifdef _OPENMP
  return omp_get_num_procs();
#else
  return 1;
#endif
}

/* ------------------------------------------------------------------ */
/* Distributions                                                       */
/* ------------------------------------------------------------------ */

/* normal.jl:56-59:  z = (x - mu) / std;  -(abs2(z) + log(2pi))/2 - log(std) */
double go_logpdf_normal(double x, double mu, double std) {
  double z = (x - mu) / std;
  return -(z * z + log(2.0 * M_PI)) / 2.0 - log(std);
}
/* normal.jl:128:  mu + std * randn() */
double go_random_normal(double mu, double std, double eps) { return mu + std * eps; }

/* gamma.jl:10-16 */
double go_logpdf_gamma(double x, double shape, double scale) {
  if (x > 0.0) return (shape - 1.0) * log(x) - (x / scale) - shape * log(scale) - lgamma(shape);
  return -INFINITY;
}
/* bernoulli.jl:10-12 */
double go_logpdf_bernoulli(int x, double p) { return x ? log(p) : log(1.0 - p); }
/* bernoulli.jl:19:  rand() < prob */
int go_random_bernoulli(double p, double u) { return u < p; }
/* categorical.jl:10-12 (x is 1-based) */
double go_logpdf_categorical(int64_t x, const double *probs, int64_t len) {
  return (x > 0 && x <= len) ? log(probs[x - 1]) : -INFINITY;
}
/* categorical.jl:20-22 -> rand(Distributions.Categorical(probs)): single draw is an
 * inverse-CDF linear scan (third-party, unpinned); restated with a pre-drawn uniform. */
int64_t go_random_categorical(const double *probs, int64_t len, double u) {
  double c = 0.0;
  for (int64_t i = 0; i < len; i++) {
    c += probs[i];
    if (u < c) return i + 1;
  }
  return len;
}
/* mvnormal.jl:12-16 -> Distributions.MvNormal(mu, Symmetric(cov)):
 * -(k log 2pi + logdet S + (x-mu)' S^-1 (x-mu)) / 2 via Cholesky */
double go_logpdf_mvnormal(const double *x, const double *mu, const double *cov, int k) {
  double *L = (double *)calloc((size_t)k * k, sizeof(double));
  double *z = (double *)malloc((size_t)k * sizeof(double));
  double logdet = 0.0, quad = 0.0;
  for (int j = 0; j < k; j++) {
    double d = cov[j * k + j];
    for (int p = 0; p < j; p++) d -= L[j * k + p] * L[j * k + p];
    if (!(d > 0.0)) { free(L); free(z); return NAN; }
    d = sqrt(d);
    L[j * k + j] = d;
    logdet += 2.0 * log(d);
    for (int i = j + 1; i < k; i++) {
      double s = cov[i * k + j]; /* Symmetric(cov) uses the upper triangle; inputs are symmetric */
      for (int p = 0; p < j; p++) s -= L[i * k + p] * L[j * k + p];
      L[i * k + j] = s / d;
    }
  }
  for (int i = 0; i < k; i++) {
    double s = x[i] - mu[i];
    for (int p = 0; p < i; p++) s -= L[i * k + p] * z[p];
    z[i] = s / L[i * k + i];
    quad += z[i] * z[i];
  }
  free(L); free(z);
  return -(k * log(2.0 * M_PI) + logdet + quad) / 2.0;
}

/* ------------------------------------------------------------------ */
/* logsumexp / ESS                                                     */
/* ------------------------------------------------------------------ */

/* Julia's sum() is pairwise with a sequential base case; restated with base 1024. */
static double pairwise_sum_exp(const double *a, int64_t n, double shift, double scale) {
  if (n <= 1024) {
    double s = 0.0;
    for (int64_t i = 0; i < n; i++) s += exp(scale * a[i] - shift);
    return s;
  }
  int64_t h = n / 2;
  return pairwise_sum_exp(a, h, shift, scale) + pairwise_sum_exp(a + h, n - h, shift, scale);
}
static double sum_exp_shifted(const double *a, int64_t n, double shift, double scale) {
#ifdef _OPENMP
  if (g_threads > 1 && n > (1 << 16)) {
    double s = 0.0;
#pragma omp parallel for reduction(+ : s) schedule(static)
    for (int64_t c = 0; c < (n + 65535) / 65536; c++) {
      int64_t lo = c * 65536, hi = lo + 65536 > n ? n : lo + 65536;
      s += pairwise_sum_exp(a + lo, hi - lo, shift, scale);
    }
    return s;
  }
#endif
  return pairwise_sum_exp(a, n, shift, scale);
}
static double array_max(const double *a, int64_t n) {
  double m = -INFINITY;
#ifdef _OPENMP
  if (g_threads > 1 && n > (1 << 16)) {
#pragma omp parallel for reduction(max : m) schedule(static)
    for (int64_t i = 0; i < n; i++) if (a[i] > m) m = a[i];
    return m;
  }
#endif
  for (int64_t i = 0; i < n; i++) if (a[i] > m) m = a[i];
  return m;
}
/* inference.jl:3-6 */
double go_logsumexp(const double *a, int64_t n) {
  double m = array_max(a, n);
  if (m == -INFINITY) return -INFINITY;
  return m + log(sum_exp_shifted(a, n, m, 1.0));
}
/* inference.jl:8-11 */
double go_logsumexp2(double x1, double x2) {
  double m = x1 > x2 ? x1 : x2;
  if (m == -INFINITY) return m;
  return m + log(exp(x1 - m) + exp(x2 - m));
}
/* particle_filter.jl:3-6:  exp(-logsumexp(2 .* lnw)) */
double go_effective_sample_size(const double *lognormw, int64_t n) {
  double m = array_max(lognormw, n);
  double lse2;
  if (m == -INFINITY || isnan(m)) {
    /* all -Inf (or NaN after -Inf - -Inf): Julia yields NaN here, `NaN < thr` is false */
    int anynan = 0;
    for (int64_t i = 0; i < n; i++) if (isnan(lognormw[i])) anynan = 1;
    if (anynan) return NAN;
    lse2 = -INFINITY;
  } else {
    lse2 = 2.0 * m + log(sum_exp_shifted(lognormw, n, 2.0 * m, 2.0));
  }
  return exp(-lse2);
}
/* particle_filter.jl:8-12 */
double go_normalize_weights(const double *logw, int64_t n, double *out) {
  double lse = go_logsumexp(logw, n);
  if (out) for (int64_t i = 0; i < n; i++) out[i] = logw[i] - lse;
  return lse;
}

/* ------------------------------------------------------------------ */
/* Philox4x32-10 noise spec (shared, by value, with the engine)        */
/* ------------------------------------------------------------------ */
enum { PK_NORMAL = 1, PK_UNIFORM = 2, PK_RESAMPLE = 3, PK_MULTI = 4, PK_GAMMA = 5 };

void go_philox4x32_10(const uint32_t c_in[4], const uint32_t k_in[2], uint32_t out[4]) {
  uint32_t c0 = c_in[0], c1 = c_in[1], c2 = c_in[2], c3 = c_in[3];
  uint32_t k0 = k_in[0], k1 = k_in[1];
  for (int r = 0; r < 10; r++) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
static void philox_block(uint64_t seed, uint64_t pid, uint32_t step, uint32_t kind, uint32_t idx, uint32_t w[4]) {
  uint32_t c[4] = {(uint32_t)pid, (uint32_t)(pid >> 32), step, (kind << 24) | (idx & 0xFFFFFFu)};
  uint32_t k[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
  go_philox4x32_10(c, k, w);
}
/* 52-bit uniform strictly inside (0,1): ((hi:lo) >> 12 + 0.5) * 2^-52 (exact in fp64) */
static double u52(uint32_t hi, uint32_t lo) {
  uint64_t v = (((uint64_t)hi << 32) | lo) >> 12;
  return ((double)v + 0.5) * 0x1p-52;
}
static void philox_normal_pair(uint64_t seed, uint64_t pid, uint32_t step, uint32_t pair, double *n0, double *n1) {
  uint32_t w[4];
  philox_block(seed, pid, step, PK_NORMAL, pair, w);
  double u1 = u52(w[0], w[1]), u2 = u52(w[2], w[3]);
  double r = sqrt(-2.0 * log(u1));
  double th = 2.0 * M_PI * u2;
  *n0 = r * cos(th);
  *n1 = r * sin(th);
}
static void philox_uniform_pair(uint64_t seed, uint64_t pid, uint32_t step, uint32_t pair, double *a, double *b) {
  uint32_t w[4];
  philox_block(seed, pid, step, PK_UNIFORM, pair, w);
  *a = u52(w[0], w[1]);
  *b = u52(w[2], w[3]);
}
void go_philox_normals(uint64_t seed, uint32_t step, int64_t pid0, int64_t n, int ndraw, double *out) {
#pragma omp parallel for schedule(static) if (g_threads > 1)
  for (int64_t i = 0; i < n; i++)
    for (int d = 0; d < ndraw; d += 2) {
      double a, b;
      philox_normal_pair(seed, (uint64_t)(pid0 + i), step, (uint32_t)(d / 2), &a, &b);
      out[(int64_t)d * n + i] = a;
      if (d + 1 < ndraw) out[(int64_t)(d + 1) * n + i] = b;
    }
}
void go_philox_uniforms(uint64_t seed, uint32_t step, int64_t pid0, int64_t n, int ndraw, double *out) {
#pragma omp parallel for schedule(static) if (g_threads > 1)
  for (int64_t i = 0; i < n; i++)
    for (int d = 0; d < ndraw; d += 2) {
      double a, b;
      philox_uniform_pair(seed, (uint64_t)(pid0 + i), step, (uint32_t)(d / 2), &a, &b);
      out[(int64_t)d * n + i] = a;
      if (d + 1 < ndraw) out[(int64_t)(d + 1) * n + i] = b;
    }
}
uint32_t go_philox_resample_u0(uint64_t seed, uint32_t step) {
  uint32_t w[4];
  philox_block(seed, 0, step, PK_RESAMPLE, 0, w);
  return w[0];
}
void go_philox_multinomial_u64(uint64_t seed, uint32_t step, int64_t k0, int64_t n, uint64_t *out) {
  for (int64_t i = 0; i < n; i++) {
    uint32_t w[4];
    philox_block(seed, (uint64_t)(k0 + i), step, PK_MULTI, 0, w);
    out[i] = ((uint64_t)w[0] << 32) | w[1];
  }
}
/* gamma.jl:29-31 -> rand(Distributions.Gamma(shape, scale)) (third-party Marsaglia-Tsang;
 * unpinned).  Spec of this repo: attempt i uses Philox block (PK_GAMMA, 2i) for the normal
 * (first Box-Muller output) and block (PK_GAMMA, 2i+1) word pair 0 for the accept uniform;
 * shape < 1 boosts with word pair 1 of block (PK_GAMMA, 0xFFFFFF). */
double go_random_gamma_philox(double shape, double scale, uint64_t seed, uint64_t pid, uint32_t step) {
  double a = shape < 1.0 ? shape + 1.0 : shape;
  double d = a - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d), g = d;
  for (uint32_t i = 0; i < 64; i++) {
    uint32_t w[4], w2[4];
    philox_block(seed, pid, step, PK_GAMMA, 2 * i, w);
    philox_block(seed, pid, step, PK_GAMMA, 2 * i + 1, w2);
    double u1 = u52(w[0], w[1]), u2 = u52(w[2], w[3]);
    double x = sqrt(-2.0 * log(u1)) * cos(2.0 * M_PI * u2);
    double v = 1.0 + c * x;
    if (v <= 0.0) continue;
    v = v * v * v;
    double u = u52(w2[0], w2[1]);
    if (log(u) < 0.5 * x * x + d - d * v + d * log(v)) { g = d * v; break; }
  }
  if (shape < 1.0) {
    uint32_t w[4];
    philox_block(seed, pid, step, PK_GAMMA, 0xFFFFFFu, w);
    g *= pow(u52(w[2], w[3]), 1.0 / shape);
  }
  return g * scale;
}

/* ------------------------------------------------------------------ */
/* Integer-CDF resamplers                       */
/* ------------------------------------------------------------------ */

/* Deterministic exp(x), -708 < x <= 0: only rint, fma, * and exact power-of-two scaling, so the CUDA restatement
 * (csrc/gbmath.cuh wexp_parts) produces identical bits.  x = j ln2/64 + r, j = rint(x 64/ln2), |r| <= ln2/128:
 * exp(x) = 2^(j >> 6) * T[j & 63] * (1 + r + r^2/2 + r^3/6 + r^4/24 + r^5/120); T = 2^(i/64) correctly rounded
 * (oracle/wexp_table.inc, generated by tools/gen_math_tables.py together with the engine's copy).  ~1.5 ulp; not libm. */
#include "wexp_table.inc"
static const double wexp_tab[64] = {GB_WEXP_TAB_VALUES};
static double dexp_parts(double x, int *kout) {
  double jd = rint(x * GB_WEXP_SCALE);
  int j = (int)jd;
  double r = fma(jd, GB_WEXP_NL_HI, x);
  r = fma(jd, GB_WEXP_NL_LO, r);
  double r2 = r * r;
  double a = fma(r, 1.0 / 120.0, 1.0 / 24.0);
  double b = fma(r, 1.0 / 6.0, 0.5);
  double c = fma(r2, a, b);
  double poly = fma(r2, c, r);
  double t = wexp_tab[j & 63];
  *kout = j >> 6; /* arithmetic shift: floor(j / 64) */
  return fma(t, poly, t);
}
static double pow2i(int e) { /* exact 2^e for -1022 <= e <= 1023 */
  uint64_t bits = (uint64_t)(e + 1023) << 52;
  double d;
  memcpy(&d, &bits, 8);
  return d;
}
double go_dexp(double x) {
  if (isnan(x)) return x;
  if (x > 0.0) x = 0.0;
  if (x < -708.0) return 0.0;
  int k;
  double p = dexp_parts(x, &k);
  return p * pow2i(k);
}
int go_quant_bits(int64_t n) {
  int lg = 0;
  while (((int64_t)1 << lg) < n) lg++;
  int b = 62 - lg;
  return b > 52 ? 52 : b;
}
uint64_t go_quantize(double x, int b) {
  if (!(x > -708.0)) return 0; /* also NaN, -Inf */
  if (x > 0.0) x = 0.0;
  int k;
  double p = dexp_parts(x, &k);
  int e = k + b;
  if (e < 0) return 0;
  return (uint64_t)(p * pow2i(e));
}
uint64_t go_quantized_weights(const double *logw, int64_t n, double m, int b, uint64_t *q_out) {
  uint64_t total = 0;
  for (int64_t i = 0; i < n; i++) {
    uint64_t q = go_quantize(logw[i] - m, b);
    if (q_out) q_out[i] = q;
    total += q;
  }
  return total;
}
/* systematic: tau_k = floor((k*2^32 + U0) * Q_N / (N * 2^32)); a_k = min{ j : Q_j > tau_k } */
static void systematic_from_q(const uint64_t *q, uint64_t qn, int64_t nslots, uint32_t u0, int64_t *anc) {
  int64_t j = 0;
  u128 cdf = q[0];
  for (int64_t k = 0; k < nslots; k++) {
    u128 num = ((u128)k << 32) + u0;
    u128 tau = num * qn / ((u128)nslots << 32);
    while (!(cdf > tau)) { j++; cdf += q[j]; }
    anc[k] = j;
  }
}
int go_resample_systematic(const double *logw, int64_t n, uint32_t u0, int64_t *anc) {
  double m = array_max(logw, n);
  if (m == -INFINITY || isnan(m)) return -1;
  int b = go_quant_bits(n);
  uint64_t *q = (uint64_t *)malloc((size_t)n * 8);
  uint64_t qn = go_quantized_weights(logw, n, m, b, q);
  systematic_from_q(q, qn, n, u0, anc);
  free(q);
  return 0;
}
/* multinomial (particle_filter.jl:261-262 semantics: N iid categorical draws, slot order):
 * tau_k = floor(U_k * Q_N / 2^64) */
int go_resample_multinomial(const double *logw, int64_t n, const uint64_t *u, int64_t *anc) {
  double m = array_max(logw, n);
  if (m == -INFINITY || isnan(m)) return -1;
  int b = go_quant_bits(n);
  uint64_t *cdf = (uint64_t *)malloc((size_t)n * 8);
  uint64_t qn = go_quantized_weights(logw, n, m, b, cdf);
  uint64_t run = 0;
  for (int64_t i = 0; i < n; i++) { run += cdf[i]; cdf[i] = run; }
  for (int64_t k = 0; k < n; k++) {
    uint64_t tau = (uint64_t)(((u128)u[k] * qn) >> 64);
    int64_t lo = 0, hi = n - 1; /* smallest j with cdf[j] > tau */
    while (lo < hi) {
      int64_t mid = (lo + hi) / 2;
      if (cdf[mid] > tau) hi = mid; else lo = mid + 1;
    }
    anc[k] = lo;
  }
  free(cdf);
  return 0;
}
/* sorted multinomial (production mode, not in the reference): the ORDER STATISTICS of N iid uniforms on [0, Q_N) through
 * exponential spacings -- U_(k) = (E_0 + ... + E_k) / (E_0 + ... + E_N) for iid exponential E_j, a classical identity
 * (e.g. Devroye 1986, ch. V.3) -- so the draws come out sorted by ancestor and one merge with the integer CDF replaces N
 * binary searches.  The spacings are INTEGERS e[0..n] (>= 1; the engine derives them from Philox, the tests pass them
 * in), everything below is exact:  T_k = e_0 + ... + e_k,  T = T_n,
 *     a_k = min{ j : C_j * T > T_k * Q_N }      (C = inclusive CDF of the quantised weights; T_k < T, so a_k exists).
 * As a multiset the ancestors have the distribution of particle_filter.jl:261-262's N iid categorical draws (up to the
 * 2^-30 resolution of the spacings); slot order differs (sorted instead of iid), which Gen's particle filter never
 * looks at: traces are exchangeable after a resample. */
int go_resample_multinomial_sorted(const double *logw, int64_t n, const uint64_t *e, int64_t *anc) {
  double m = array_max(logw, n);
  if (m == -INFINITY || isnan(m)) return -1;
  int b = go_quant_bits(n);
  uint64_t *q = (uint64_t *)malloc((size_t)n * 8);
  uint64_t qn = go_quantized_weights(logw, n, m, b, q);
  u128 total = 0;
  for (int64_t k = 0; k <= n; k++) total += e[k];
  int64_t j = 0;
  u128 cdf = q[0], t = 0;
  for (int64_t k = 0; k < n; k++) {
    t += e[k];
    while (!(cdf * total > t * qn)) { j++; cdf += q[j]; }
    anc[k] = j;
  }
  free(q);
  return 0;
}
/* residual: c_j = floor(N q_j / Q_N) copies; R = N - sum c_j slots by systematic resampling on
 * r_j = N q_j - c_j Q_N (sum r_j = R * Q_N).  Output sorted by ancestor. */
int go_resample_residual(const double *logw, int64_t n, uint32_t u0, int64_t *anc) {
  double m = array_max(logw, n);
  if (m == -INFINITY || isnan(m)) return -1;
  int b = go_quant_bits(n);
  uint64_t *q = (uint64_t *)malloc((size_t)n * 8);
  uint64_t qn = go_quantized_weights(logw, n, m, b, q);
  int64_t *cnt = (int64_t *)calloc((size_t)n, 8);
  int64_t copies = 0;
  for (int64_t j = 0; j < n; j++) {
    u128 nq = (u128)n * q[j];
    cnt[j] = (int64_t)(nq / qn);
    q[j] = (uint64_t)(nq - (u128)cnt[j] * qn); /* residual r_j < Q_N */
    copies += cnt[j];
  }
  int64_t R = n - copies;
  if (R > 0) {
    /* tau'_k = floor((k*2^32+u0) * (R*Q_N) / (R*2^32)) = floor((k*2^32+u0) * Q_N / 2^32) */
    int64_t j = 0;
    u128 cdf = q[0];
    for (int64_t k = 0; k < R; k++) {
      u128 tau = ((((u128)k << 32) + u0) * qn) >> 32;
      while (!(cdf > tau)) { j++; cdf += q[j]; }
      cnt[j]++;
    }
  }
  int64_t k = 0;
  for (int64_t j = 0; j < n; j++)
    for (int64_t c = 0; c < cnt[j]; c++) anc[k++] = j;
  free(q); free(cnt);
  return k == n ? 0 : -2;
}

/* ------------------------------------------------------------------ */
/* Exact log-ML oracles                                                */
/* ------------------------------------------------------------------ */

/* scalar Kalman filter for cfg 1: x0~N(m0,s0), x_t~N(a x,q), y_t~N(x_t,r) */
double go_kalman_logml(double m0, double s0, double a, double q, double r, const double *y, int T1) {
  double mean = m0, P = s0 * s0, ll = 0.0;
  for (int t = 0; t < T1; t++) {
    if (t > 0) { mean = a * mean; P = a * a * P + q * q; }
    double v = y[t] - mean, F = P + r * r;
    ll += -0.5 * (log(2.0 * M_PI) + log(F) + v * v / F);
    double K = P / F;
    mean = mean + K * v;
    P = (1.0 - K) * P;
  }
  return ll;
}
/* test/inference/particle_filter.jl:1-27 (hmm_forward_alg), emissions 1-based */
double go_hmm_forward(const double *prior, const double *emission, const double *transition, int K, int M,
                      const int64_t *obs, int T) {
  double marg = 1.0;
  double *alpha = (double *)malloc(K * sizeof(double)), *post = (double *)malloc(K * sizeof(double));
  for (int i = 0; i < K; i++) alpha[i] = prior[i];
  for (int t = 1; t < T; t++) {
    double denom = 0.0;
    for (int z = 0; z < K; z++) { post[z] = alpha[z] * emission[z * M + (obs[t - 1] - 1)]; denom += post[z]; }
    for (int z = 0; z < K; z++) post[z] /= denom;
    for (int z = 0; z < K; z++) {
      double s = 0.0;
      for (int p = 0; p < K; p++) s += transition[p * K + z] * post[p];
      alpha[z] = s;
    }
    marg *= denom;
  }
  double denom = 0.0;
  for (int z = 0; z < K; z++) denom += alpha[z] * emission[z * M + (obs[T - 1] - 1)];
  marg *= denom;
  free(alpha); free(post);
  return marg;
}
/* cfg 4 exact evidence: log N(y; 0, sigma^2 I_n + X X') */
double go_blr_log_evidence(const double *X, const double *y, int n, int D, double sigma) {
  double *C = (double *)malloc((size_t)n * n * sizeof(double));
  double *zero = (double *)calloc(n, sizeof(double));
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++) {
      double s = 0.0;
      for (int d = 0; d < D; d++) s += X[i * D + d] * X[j * D + d];
      C[i * n + j] = s + (i == j ? sigma * sigma : 0.0);
    }
  double r = go_logpdf_mvnormal(y, zero, C, n);
  free(C); free(zero);
  return r;
}

/* ------------------------------------------------------------------ */
/* Models and the particle filter               */
/* ------------------------------------------------------------------ */
struct go_pf {
  int kind, nparams, S, step;
  double *params;
  int64_t n;
  uint64_t seed;
  double *state;     /* [S][n]  current traces (SoA view of `traces`)            */
  double *new_state; /* [S][n]  `new_traces`                                      */
  double *logw;      /* particle_filter.jl:38 */
  double log_ml_est; /* :39 */
  int64_t *parents;  /* :40, 1-based */
  int mh_ok;         /* new_state still holds the parents of `state` (rows aligned) */
  int storage_f32;   /* GB_F32 storage model: traces, log weights and increments are rounded to float when stored */
  int win_W, win_len; /* sliding-window rejuvenation: the last win_W states of every particle's own lineage */
  double *win;       /* [win_W][S][n]: win[k] = the state k+1 steps before the current one */
};

/* Storage model of the engine's GB_F32 filters (this repo's extension; Gen itself is Float64 only): the per-particle
 * arithmetic below stays fp64, but every value written to the particle store -- trace fields, log weights, log
 * incremental weights -- is rounded to single precision first.  Applies to filters created AFTER the call. */
static int g_storage_f32 = 0;
void go_set_storage_f32(int on) { g_storage_f32 = on != 0; }
static inline double st_round(const go_pf_t *pf, double v) { return pf->storage_f32 ? (double)(float)v : v; }

int go_model_state_dim(int kind, const double *p) {
  switch (kind) {
    case GO_MODEL_LGSSM: case GO_MODEL_SV: case GO_MODEL_HMM: return 1;
    case GO_MODEL_BEARINGS: return 4;
    case GO_MODEL_BLR: return (int)p[0];
  }
  return -1;
}
int go_model_num_obs(int kind, const double *p) { return kind == GO_MODEL_BLR ? (int)p[1] : 1; }
int go_model_num_normals(int kind, const double *p, int proposal_kind, int is_init) {
  (void)proposal_kind;
  switch (kind) {
    case GO_MODEL_LGSSM: case GO_MODEL_SV: return 1;
    case GO_MODEL_BEARINGS: return is_init ? 4 : 2;
    case GO_MODEL_HMM: return 0;
    case GO_MODEL_BLR: return is_init ? (int)p[0] : 0;
  }
  return 0;
}
int go_model_num_uniforms(int kind, const double *p, int proposal_kind, int is_init) {
  (void)p; (void)proposal_kind; (void)is_init;
  return kind == GO_MODEL_HMM ? 1 : 0;
}

/* One particle through `generate` (t == 0) or the extend-by-one `update` (t >= 1).
 * Weight semantics: only constrained choices contribute (dynamic/generate.jl:26-41,
 * static_ir/generate.jl:31-42, unfold/update.jl:54-78); with a custom proposal the proposed
 * latents are constrained too and the proposal score is subtracted once
 * (particle_filter.jl:128-130, trace_translators.jl:854, importance.jl:53). */
static double particle_transition(int kind, const double *p, int t, const double *prev, double *next,
                                  const double *obs, int pk, const double *pa, const double *eps,
                                  const double *us) {
  switch (kind) {
    case GO_MODEL_LGSSM: {
      /* p = [m0, s0, a, q, r] */
      double mu = t == 0 ? p[0] : p[2] * prev[0];
      double sd = t == 0 ? p[1] : p[3];
      if (pk == GO_PROPOSAL_AFFINE_NORMAL) {
        double qmu = t == 0 ? pa[0] : pa[0] + pa[1] * prev[0];
        double x = go_random_normal(qmu, pa[2], eps[0]);
        double lq = go_logpdf_normal(x, qmu, pa[2]);
        double w = 0.0;
        w += go_logpdf_normal(x, mu, sd);
        w += go_logpdf_normal(obs[0], x, p[4]);
        next[0] = x;
        return w - lq;
      }
      double x = go_random_normal(mu, sd, eps[0]);
      next[0] = x;
      return go_logpdf_normal(obs[0], x, p[4]);
    }
    case GO_MODEL_SV: {
      /* p = [mu, phi, sigma, sd0]  (sd0 = sigma / sqrt(1 - phi^2), computed by the caller) */
      double mean = t == 0 ? p[0] : p[0] + p[1] * (prev[0] - p[0]);
      double sd = t == 0 ? p[3] : p[2];
      double h = go_random_normal(mean, sd, eps[0]);
      next[0] = h;
      return go_logpdf_normal(obs[0], 0.0, exp(h / 2.0));
    }
    case GO_MODEL_BEARINGS: {
      /* docs/src/tutorials/smc.md:607-661.
       * p = [x0mu,x0sd,y0mu,y0sd,vx0mu,vx0sd,vy0mu,vy0sd, velocity_var, measurement_noise] */
      double x, y, vx, vy;
      if (t == 0) {
        x = go_random_normal(p[0], p[1], eps[0]);
        y = go_random_normal(p[2], p[3], eps[1]);
        vx = go_random_normal(p[4], p[5], eps[2]);
        vy = go_random_normal(p[6], p[7], eps[3]);
      } else {
        double sdv = sqrt(p[8]);
        vx = go_random_normal(prev[2], sdv, eps[0]);
        vy = go_random_normal(prev[3], sdv, eps[1]);
        x = prev[0] + vx;
        y = prev[1] + vy;
      }
      next[0] = x; next[1] = y; next[2] = vx; next[3] = vy;
      return go_logpdf_normal(obs[0], atan2(y, x), p[9]);
    }
    case GO_MODEL_HMM: {
      /* test/inference/particle_filter.jl:52-78.
       * p = [K, M, prior[K], transition[K*K] (col = prev_z), emission[M*K] (col = z)] */
      int K = (int)p[0], M = (int)p[1];
      const double *prior = p + 2, *trans = p + 2 + K, *emis = p + 2 + K + K * K;
      const double *zp = t == 0 ? prior : trans + ((int64_t)prev[0] - 1) * K;
      int64_t xo = (int64_t)obs[0];
      if (pk == GO_PROPOSAL_CATEGORICAL_TABLE) {
        const double *qp = t == 0 ? pa : pa + ((int64_t)prev[0] - 1) * K;
        int64_t z = go_random_categorical(qp, K, us[0]);
        double lq = go_logpdf_categorical(z, qp, K);
        double w = 0.0;
        w += go_logpdf_categorical(z, zp, K);
        w += go_logpdf_categorical(xo, emis + (z - 1) * M, M);
        next[0] = (double)z;
        return w - lq;
      }
      int64_t z = go_random_categorical(zp, K, us[0]);
      next[0] = (double)z;
      return go_logpdf_categorical(xo, emis + (z - 1) * M, M);
    }
    case GO_MODEL_BLR: {
      /* p = [D, n, sigma, X[n*D] row-major]; w ~ Map(normal)(0,1); y_j ~ normal(x_j . w, sigma) */
      int D = (int)p[0], n = (int)p[1];
      double sigma = p[2];
      const double *X = p + 3;
      for (int d = 0; d < D; d++) next[d] = go_random_normal(0.0, 1.0, eps[d]);
      double w = 0.0;
      for (int j = 0; j < n; j++) {
        double mu = 0.0;
        for (int d = 0; d < D; d++) mu += X[j * D + d] * next[d];
        w += go_logpdf_normal(obs[j], mu, sigma); /* map/generate.jl:16-17 accumulates in order */
      }
      return w;
    }
  }
  return NAN;
}

static void run_particles(go_pf_t *pf, int t, const double *obs, int pk, const double *pa,
                          const double *normals, const double *uniforms, double *w_out) {
  const int S = pf->S, kind = pf->kind;
  const int64_t n = pf->n;
  const int nn = go_model_num_normals(kind, pf->params, pk, t == 0);
  const int nu = go_model_num_uniforms(kind, pf->params, pk, t == 0);
#pragma omp parallel for schedule(static) if (g_threads > 1)
  for (int64_t i = 0; i < n; i++) {
    double prev[64], next[64], eps[64], us[4];
    for (int s = 0; s < S && t > 0; s++) prev[s] = pf->state[(int64_t)s * n + i];
    if (normals) {
      for (int d = 0; d < nn; d++) eps[d] = normals[(int64_t)d * n + i];
    } else {
      for (int d = 0; d < nn; d += 2) {
        double a, b;
        philox_normal_pair(pf->seed, (uint64_t)i, (uint32_t)t, (uint32_t)(d / 2), &a, &b);
        eps[d] = a;
        if (d + 1 < nn) eps[d + 1] = b;
      }
    }
    if (uniforms) {
      for (int d = 0; d < nu; d++) us[d] = uniforms[(int64_t)d * n + i];
    } else {
      for (int d = 0; d < nu; d += 2) {
        double a, b;
        philox_uniform_pair(pf->seed, (uint64_t)i, (uint32_t)t, (uint32_t)(d / 2), &a, &b);
        us[d] = a;
        if (d + 1 < nu) us[d + 1] = b;
      }
    }
    w_out[i] = particle_transition(kind, pf->params, t, prev, next, obs, pk, pa, eps, us);
    for (int s = 0; s < S; s++) pf->new_state[(int64_t)s * n + i] = st_round(pf, next[s]);
  }
}
static void swap_traces(go_pf_t *pf) { double *t = pf->state; pf->state = pf->new_state; pf->new_state = t; }

/* particle_filter.jl:118-154 */
go_pf_t *go_pf_init(int kind, const double *params, int nparams, int64_t n, uint64_t seed,
                    const double *obs, int pk, const double *pa, const double *normals,
                    const double *uniforms) {
  int S = go_model_state_dim(kind, params);
  if (S < 1 || S > 64 || n < 1) return NULL;
  go_pf_t *pf = (go_pf_t *)calloc(1, sizeof(*pf));
  pf->kind = kind; pf->nparams = nparams; pf->S = S; pf->n = n; pf->seed = seed; pf->step = 0;
  pf->storage_f32 = g_storage_f32;
  pf->params = (double *)malloc((size_t)nparams * 8);
  memcpy(pf->params, params, (size_t)nparams * 8);
  pf->state = (double *)calloc((size_t)S * n, 8);
  pf->new_state = (double *)calloc((size_t)S * n, 8);
  pf->logw = (double *)malloc((size_t)n * 8);
  pf->parents = (int64_t *)malloc((size_t)n * 8);
  run_particles(pf, 0, obs, pk, pa, normals, uniforms, pf->logw); /* :127-131, :149-151 */
  if (pf->storage_f32)
    for (int64_t i = 0; i < n; i++) pf->logw[i] = (double)(float)pf->logw[i];
  swap_traces(pf);
  pf->log_ml_est = 0.0;                                            /* :133, :153 */
  for (int64_t i = 0; i < n; i++) pf->parents[i] = i + 1;         /* collect(1:num_particles) */
  pf->mh_ok = kind != GO_MODEL_BLR;
  return pf;
}
/* particle_filter.jl:186-240 */
int go_pf_step(go_pf_t *pf, const double *obs, int pk, const double *pa, const double *normals,
               const double *uniforms, double *incr_out) {
  double *incr = incr_out ? incr_out : (double *)malloc((size_t)pf->n * 8);
  pf->step += 1;
  run_particles(pf, pf->step, obs, pk, pa, normals, uniforms, incr);
#pragma omp parallel for schedule(static) if (g_threads > 1)
  for (int64_t i = 0; i < pf->n; i++) {
    pf->logw[i] = st_round(pf, pf->logw[i] + incr[i]);            /* :199, :231 */
    incr[i] = st_round(pf, incr[i]);
  }
  if (pf->win_W > 0) { /* the window follows the trace: the state the step extended becomes lag 1 */
    const size_t blk = (size_t)pf->S * pf->n;
    for (int k = pf->win_W - 1; k >= 1; k--) memcpy(pf->win + (size_t)k * blk, pf->win + (size_t)(k - 1) * blk, blk * 8);
    memcpy(pf->win, pf->state, blk * 8);
    if (pf->win_len < pf->win_W) pf->win_len++;
  }
  swap_traces(pf);                                                 /* :203-205, :234-237 */
  pf->mh_ok = 1;
  if (!incr_out) free(incr);
  return 0;
}
/* Gen's traces are persistent copies: resampling particle i from ancestor a copies a's whole past.  Here: the window rows. */
static void window_gather(go_pf_t *pf, const int64_t *anc0 /* 0-based */) {
  if (pf->win_W == 0) return;
  const int64_t n = pf->n;
  const size_t rows = (size_t)pf->win_W * pf->S;
  double *tmp = (double *)malloc((size_t)n * 8);
  for (size_t r = 0; r < rows; r++) {
    double *col = pf->win + r * n;
    for (int64_t i = 0; i < n; i++) tmp[i] = col[anc0[i]];
    memcpy(col, tmp, (size_t)n * 8);
  }
  free(tmp);
}
void go_pf_enable_window(go_pf_t *pf, int W) {
  free(pf->win);
  pf->win = W > 0 ? (double *)calloc((size_t)W * pf->S * pf->n, 8) : NULL;
  pf->win_W = W; pf->win_len = 0;
}
/* particle_filter.jl:249-275 */
int go_pf_maybe_resample(go_pf_t *pf, double ess_threshold, int kind, int use_given_u0, uint32_t u0,
                         const uint64_t *u64s, double *ess_out) {
  const int64_t n = pf->n;
  double *lnw = (double *)malloc((size_t)n * 8);
  double lse = go_normalize_weights(pf->logw, n, lnw);             /* :254 */
  double ess = go_effective_sample_size(lnw, n);                   /* :255 */
  free(lnw);
  if (ess_out) *ess_out = ess;
  int do_resample = ess < ess_threshold;                           /* :256 (strict; NaN => false) */
  if (!do_resample) return 0;
  int64_t *anc = (int64_t *)malloc((size_t)n * 8);
  /* resampling randomness is keyed by the index of the step that produced the weights */
  uint32_t u = use_given_u0 ? u0 : go_philox_resample_u0(pf->seed, (uint32_t)pf->step);
  int rc = 0;
  if (kind == GO_RESAMPLE_SYSTEMATIC) rc = go_resample_systematic(pf->logw, n, u, anc);
  else if (kind == GO_RESAMPLE_RESIDUAL) rc = go_resample_residual(pf->logw, n, u, anc);
  else if (kind == GO_RESAMPLE_MULTINOMIAL_SORTED) {
    /* u64s = the n + 1 integer spacings (required: the engine's come from its own table-driven log, exported to the tests) */
    rc = u64s ? go_resample_multinomial_sorted(pf->logw, n, u64s, anc) : -2;
  } else {
    uint64_t *tmp = NULL;
    if (!u64s) {
      tmp = (uint64_t *)malloc((size_t)n * 8);
      go_philox_multinomial_u64(pf->seed, (uint32_t)pf->step, 0, n, tmp);
    }
    rc = go_resample_multinomial(pf->logw, n, u64s ? u64s : tmp, anc); /* :261-262 */
    free(tmp);
  }
  if (rc != 0) { free(anc); return rc; }
  pf->log_ml_est += lse - log((double)n);                          /* :263 */
#pragma omp parallel for schedule(static) if (g_threads > 1)
  for (int64_t i = 0; i < n; i++) {                                /* :264-267 */
    pf->parents[i] = anc[i] + 1;
    for (int s = 0; s < pf->S; s++) pf->new_state[(int64_t)s * n + i] = pf->state[(int64_t)s * n + anc[i]];
    pf->logw[i] = 0.0;
  }
  swap_traces(pf);                                                 /* :270-272 */
  window_gather(pf, anc);
  pf->mh_ok = 0;
  free(anc);
  return 1;
}
/* The resampling branch of maybe_resample! (particle_filter.jl:261-272) with the ancestors GIVEN (1-based `parents`)
 * instead of drawn: log_ml_est is updated from this filter's own weights, traces are gathered, weights reset.  Used to
 * keep two filters on the same genealogy while each keeps its own arithmetic (fp32-storage engine vs fp64 oracle). */
int go_pf_resample_given(go_pf_t *pf, const int64_t *parents) {
  const int64_t n = pf->n;
  for (int64_t i = 0; i < n; i++)
    if (parents[i] < 1 || parents[i] > n) return -1;
  double lse = go_logsumexp(pf->logw, n);
  pf->log_ml_est += lse - log((double)n);                          /* :263 */
#pragma omp parallel for schedule(static) if (g_threads > 1)
  for (int64_t i = 0; i < n; i++) {                                /* :264-267 */
    pf->parents[i] = parents[i];
    for (int s = 0; s < pf->S; s++) pf->new_state[(int64_t)s * n + i] = pf->state[(int64_t)s * n + (parents[i] - 1)];
    pf->logw[i] = 0.0;
  }
  swap_traces(pf);                                                 /* :270-272 */
  if (pf->win_W > 0) {
    int64_t *anc0 = (int64_t *)malloc((size_t)n * 8);
    for (int64_t i = 0; i < n; i++) anc0[i] = parents[i] - 1;
    window_gather(pf, anc0);
    free(anc0);
  }
  pf->mh_ok = 0;
  return 1;
}
/* ---- MH rejuvenation of the latest latent choices (src/inference/mh.jl) ----------------------------------------
 * Scores of the choices of the latest kernel application: the observation given the state and the latent choices
 * given the parent state.  Same distributions and arguments as particle_transition above. */
static double model_obs_logpdf(int kind, const double *p, const double *st, const double *obs) {
  switch (kind) {
    case GO_MODEL_LGSSM: return go_logpdf_normal(obs[0], st[0], p[4]);
    case GO_MODEL_SV: return go_logpdf_normal(obs[0], 0.0, exp(st[0] / 2.0));
    case GO_MODEL_BEARINGS: return go_logpdf_normal(obs[0], atan2(st[1], st[0]), p[9]);
    case GO_MODEL_HMM: {
      int K = (int)p[0], M = (int)p[1];
      return go_logpdf_categorical((int64_t)obs[0], p + 2 + K + K * K + ((int64_t)st[0] - 1) * M, M);
    }
  }
  return NAN;
}
/* number of latent choices of the latest kernel application and their values c[] inside the state */
static int model_latent_choices(int kind, int t, const int **idx) {
  static const int one[1] = {0}, all4[4] = {0, 1, 2, 3}, vel[2] = {2, 3};
  if (kind == GO_MODEL_BEARINGS) { *idx = t == 0 ? all4 : vel; return t == 0 ? 4 : 2; }
  *idx = one;
  return 1;
}
/* score of latent choice k (value v) given the parent state */
static double model_latent_score(int kind, const double *p, int t, const double *prev, int k, double v) {
  switch (kind) {
    case GO_MODEL_LGSSM: return go_logpdf_normal(v, t == 0 ? p[0] : p[2] * prev[0], t == 0 ? p[1] : p[3]);
    case GO_MODEL_SV: return go_logpdf_normal(v, t == 0 ? p[0] : p[0] + p[1] * (prev[0] - p[0]), t == 0 ? p[3] : p[2]);
    case GO_MODEL_BEARINGS:
      if (t == 0) return go_logpdf_normal(v, p[2 * k], p[2 * k + 1]);
      return go_logpdf_normal(v, prev[2 + k], sqrt(p[8]));
    case GO_MODEL_HMM: {
      int K = (int)p[0];
      return go_logpdf_categorical((int64_t)v, t == 0 ? p + 2 : p + 2 + K + ((int64_t)prev[0] - 1) * K, K);
    }
  }
  return NAN;
}
/* the state as a function of the latent choices (bearings: positions follow from the velocities) */
static void model_state_from_choices(int kind, int t, const double *prev, const double *c, double *st) {
  if (kind == GO_MODEL_BEARINGS && t > 0) {
    st[2] = c[0]; st[3] = c[1];
    st[0] = prev[0] + c[0];
    st[1] = prev[1] + c[1];
    return;
  }
  int S = kind == GO_MODEL_BEARINGS ? 4 : 1;
  for (int s = 0; s < S; s++) st[s] = c[s];
}

/* One MH move on every particle.  move 0: mh(trace, selection of the latest latents) -- mh.jl:16-27, the weight is
 * regenerate's: sum of (score - prev_score) over retained choices (dynamic/regenerate.jl:47-50) = the observation's.
 * move 1: mh(trace, proposal, args) with c' ~ normal(c, drift_std) per latent choice -- mh.jl:37-58; the update
 * weight is sum of (score - prev_score) over constrained latents and the observation (dynamic/update.jl:46-50).
 * Noise: pre-drawn rows (normals: the transition's / one per drifted choice; uniforms: the transition's rows then one
 * accept row) or Philox draw-index block 0x800000 + 16 * move_index.  Requires the parents to be resident: call
 * after init/step and before a resample.  alpha_out/accepted_out may be NULL.  Returns #accepted or -1. */
int64_t go_pf_rejuvenate(go_pf_t *pf, int move, const double *obs, double drift_std, uint32_t move_index,
                         const double *normals, const double *uniforms, double *alpha_out, uint8_t *accepted_out) {
  const int kind = pf->kind, S = pf->S, t = pf->step;
  const int64_t n = pf->n;
  if (!pf->mh_ok || kind == GO_MODEL_BLR || (move == 1 && kind == GO_MODEL_HMM)) return -1;
  const int *cidx;
  const int nc = model_latent_choices(kind, t, &cidx);
  const int nn = move == 1 ? nc : go_model_num_normals(kind, pf->params, 0, t == 0);
  const int nu = move == 1 ? 0 : go_model_num_uniforms(kind, pf->params, 0, t == 0);
  const uint32_t base = 0x800000u + 16u * move_index;
  int64_t total = 0;
#pragma omp parallel for schedule(static) reduction(+ : total) if (g_threads > 1)
  for (int64_t i = 0; i < n; i++) {
    double prev[8] = {0}, st[8], next[8], eps[8], us[4], ua;
    for (int s = 0; s < S; s++) { st[s] = pf->state[(int64_t)s * n + i]; if (t > 0) prev[s] = pf->new_state[(int64_t)s * n + i]; }
    if (normals || uniforms) {
      for (int d = 0; d < nn; d++) eps[d] = normals[(int64_t)d * n + i];
      for (int d = 0; d < nu; d++) us[d] = uniforms[(int64_t)d * n + i];
      ua = uniforms[(int64_t)nu * n + i];
    } else {
      for (int d = 0; d < nn; d += 2) {
        double a, b;
        philox_normal_pair(pf->seed, (uint64_t)i, (uint32_t)t, base + (uint32_t)(d / 2), &a, &b);
        eps[d] = a;
        if (d + 1 < nn) eps[d + 1] = b;
      }
      double a, b;
      philox_uniform_pair(pf->seed, (uint64_t)i, (uint32_t)t, base + 15u, &a, &b);
      ua = a;
      us[0] = b;
    }
    double alpha;
    if (move == 0) {
      double w_new = particle_transition(kind, pf->params, t, prev, next, obs, 0, NULL, eps, us);
      alpha = 0.0;
      alpha += w_new - model_obs_logpdf(kind, pf->params, st, obs);          /* regenerate.jl:47-50 */
    } else {
      double c_new[4], fwd = 0.0, bwd = 0.0, weight = 0.0;
      for (int k = 0; k < nc; k++) {
        double c = st[cidx[k]];
        c_new[k] = go_random_normal(c, drift_std, eps[k]);                   /* proposal: {addr} ~ normal(choices[addr], std) */
        fwd += go_logpdf_normal(c_new[k], c, drift_std);                     /* propose: mh.jl:44 */
      }
      model_state_from_choices(kind, t, prev, c_new, next);
      for (int k = 0; k < nc; k++)                                           /* update.jl:46-50 */
        weight += model_latent_score(kind, pf->params, t, prev, k, c_new[k]) -
                  model_latent_score(kind, pf->params, t, prev, k, st[cidx[k]]);
      weight += model_obs_logpdf(kind, pf->params, next, obs) - model_obs_logpdf(kind, pf->params, st, obs);
      for (int k = 0; k < nc; k++) bwd += go_logpdf_normal(st[cidx[k]], c_new[k], drift_std);   /* assess: mh.jl:48 */
      alpha = weight - fwd + bwd;                                            /* mh.jl:49 */
    }
    int acc = log(ua) < alpha;                                               /* mh.jl:20, :52 */
    if (alpha_out) alpha_out[i] = alpha;
    if (accepted_out) accepted_out[i] = (uint8_t)acc;
    if (acc) {
      for (int s = 0; s < S; s++) pf->state[(int64_t)s * n + i] = st_round(pf, next[s]);
      total++;
    }
  }
  return total;
}
/* Sliding-window block move (docs/src/tutorials/smc.md:492-541): mh(trace, perturbation_proposal, (a, b)) with
 * {latent choices of step t} ~ normal(current value, drift_std) for a <= t <= b (mh.jl:37-58).  The update weight
 * (dynamic/update.jl:46-62) sums score' - score over every choice whose value or arguments change: the latent choices of
 * steps a..T under their parents (unperturbed ones cancel unless their parent moved) and the observations of steps a..T.
 * Noise: pre-drawn rows (normals [(b-a+1) * choices per step][n], step-major; uniforms [1][n] = the accept draw) or Philox
 * draw-index block 0x900000 + 64 * move_index (pairs) and pair + 63 for the accept draw.  Returns #accepted, -1 if the
 * window does not reach back to step a - 1. */
int64_t go_pf_rejuvenate_window(go_pf_t *pf, int a, int b, const double *obs_window, double drift_std, uint32_t move_index,
                                const double *normals, const double *uniforms, double *alpha_out, uint8_t *accepted_out) {
  const int kind = pf->kind, S = pf->S, T = pf->step;
  const int64_t n = pf->n;
  const int L = T - a + 1, nb = b - a + 1;
  if (kind == GO_MODEL_HMM || kind == GO_MODEL_BLR || a < 1 || b < a || b > T || L > pf->win_len) return -1;
  const uint32_t base = 0x900000u + 64u * move_index;
  int64_t total = 0;
#pragma omp parallel for schedule(static) reduction(+ : total) if (g_threads > 1)
  for (int64_t i = 0; i < n; i++) {
    double told[9][8], tnew[9][8];
    for (int j = 0; j <= L; j++) {
      const int lag = L - j;
      for (int s = 0; s < S; s++)
        told[j][s] = lag == 0 ? pf->state[(int64_t)s * n + i] : pf->win[((int64_t)(lag - 1) * S + s) * n + i];
    }
    for (int s = 0; s < S; s++) tnew[0][s] = told[0][s];
    double weight = 0.0, fwd = 0.0, bwd = 0.0;
    for (int j = 1; j <= L; j++) {
      const int t = a - 1 + j;
      const int *cidx;
      const int nc = model_latent_choices(kind, t, &cidx);
      double c_new[4], c_old[4];
      for (int k = 0; k < nc; k++) {
        c_old[k] = told[j][cidx[k]];
        if (j <= nb) {
          const int idx = (j - 1) * nc + k;
          double e;
          if (normals) e = normals[(int64_t)idx * n + i];
          else {
            double e0, e1;
            philox_normal_pair(pf->seed, (uint64_t)i, (uint32_t)T, base + (uint32_t)(idx >> 1), &e0, &e1);
            e = (idx & 1) ? e1 : e0;
          }
          c_new[k] = go_random_normal(c_old[k], drift_std, e);                   /* proposal: {addr} ~ normal(choices[addr], std) */
          fwd += go_logpdf_normal(c_new[k], c_old[k], drift_std);                /* propose: mh.jl:44 */
        } else {
          c_new[k] = c_old[k];
        }
      }
      model_state_from_choices(kind, t, tnew[j - 1], c_new, tnew[j]);
      for (int k = 0; k < nc; k++)                                               /* update.jl:46-50 */
        weight += model_latent_score(kind, pf->params, t, tnew[j - 1], k, c_new[k]) -
                  model_latent_score(kind, pf->params, t, told[j - 1], k, c_old[k]);
      weight += model_obs_logpdf(kind, pf->params, tnew[j], obs_window + (j - 1)) -
                model_obs_logpdf(kind, pf->params, told[j], obs_window + (j - 1));
      if (j <= nb)
        for (int k = 0; k < nc; k++) bwd += go_logpdf_normal(c_old[k], c_new[k], drift_std);   /* assess: mh.jl:48 */
    }
    const double alpha = weight - fwd + bwd;                                     /* mh.jl:49 */
    double ua;
    if (uniforms) ua = uniforms[i];
    else {
      double u0, u1;
      philox_uniform_pair(pf->seed, (uint64_t)i, (uint32_t)T, base + 63u, &u0, &u1);
      ua = u0;
    }
    const int acc = log(ua) < alpha;                                             /* mh.jl:52 */
    if (alpha_out) alpha_out[i] = alpha;
    if (accepted_out) accepted_out[i] = (uint8_t)acc;
    if (acc) {
      for (int j = 1; j <= L; j++) {
        const int lag = L - j;
        for (int s = 0; s < S; s++) {
          const double v = st_round(pf, tnew[j][s]);
          if (lag == 0) pf->state[(int64_t)s * n + i] = v;
          else pf->win[((int64_t)(lag - 1) * S + s) * n + i] = v;
        }
      }
      total++;
    }
  }
  pf->mh_ok = 0;
  return total;
}
/* particle_filter.jl:91-94 */
double go_pf_log_ml_estimate(const go_pf_t *pf) {
  return pf->log_ml_est + go_logsumexp(pf->logw, pf->n) - log((double)pf->n);
}
const double *go_pf_log_weights(const go_pf_t *pf) { return pf->logw; }
const int64_t *go_pf_parents(const go_pf_t *pf) { return pf->parents; }
const double *go_pf_state(const go_pf_t *pf, int f) { return pf->state + (int64_t)f * pf->n; }
double go_pf_log_ml_est_field(const go_pf_t *pf) { return pf->log_ml_est; }
int64_t go_pf_num_particles(const go_pf_t *pf) { return pf->n; }
int go_pf_step_index(const go_pf_t *pf) { return pf->step; }
void go_pf_set_log_weights(go_pf_t *pf, const double *lw) { memcpy(pf->logw, lw, (size_t)pf->n * 8); }
/* particle_filter.jl:43-51 */
go_pf_t *go_pf_copy(const go_pf_t *s) {
  go_pf_t *pf = (go_pf_t *)malloc(sizeof(*pf));
  *pf = *s;
  size_t sn = (size_t)s->S * s->n * 8, nn = (size_t)s->n * 8;
  pf->params = (double *)malloc((size_t)s->nparams * 8); memcpy(pf->params, s->params, (size_t)s->nparams * 8);
  pf->state = (double *)malloc(sn); memcpy(pf->state, s->state, sn);
  pf->new_state = (double *)malloc(sn); memcpy(pf->new_state, s->new_state, sn);
  pf->logw = (double *)malloc(nn); memcpy(pf->logw, s->logw, nn);
  pf->parents = (int64_t *)malloc(nn); memcpy(pf->parents, s->parents, nn);
  if (s->win) {
    const size_t wb = (size_t)s->win_W * sn;
    pf->win = (double *)malloc(wb); memcpy(pf->win, s->win, wb);
  }
  return pf;
}
void go_pf_free(go_pf_t *pf) {
  if (!pf) return;
  free(pf->win);
  free(pf->params); free(pf->state); free(pf->new_state); free(pf->logw); free(pf->parents); free(pf);
}
/* importance.jl:20-59 */
double go_importance_sampling(int kind, const double *params, int nparams, int64_t n, uint64_t seed,
                              const double *obs, int pk, const double *pa, const double *normals,
                              const double *uniforms, double *lognormw_out, go_pf_t **pf_out) {
  go_pf_t *pf = go_pf_init(kind, params, nparams, n, seed, obs, pk, pa, normals, uniforms);
  if (!pf) return NAN;
  double lse = go_logsumexp(pf->logw, n);                          /* :32, :55 */
  double lml = lse - log((double)n);                               /* :33, :56 */
  if (lognormw_out) for (int64_t i = 0; i < n; i++) lognormw_out[i] = pf->logw[i] - lse; /* :34, :57 */
  if (pf_out) *pf_out = pf; else go_pf_free(pf);
  return lml;
}
