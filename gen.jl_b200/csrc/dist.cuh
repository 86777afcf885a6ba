// dist.cuh -- device logpdf / random of Gen's built-in distributions on the hot path
// (src/modeling_library/distributions/{normal,gamma,bernoulli,categorical,mvnormal}.jl).
// Templated on the arithmetic type R (double = reference precision, float = fp32 mode).
#pragma once
#include "gbmath.cuh"

namespace gb {

template <typename R> struct Math;
template <> struct Math<double> {
  // hot-loop versions (csrc/gbmath.cuh): constant-bank coefficients + shared-memory tables
  static GB_D double log_t(double x, const MathTables &T) { return gb_log(x, T); }
  static GB_D double exp_t(double x) { return gb_exp(x); }
  static GB_D double atan2_t(double y, double x, const MathTables &T) { return gb_atan2(y, x, T); }
  static GB_D double log_(double x) { return log(x); }
  static GB_D double exp_(double x) { return exp(x); }
  static GB_D double sqrt_(double x) { return sqrt(x); }
  static GB_D double atan2_(double y, double x) { return atan2(y, x); }
  static GB_D double lgamma_(double x) { return lgamma(x); }
  static GB_D double pow_(double x, double y) { return pow(x, y); }
  static GB_D double neg_inf() { return -INFINITY; }
  static GB_D double log2pi() { return kLog2Pi; }
};
template <> struct Math<float> {
  static GB_D float log_t(float x, const MathTables &) { return logf(x); }
  static GB_D float exp_t(float x) { return expf(x); }
  static GB_D float atan2_t(float y, float x, const MathTables &) { return atan2f(y, x); }
  static GB_D float log_(float x) { return logf(x); }
  static GB_D float exp_(float x) { return expf(x); }
  static GB_D float sqrt_(float x) { return sqrtf(x); }
  static GB_D float atan2_(float y, float x) { return atan2f(y, x); }
  static GB_D float lgamma_(float x) { return lgammaf(x); }
  static GB_D float pow_(float x, float y) { return powf(x, y); }
  static GB_D float neg_inf() { return -INFINITY; }
  static GB_D float log2pi() { return (float)kLog2Pi; }
};

// normal.jl:56-59:  z = (x - mu) / std;  -(abs2(z) + log(2pi))/2 - log(std)
template <typename R> GB_D R logpdf_normal(R x, R mu, R std) {
  R z = (x - mu) / std;
  return -(z * z + Math<R>::log2pi()) / R(2) - Math<R>::log_(std);
}
// same with log(std) hoisted by the caller (std is a per-step constant on the hot path)
template <typename R> GB_D R logpdf_normal_ls(R x, R mu, R std, R log_std) {
  R z = (x - mu) / std;
  return -(z * z + Math<R>::log2pi()) / R(2) - log_std;
}
// hot-loop form: z = (x - mu) * (1/std) with the reciprocal hoisted by the caller.  Differs from the
// reference's division by <= 1 ulp of z, far inside the 1e-10 log-weight tolerance; the state arithmetic
// (random_normal) is untouched, so resampled states stay bit-exact.
template <typename R> GB_D R logpdf_normal_fast(R x, R mu, R inv_std, R log_std) {
  R z = (x - mu) * inv_std;
  return -(z * z + Math<R>::log2pi()) / R(2) - log_std;
}
// normal.jl:128:  mu + std * randn()   (two roundings: -fmad=false)
template <typename R> GB_D R random_normal(R mu, R std, R eps) { return mu + std * eps; }

// gamma.jl:10-16
template <typename R> GB_D R logpdf_gamma(R x, R shape, R scale) {
  if (x > R(0))
    return (shape - R(1)) * Math<R>::log_(x) - (x / scale) - shape * Math<R>::log_(scale) - Math<R>::lgamma_(shape);
  return Math<R>::neg_inf();
}
// bernoulli.jl:10-12 / :19
template <typename R> GB_D R logpdf_bernoulli(bool x, R p) { return x ? Math<R>::log_(p) : Math<R>::log_(R(1) - p); }
template <typename R> GB_D bool random_bernoulli(R p, R u) { return u < p; }
// categorical.jl:10-12 (x is 1-based); probs are always fp64 tables
template <typename R> GB_D R logpdf_categorical(long long x, const double *__restrict__ probs, int len) {
  return (x > 0 && x <= len) ? Math<R>::log_((R)__ldg(probs + (x - 1))) : Math<R>::neg_inf();
}
template <typename R> GB_D R logpdf_categorical_t(long long x, const double *__restrict__ probs, int len, const MathTables &T) {
  if (!(x > 0 && x <= len)) return Math<R>::neg_inf();
  const double p = __ldg(probs + (x - 1));
  return p > 0x1p-1000 ? Math<R>::log_t((R)p, T) : Math<R>::log_((R)p);   // log(0) = -Inf via libm
}
// categorical.jl:20-22: single draw = inverse-CDF linear scan (oracle go_random_categorical)
template <typename R> GB_D long long random_categorical(const double *__restrict__ probs, int len, R u) {
  double c = 0.0;
  for (int i = 0; i < len; i++) {
    c += __ldg(probs + i);
    if ((double)u < c) return i + 1;
  }
  return len;
}
// gamma.jl:29-31 -> Marsaglia-Tsang on the engine's Philox stream (oracle go_random_gamma_philox)
template <typename R> GB_D R random_gamma_philox(R shape, R scale, uint64_t seed, uint64_t pid, uint32_t step) {
  double sh = (double)shape;
  double a = sh < 1.0 ? sh + 1.0 : sh;
  double d = a - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d), g = d;
  for (uint32_t i = 0; i < 64; i++) {
    U4 w = philox_block(seed, pid, step, PK_GAMMA, 2 * i);
    U4 w2 = philox_block(seed, pid, step, PK_GAMMA, 2 * i + 1);
    double u1 = u52(w.x, w.y), u2 = u52(w.z, w.w);
    double x = sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
    double v = 1.0 + c * x;
    if (v <= 0.0) continue;
    v = v * v * v;
    double u = u52(w2.x, w2.y);
    if (log(u) < 0.5 * x * x + d - d * v + d * log(v)) { g = d * v; break; }
  }
  if (sh < 1.0) {
    U4 w = philox_block(seed, pid, step, PK_GAMMA, 0xFFFFFFu);
    g *= pow(u52(w.z, w.w), 1.0 / sh);
  }
  return (R)(g * (double)scale);
}

// ---- the remaining scalar built-ins (modeling_library/distributions/*.jl), fp64 ----
GB_D double logpdf_uniform(double x, double lo, double hi) {            // uniform_continuous.jl:12-14
  return (x >= lo && x <= hi) ? -log(hi - lo) : -INFINITY;
}
GB_D double logpdf_exponential(double x, double rate) {                 // exponential.jl:10-13 (Exponential(1/rate))
  return x >= 0.0 ? log(rate) - rate * x : -INFINITY;
}
GB_D double logpdf_laplace(double x, double loc, double scale) {        // laplace.jl:10-13
  const double diff = fabs(x - loc);
  return -diff / scale - log(2.0 * scale);
}
GB_D double logpdf_cauchy(double x, double x0, double gamma) {          // cauchy.jl:10-12 (Distributions.Cauchy)
  const double z = (x - x0) / gamma;
  return -log(M_PI) - log(gamma) - log1p(z * z);
}
GB_D double logpdf_inv_gamma(double x, double shape, double scale) {    // inv_gamma.jl:12-18
  if (x > 0.0) return shape * log(scale) - (shape + 1.0) * log(x) - lgamma(shape) - (scale / x);
  return -INFINITY;
}
GB_D double logpdf_beta(double x, double a, double b) {                 // beta.jl:13-24 incl. the endpoint special cases
  const double eps = 2.220446049250313e-16, lbeta = lgamma(a) + lgamma(b) - lgamma(a + b);
  if (x < 0.0 || x > 1.0) return -INFINITY;
  if (x == 0.0 && a <= 1.0) return (a - 1.0) * log(eps) - a + 1.0 + (b - 1.0) * log1p(-eps) - lbeta;
  if (x == 1.0 && b <= 1.0) return (a - 1.0) * log1p(-eps) - b + 1.0 + (b - 1.0) * log(eps) - lbeta;
  return (a - 1.0) * log(x) + (b - 1.0) * log1p(-x) - lbeta;
}
GB_D double logpdf_poisson(double xi, double lambda) {                  // poisson.jl:10-12
  return xi < 0.0 ? -INFINITY : xi * log(lambda) - lambda - lgamma(xi + 1.0);
}
GB_D double logpdf_geometric(double xi, double p) {                     // geometric.jl:10-12 (failures before 1st success)
  return xi < 0.0 ? -INFINITY : log(p) + xi * log1p(-p);
}
GB_D double logpdf_binom(double xi, double n, double p) {               // binom.jl:10-12
  if (xi < 0.0 || xi > n) return -INFINITY;
  return lgamma(n + 1.0) - lgamma(xi + 1.0) - lgamma(n - xi + 1.0) + xi * log(p) + (n - xi) * log1p(-p);
}
GB_D double logpdf_neg_binom(double xi, double r, double p) {           // neg_binom.jl:12-14 (failures before r successes)
  if (xi < 0.0) return -INFINITY;
  return lgamma(xi + r) - lgamma(xi + 1.0) - lgamma(r) + r * log(p) + xi * log1p(-p);
}
GB_D double logpdf_uniform_discrete(double xi, double lo, double hi) {  // uniform_discrete.jl:10-13
  return (xi >= lo && xi <= hi) ? -log(hi - lo + 1.0) : -INFINITY;
}

// distribution ids of the scalar built-ins above (== GB_DIST_* of include/genb200.h, static_assert in genb200.cu)
enum { DIST_UNIFORM = 6, DIST_EXPONENTIAL = 7, DIST_LAPLACE = 8, DIST_CAUCHY = 9, DIST_INV_GAMMA = 10, DIST_BETA = 11,
       DIST_POISSON = 12, DIST_GEOMETRIC = 13, DIST_BINOM = 14, DIST_NEG_BINOM = 15, DIST_UNIFORM_DISCRETE = 16,
       DIST_BROADCASTED_NORMAL = 17, DIST_DIRICHLET = 18, DIST_PIECEWISE_UNIFORM = 19, DIST_BETA_UNIFORM = 20 };
GB_D double logpdf_scalar(int dist, double x, double a, double b) {
  switch (dist) {
    case DIST_UNIFORM: return logpdf_uniform(x, a, b);
    case DIST_EXPONENTIAL: return logpdf_exponential(x, a);
    case DIST_LAPLACE: return logpdf_laplace(x, a, b);
    case DIST_CAUCHY: return logpdf_cauchy(x, a, b);
    case DIST_INV_GAMMA: return logpdf_inv_gamma(x, a, b);
    case DIST_BETA: return logpdf_beta(x, a, b);
    case DIST_POISSON: return logpdf_poisson(x, a);
    case DIST_GEOMETRIC: return logpdf_geometric(x, a);
    case DIST_BINOM: return logpdf_binom(x, a, b);
    case DIST_NEG_BINOM: return logpdf_neg_binom(x, a, b);
    case DIST_UNIFORM_DISCRETE: return logpdf_uniform_discrete(x, a, b);
  }
  return NAN;
}
// inverse-CDF draws on one uniform in (0,1) for the built-ins that have a closed-form quantile function
GB_D bool has_invcdf(int dist) {
  return dist == DIST_UNIFORM || dist == DIST_EXPONENTIAL || dist == DIST_LAPLACE || dist == DIST_CAUCHY || dist == DIST_GEOMETRIC ||
         dist == DIST_UNIFORM_DISCRETE;
}
GB_D double random_invcdf(int dist, double p0, double p1, double u) {
  switch (dist) {
    case DIST_UNIFORM: return p0 + (p1 - p0) * u;
    case DIST_EXPONENTIAL: return -log1p(-u) / p0;
    case DIST_LAPLACE: return u < 0.5 ? p0 + p1 * log(2.0 * u) : p0 - p1 * log(2.0 * (1.0 - u));
    case DIST_CAUCHY: return p0 + p1 * (sinpi(u - 0.5) / cospi(u - 0.5));
    case DIST_GEOMETRIC: return floor(log1p(-u) / log1p(-p0));
    case DIST_UNIFORM_DISCRETE: return p0 + floor((p1 - p0 + 1.0) * u);
  }
  return NAN;
}

// ---- dirichlet / piecewise_uniform / beta_uniform (the non-scalar-parameter built-ins) ----
GB_D double logsumexp2(double x1, double x2) {                           // inference.jl:8-11
  const double m = x1 > x2 ? x1 : x2;
  return m == -INFINITY ? m : m + log(exp(x1 - m) + exp(x2 - m));
}
GB_D double logpdf_dirichlet(const double *__restrict__ x, const double *__restrict__ alpha, int k) {   // dirichlet.jl:12-21
  double s = 0.0, sa = 0.0, lg = 0.0;
  bool ok = true;
  for (int j = 0; j < k; j++) { s += x[j]; ok = ok && x[j] >= 0.0 && alpha[j] >= 0.0; sa += alpha[j]; lg += lgamma(alpha[j]); }
  // isapprox(sum(x), 1): |s - 1| <= sqrt(eps) * max(|s|, 1)
  if (!ok || !(fabs(s - 1.0) <= 1.4901161193847656e-08 * fmax(fabs(s), 1.0))) return -INFINITY;
  double ll = 0.0;
  for (int j = 0; j < k; j++) ll += (alpha[j] - 1.0) * log(x[j]);
  return ll - (lg - lgamma(sa));
}
GB_D double logpdf_piecewise_uniform(double x, const double *__restrict__ bounds, const double *__restrict__ probs, int nb) {
  if (x <= bounds[0] || x >= bounds[nb]) return -INFINITY;               // piecewise_uniform.jl:38-44
  int bin = 0;
  while (x > bounds[bin + 1]) bin++;                                     // get_bin, :19-27
  return log(probs[bin]) - log(bounds[bin + 1] - bounds[bin]);
}
GB_D double logpdf_beta_uniform(double x, double theta, double a, double b) {   // beta_uniform.jl:10-18
  if (x < 0.0 || x > 1.0) return -INFINITY;
  return logsumexp2(log(theta) + logpdf_beta(x, a, b), log(1.0 - theta));
}
// Exact inversion for unimodal integer distributions: the uniform is compared against the probability masses in the order
// mode, mode+1, mode-1, mode+2, ... (each k still receives exactly p_k of the unit interval), so the search takes
// O(standard deviation) steps for any parameter.  up(k) = p_{k+1} / p_k, down(k) = p_{k-1} / p_k.
template <typename UP, typename DOWN>
GB_D double sample_unimodal(double u, long long mode, double p_mode, long long kmin, long long kmax, UP up, DOWN down) {
  if (u < p_mode) return (double)mode;
  u -= p_mode;
  double pu = p_mode, pd = p_mode;
  long long ku = mode, kd = mode;
  for (int it = 0; it < (1 << 22); it++) {
    bool moved = false;
    if (ku < kmax) { pu *= up(ku); ku++; moved = true; if (u < pu) return (double)ku; u -= pu; }
    if (kd > kmin) { pd *= down(kd); kd--; moved = true; if (u < pd) return (double)kd; u -= pd; }
    if (!moved) break;
  }
  return (double)ku;     // the rounding leftover of the unit interval (~1e-16) goes to the upper tail
}
GB_D double random_poisson(double lambda, double u) {                     // poisson.jl:19-21
  const long long mode = (long long)floor(lambda);
  const double pm = exp((double)mode * log(lambda) - lambda - lgamma((double)mode + 1.0));
  return sample_unimodal(u, mode, pm, 0, (1ll << 62), [=](long long k) { return lambda / (double)(k + 1); },
                         [=](long long k) { return (double)k / lambda; });
}
GB_D double random_binom(double n, double p, double u) {                  // binom.jl:18-20
  if (p <= 0.0) return 0.0;
  if (p >= 1.0) return n;
  long long mode = (long long)floor((n + 1.0) * p);
  if (mode > (long long)n) mode = (long long)n;
  const double pm = exp(logpdf_binom((double)mode, n, p));
  const double odds = p / (1.0 - p);
  return sample_unimodal(u, mode, pm, 0, (long long)n, [=](long long k) { return (n - (double)k) / ((double)k + 1.0) * odds; },
                         [=](long long k) { return (double)k / (n - (double)k + 1.0) / odds; });
}
GB_D double random_neg_binom(double r, double p, double u) {              // neg_binom.jl:22-24 (failures before r successes)
  const double q = 1.0 - p;
  long long mode = r > 1.0 ? (long long)floor((r - 1.0) * q / p) : 0;
  const double pm = exp(logpdf_neg_binom((double)mode, r, p));
  return sample_unimodal(u, mode, pm, 0, (1ll << 62), [=](long long k) { return ((double)k + r) / ((double)k + 1.0) * q; },
                         [=](long long k) { return (double)k / ((double)k - 1.0 + r) / q; });
}

// mvnormal.jl:12-16: the covariance is shared by all particles, so it is factored ONCE (host, in
// fp64) instead of once per call as the reference does; the kernel consumes L (row-major lower
// Cholesky factor, k*k doubles) and logdet:  -(k log 2pi + logdet + |L^-1 (x-mu)|^2) / 2
template <typename R, int KMAX>
GB_D R logpdf_mvnormal_chol(const R *x, const double *__restrict__ mu, const double *__restrict__ L, double logdet, int k) {
  R z[KMAX];
  R quad = R(0);
  for (int i = 0; i < k; i++) {
    R s = x[i] - (R)__ldg(mu + i);
    for (int p = 0; p < i; p++) s -= (R)__ldg(L + i * k + p) * z[p];
    z[i] = s / (R)__ldg(L + i * k + i);
    quad += z[i] * z[i];
  }
  return -(R(k) * Math<R>::log2pi() + (R)logdet + quad) / R(2);
}

}  // namespace gb
