// models.cuh -- per-particle `generate` (t = 0) / extend-by-one `update` (t >= 1) of the supported
// model class, written against the reference's weight semantics:
//   * only constrained choices contribute to the weight (dynamic/generate.jl:26-41,
//     static_ir/generate.jl:31-42, unfold/update.jl:54-78);
//   * with a custom proposal the proposed latents are constrained as well and the proposal score is
//     subtracted once (particle_filter.jl:128-130, trace_translators.jl:854, importance.jl:53).
// The state arithmetic (propose + deterministic return value) uses exactly the operations of the
// Julia programs in SURVEY.md Appendix A so that resampled states are bit-exact in fp64.
#pragma once
#include "dist.cuh"

namespace gb {

enum { MODEL_LGSSM = 1, MODEL_SV = 2, MODEL_BEARINGS = 3, MODEL_HMM = 4, MODEL_BLR = 5, MODEL_SPEC = 6 };
enum { PROPOSAL_DEFAULT = 0, PROPOSAL_AFFINE_NORMAL = 1, PROPOSAL_CATEGORICAL_TABLE = 2 };

struct ModelParams {
  double p[12];        // scalar parameters (layout: include/genb200.h)
  double pa[4];        // scalar proposal arguments
  double obs0;         // the step's scalar observation
  const double *tab;   // DEVICE: HMM prior | transition | emission
  const double *ptab;  // DEVICE: HMM proposal table for this step
  int K, M;
  // GB_MODEL_SPEC programs lowered onto the step kernel at run time (spec_jit.inc): observations 1..15, step index
  double obsx[15];
  unsigned int step;
};

// Arithmetic type of the per-particle transition for a filter whose STORAGE type is R: always fp64.  A GB_F32 filter
// keeps traces, log weights and noise in fp32 (half the HBM traffic) but widens them for propose + logpdf, so its log
// weights differ from the fp64 filter's only through the rounding of what is stored: a single-precision atan2 or
// (obs - mean) / sigma residual is amplified by the likelihood's curvature (1/sigma = 200 for the bearings model) to
// 1e-4 relative per particle, ten times the 1e-5 bar.  The kernels are HBM / issue bound, not fp64-pipe bound.
template <typename R> struct Arith { typedef double type; };

// Per-thread constants hoisted out of the particle loop (log of fixed standard deviations etc.).
template <typename R> struct Hoisted { R c0, c1, c2, i0, i1, i2; };

template <typename R, int KIND, int PK, bool INIT> struct Transition;

// ---- cfg 1: 1-D linear-Gaussian state-space model ----
template <typename R, int PK, bool INIT> struct Transition<R, MODEL_LGSSM, PK, INIT> {
  static constexpr int S = 1, NN = 1, NU = 0;
  static GB_D Hoisted<R> hoist(const ModelParams &mp) {
    Hoisted<R> h;
    h.c0 = Math<R>::log_((R)mp.p[4]);                  // log r
    h.c1 = Math<R>::log_((R)(INIT ? mp.p[1] : mp.p[3])); // log model std of x
    h.c2 = Math<R>::log_((R)mp.pa[2]);                 // log proposal std
    h.i0 = R(1) / (R)mp.p[4];
    h.i1 = R(1) / (R)(INIT ? mp.p[1] : mp.p[3]);
    h.i2 = R(1) / (R)mp.pa[2];
    return h;
  }
  static GB_D R apply(const ModelParams &mp, const Hoisted<R> &h, const MathTables &T, const R *prev, R *next, const R *eps, const R *) {
    R mu = INIT ? (R)mp.p[0] : (R)mp.p[2] * prev[0];
    R sd = INIT ? (R)mp.p[1] : (R)mp.p[3];
    if (PK == PROPOSAL_AFFINE_NORMAL) {
      R qmu = INIT ? (R)mp.pa[0] : (R)mp.pa[0] + (R)mp.pa[1] * prev[0];
      R x = random_normal<R>(qmu, (R)mp.pa[2], eps[0]);
      R lq = logpdf_normal_fast<R>(x, qmu, h.i2, h.c2);
      R w = R(0);
      w += logpdf_normal_fast<R>(x, mu, h.i1, h.c1);
      w += logpdf_normal_fast<R>((R)mp.obs0, x, h.i0, h.c0);
      next[0] = x;
      return w - lq;
    }
    R x = random_normal<R>(mu, sd, eps[0]);
    next[0] = x;
    return obs_logpdf(mp, h, T, next);
  }
  // scores used by the MH rejuvenation moves (kernels_mh.cuh): the observed choice given the state, and the
  // latent choices of the latest kernel application given the parent state; `drift` perturbs those choices
  static constexpr int ND = 1;
  static GB_D R obs_logpdf(const ModelParams &mp, const Hoisted<R> &h, const MathTables &, const R *st) {
    return logpdf_normal_fast<R>((R)mp.obs0, st[0], h.i0, h.c0);
  }
  static GB_D R latent_logpdf(const ModelParams &mp, const Hoisted<R> &h, const MathTables &, const R *prev, const R *st) {
    R mu = INIT ? (R)mp.p[0] : (R)mp.p[2] * prev[0];
    return logpdf_normal_fast<R>(st[0], mu, h.i1, h.c1);
  }
  static GB_D R drift(const ModelParams &, const Hoisted<R> &, const R *, const R *st, R *next, const R *eps, R sd, R lsd) {
    next[0] = random_normal<R>(st[0], sd, eps[0]);
    return logpdf_normal_ls<R>(next[0], st[0], sd, lsd) - logpdf_normal_ls<R>(st[0], next[0], sd, lsd);   // fwd - bwd
  }
};

// ---- cfg 2: stochastic volatility ----
template <typename R, int PK, bool INIT> struct Transition<R, MODEL_SV, PK, INIT> {
  static constexpr int S = 1, NN = 1, NU = 0;
  static GB_D Hoisted<R> hoist(const ModelParams &) { return Hoisted<R>{R(0), R(0), R(0), R(0), R(0), R(0)}; }
  static GB_D R apply(const ModelParams &mp, const Hoisted<R> &h, const MathTables &T, const R *prev, R *next, const R *eps, const R *) {
    R mean = INIT ? (R)mp.p[0] : (R)mp.p[0] + (R)mp.p[1] * (prev[0] - (R)mp.p[0]);
    R sd = INIT ? (R)mp.p[3] : (R)mp.p[2];
    R hh = random_normal<R>(mean, sd, eps[0]);
    next[0] = hh;
    return obs_logpdf(mp, h, T, next);
  }
  static constexpr int ND = 1;
  static GB_D R obs_logpdf(const ModelParams &mp, const Hoisted<R> &, const MathTables &, const R *st) {
    // y ~ normal(0, exp(h/2)):  log(std) = h/2 and 1/std = exp(-h/2) (the reference evaluates
    // log(exp(h/2)) and divides; both agree to ~1 ulp, far inside the 1e-10 weight tolerance)
    R half = st[0] / R(2);
    return logpdf_normal_fast<R>((R)mp.obs0, R(0), Math<R>::exp_t(-half), half);
  }
  static GB_D R latent_logpdf(const ModelParams &mp, const Hoisted<R> &, const MathTables &, const R *prev, const R *st) {
    R mean = INIT ? (R)mp.p[0] : (R)mp.p[0] + (R)mp.p[1] * (prev[0] - (R)mp.p[0]);
    return logpdf_normal<R>(st[0], mean, INIT ? (R)mp.p[3] : (R)mp.p[2]);
  }
  static GB_D R drift(const ModelParams &, const Hoisted<R> &, const R *, const R *st, R *next, const R *eps, R sd, R lsd) {
    next[0] = random_normal<R>(st[0], sd, eps[0]);
    return logpdf_normal_ls<R>(next[0], st[0], sd, lsd) - logpdf_normal_ls<R>(st[0], next[0], sd, lsd);
  }
};

// ---- cfg 3: bearings-only tracking (docs/src/tutorials/smc.md:607-661) ----
template <typename R, int PK, bool INIT> struct Transition<R, MODEL_BEARINGS, PK, INIT> {
  static constexpr int S = 4, NN = INIT ? 4 : 2, NU = 0;
  static GB_D Hoisted<R> hoist(const ModelParams &mp) {
    Hoisted<R> h;
    h.c0 = Math<R>::log_((R)mp.p[9]);      // log measurement_noise
    h.c1 = Math<R>::sqrt_((R)mp.p[8]);     // sqrt(velocity_var)
    h.c2 = R(0);
    h.i0 = R(1) / (R)mp.p[9];
    h.i1 = h.i2 = R(0);
    return h;
  }
  static GB_D R apply(const ModelParams &mp, const Hoisted<R> &h, const MathTables &T, const R *prev, R *next, const R *eps, const R *) {
    R x, y, vx, vy;
    if (INIT) {
      x = random_normal<R>((R)mp.p[0], (R)mp.p[1], eps[0]);
      y = random_normal<R>((R)mp.p[2], (R)mp.p[3], eps[1]);
      vx = random_normal<R>((R)mp.p[4], (R)mp.p[5], eps[2]);
      vy = random_normal<R>((R)mp.p[6], (R)mp.p[7], eps[3]);
    } else {
      vx = random_normal<R>(prev[2], h.c1, eps[0]);
      vy = random_normal<R>(prev[3], h.c1, eps[1]);
      x = prev[0] + vx;
      y = prev[1] + vy;
    }
    next[0] = x; next[1] = y; next[2] = vx; next[3] = vy;
    return obs_logpdf(mp, h, T, next);
  }
  static constexpr int ND = INIT ? 4 : 2;
  static GB_D R obs_logpdf(const ModelParams &mp, const Hoisted<R> &h, const MathTables &T, const R *st) {
    return logpdf_normal_fast<R>((R)mp.obs0, Math<R>::atan2_t(st[1], st[0], T), h.i0, h.c0);
  }
  static GB_D R latent_logpdf(const ModelParams &mp, const Hoisted<R> &h, const MathTables &, const R *prev, const R *st) {
    R w = R(0);
    if (INIT) {
      for (int k = 0; k < 4; k++) w += logpdf_normal<R>(st[k], (R)mp.p[2 * k], (R)mp.p[2 * k + 1]);
    } else {
      w += logpdf_normal<R>(st[2], prev[2], h.c1);
      w += logpdf_normal<R>(st[3], prev[3], h.c1);
    }
    return w;
  }
  // the choices are (x0, y0, vx0, vy0) at t = 0 and (vx, vy) afterwards; positions follow deterministically
  static GB_D R drift(const ModelParams &, const Hoisted<R> &, const R *prev, const R *st, R *next, const R *eps, R sd, R lsd) {
    R d = R(0);
    if (INIT) {
      for (int k = 0; k < 4; k++) {
        next[k] = random_normal<R>(st[k], sd, eps[k]);
        d += logpdf_normal_ls<R>(next[k], st[k], sd, lsd) - logpdf_normal_ls<R>(st[k], next[k], sd, lsd);
      }
    } else {
      next[2] = random_normal<R>(st[2], sd, eps[0]);
      next[3] = random_normal<R>(st[3], sd, eps[1]);
      next[0] = prev[0] + next[2];
      next[1] = prev[1] + next[3];
      for (int k = 2; k < 4; k++) d += logpdf_normal_ls<R>(next[k], st[k], sd, lsd) - logpdf_normal_ls<R>(st[k], next[k], sd, lsd);
    }
    return d;
  }
};

// ---- the reference's own PF test fixture: 3-state HMM (test/inference/particle_filter.jl:52-78) ----
template <typename R, int PK, bool INIT> struct Transition<R, MODEL_HMM, PK, INIT> {
  static constexpr int S = 1, NN = 0, NU = 1;
  static GB_D Hoisted<R> hoist(const ModelParams &) { return Hoisted<R>{R(0), R(0), R(0), R(0), R(0), R(0)}; }
  static GB_D R apply(const ModelParams &mp, const Hoisted<R> &, const MathTables &T, const R *prev, R *next, const R *, const R *us) {
    const int K = mp.K, M = mp.M;
    const double *prior = mp.tab, *trans = mp.tab + K, *emis = mp.tab + K + K * K;
    const double *zp = INIT ? prior : trans + ((long long)prev[0] - 1) * K;
    long long xo = (long long)mp.obs0;
    if (PK == PROPOSAL_CATEGORICAL_TABLE) {
      const double *qp = INIT ? mp.ptab : mp.ptab + ((long long)prev[0] - 1) * K;
      long long z = random_categorical<R>(qp, K, us[0]);
      R lq = logpdf_categorical_t<R>(z, qp, K, T);
      R w = R(0);
      w += logpdf_categorical_t<R>(z, zp, K, T);
      w += logpdf_categorical_t<R>(xo, emis + (z - 1) * M, M, T);
      next[0] = (R)z;
      return w - lq;
    }
    long long z = random_categorical<R>(zp, K, us[0]);
    next[0] = (R)z;
    return logpdf_categorical_t<R>(xo, emis + (z - 1) * M, M, T);
  }
  static constexpr int ND = 0;   // discrete latent: resimulation move only
  static GB_D R obs_logpdf(const ModelParams &mp, const Hoisted<R> &, const MathTables &T, const R *st) {
    const int K = mp.K, M = mp.M;
    return logpdf_categorical_t<R>((long long)mp.obs0, mp.tab + K + K * K + ((long long)st[0] - 1) * M, M, T);
  }
  static GB_D R latent_logpdf(const ModelParams &mp, const Hoisted<R> &, const MathTables &T, const R *prev, const R *st) {
    const int K = mp.K;
    return logpdf_categorical_t<R>((long long)st[0], INIT ? mp.tab : mp.tab + K + ((long long)prev[0] - 1) * K, K, T);
  }
  static GB_D R drift(const ModelParams &, const Hoisted<R> &, const R *, const R *st, R *next, const R *, R, R) {
    next[0] = st[0];
    return R(0);
  }
};

}  // namespace gb
