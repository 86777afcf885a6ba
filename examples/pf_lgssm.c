/* examples/pf_lgssm.c -- the C ABI used from plain C (no Python, no torch): BASELINE config 1.
 *
 * A bootstrap particle filter on the 1-D linear-Gaussian state-space model
 *     x_0 ~ N(m0, s0), x_t ~ N(a x_{t-1}, q), y_t ~ N(x_t, r)
 * written exactly like the reference's user loop (docs/src/tutorials/smc.md:674-689):
 *     state = initialize_particle_filter(model, (0,), obs_0, N)
 *     for t = 1:T   maybe_resample!(state, ess_threshold=N/2); particle_filter_step!(state, (t,), ..., obs_t)   end
 *     log_ml_estimate(state)
 * and checked against the exact Kalman-filter log marginal likelihood computed below.
 *
 *   gcc -O2 -Iinclude examples/pf_lgssm.c -o pf_lgssm -Lgen.jl_b200 -lgenb200 -Wl,-rpath,$PWD/gen.jl_b200 -lm
 *   ./pf_lgssm [N] [T] [seeds]
 * Exit code 0: |mean_seeds(log ML) - Kalman| within 5 standard errors (+ Jensen bias allowance); 2: no GPU.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "genb200.h"

static const double M0 = 0.0, S0 = 1.0, A = 0.9, Q = 0.5, R = 1.0;

/* splitmix64 + Box-Muller: just a data generator for the synthetic series */
static uint64_t sm_state = 20240;
static double unif(void) {
  uint64_t z = (sm_state += 0x9E3779B97F4A7C15ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  return ((double)(z >> 11) + 0.5) * (1.0 / 9007199254740992.0);
}
static double randn(void) { return sqrt(-2.0 * log(unif())) * cos(6.283185307179586 * unif()); }

/* exact log p(y_0..y_T) by the scalar Kalman filter */
static double kalman_logml(const double *y, int T1) {
  double m = M0, P = S0 * S0, ll = 0.0;
  for (int t = 0; t < T1; t++) {
    if (t > 0) { m = A * m; P = A * A * P + Q * Q; }
    double S = P + R * R, e = y[t] - m, K = P / S;
    ll += -0.5 * (log(2.0 * M_PI * S) + e * e / S);
    m += K * e;
    P *= (1.0 - K);
  }
  return ll;
}

#define CHECK(call)                                                          \
  do {                                                                       \
    int rc__ = (call);                                                       \
    if (rc__ != GB_OK) {                                                     \
      fprintf(stderr, "%s -> %d: %s\n", #call, rc__, gb_last_error());       \
      return rc__ == GB_ENOGPU ? 2 : 1;                                      \
    }                                                                        \
  } while (0)

int main(int argc, char **argv) {
  const int64_t N = argc > 1 ? atoll(argv[1]) : 1000;
  const int T = argc > 2 ? atoi(argv[2]) : 100;
  const int seeds = argc > 3 ? atoi(argv[3]) : 32;
  double *y = (double *)malloc(sizeof(double) * (T + 1));
  double x = M0 + S0 * randn();
  for (int t = 0; t <= T; t++) {
    if (t > 0) x = A * x + Q * randn();
    y[t] = x + R * randn();
  }
  const double exact = kalman_logml(y, T + 1);

  const double params[5] = {M0, S0, A, Q, R};
  gb_model_t *model = NULL;
  CHECK(gb_model_create(GB_MODEL_LGSSM, params, 5, &model));
  double sum = 0.0, sum2 = 0.0;
  int resamples = 0;
  for (int s = 0; s < seeds; s++) {
    gb_pf_t *pf = NULL;
    CHECK(gb_pf_initialize(model, N, GB_F64, (uint64_t)s, &y[0], 1, GB_PROPOSAL_DEFAULT, NULL, 0, &pf));
    for (int t = 1; t <= T; t++) {
      int did = 0;
      CHECK(gb_pf_maybe_resample(pf, (double)N / 2.0, GB_RESAMPLE_SYSTEMATIC, &did));
      resamples += did;
      CHECK(gb_pf_step(pf, &y[t], 1, GB_PROPOSAL_DEFAULT, NULL, 0, NULL));
    }
    double lml = 0.0;
    CHECK(gb_pf_log_ml_estimate(pf, &lml));
    sum += lml;
    sum2 += lml * lml;
    gb_pf_free(pf);
  }
  gb_model_free(model);
  const double mean = sum / seeds, var = seeds > 1 ? (sum2 - seeds * mean * mean) / (seeds - 1) : 0.0;
  const double se = sqrt(var / seeds);
  /* E[log Z_hat] <= log Z (Jensen): bias ~ -var/2 */
  const int ok = fabs(mean - exact) <= 5.0 * se + 0.5 * var + 1e-9;
  printf("{\"N\": %lld, \"T\": %d, \"seeds\": %d, \"kalman_log_ml\": %.10f, \"pf_mean_log_ml\": %.10f, \"pf_sd\": %.4f, "
         "\"resamples\": %d, \"ok\": %s}\n", (long long)N, T, seeds, exact, mean, sqrt(var), resamples, ok ? "true" : "false");
  free(y);
  return ok ? 0 : 1;
}
