/* examples/pf_multigpu.c -- ONE particle filter sharded over several GPUs from plain C, one host thread, no NCCL / MPI /
 * Python: the gb_spf_* entries of include/genb200.h.
 *
 * Bearings-only tracking (docs/src/tutorials/smc.md:607-661), N particles over G shards.  The same loop is run on one
 * shard and on G shards -- through the synchronous calls (maybe_resample! returns its Bool) and through gb_spf_run
 * (every decision and exchange on the devices) -- and the results must be IDENTICAL: Philox noise is keyed by the global
 * particle id and maybe_resample! is global-prefix systematic resampling over an exact integer CDF.
 *
 *   gcc -O2 -Iinclude examples/pf_multigpu.c -o pf_multigpu -Lgen.jl_b200 -lgenb200 -Wl,-rpath,$PWD/gen.jl_b200 -lm
 *   ./pf_multigpu [N] [T] [G]      (G shards on devices 0..G-1; with fewer than G GPUs the shards share devices)
 * Exit code 0: all three runs agree bit for bit; 2: no GPU.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "genb200.h"

#define CHECK(call)                                                          \
  do {                                                                       \
    int rc__ = (call);                                                       \
    if (rc__ != GB_OK) {                                                     \
      fprintf(stderr, "%s -> %d: %s\n", #call, rc__, gb_last_error());       \
      return rc__ == GB_ENOGPU ? 2 : 1;                                      \
    }                                                                        \
  } while (0)

int main(int argc, char **argv) {
  const int64_t N = argc > 1 ? atoll(argv[1]) : 200000;
  const int T = argc > 2 ? atoi(argv[2]) : 12;
  const int G = argc > 3 ? atoi(argv[3]) : 2;
  int ndev = 0;
  CHECK(gb_device_count(&ndev));
  if (ndev == 0) { fprintf(stderr, "no CUDA device: libgenb200 has no CPU fallback\n"); return 2; }
  int devices[64];
  for (int r = 0; r < G && r < 64; r++) devices[r] = r % ndev;
  const int one[1] = {0};

  const double params[10] = {0.01, 0.01, 0.95, 0.01, 0.002, 0.01, -0.013, 0.01, 1e-6, 0.005};
  double *z = (double *)malloc(sizeof(double) * (T + 1));
  for (int t = 0; t <= T; t++) z[t] = atan2(0.95 - 0.013 * t, 0.01 + 0.002 * t) + 0.004 * sin(1.7 * t);
  gb_model_t *model = NULL;
  CHECK(gb_model_create(GB_MODEL_BEARINGS, params, 10, &model));

  double lml[3] = {0, 0, 0};
  int nres[3] = {0, 0, 0};
  double *x[3] = {NULL, NULL, NULL};
  for (int run = 0; run < 3; run++) {            /* 0: one shard, synchronous; 1: G shards, synchronous; 2: G shards, gb_spf_run */
    gb_spf_t *spf = NULL;
    CHECK(gb_spf_initialize(model, N, GB_F64, 20242, run == 0 ? 1 : G, run == 0 ? one : devices, &z[0], 1, GB_PROPOSAL_DEFAULT, NULL, 0, &spf));
    if (run < 2) {
      for (int t = 1; t <= T; t++) {
        int did = 0;
        CHECK(gb_spf_maybe_resample(spf, (double)N / 2.0, &did));
        nres[run] += did;
        CHECK(gb_spf_step(spf, &z[t], 1, GB_PROPOSAL_DEFAULT, NULL, 0));
      }
    } else {
      CHECK(gb_spf_run(spf, T, &z[1], 1, (double)N / 2.0, &nres[run]));
    }
    CHECK(gb_spf_log_ml_estimate(spf, &lml[run]));
    x[run] = (double *)malloc(sizeof(double) * N);
    CHECK(gb_spf_get_state(spf, 0, x[run]));
    gb_spf_free(spf);
  }
  gb_model_free(model);
  const int same_state = memcmp(x[0], x[1], sizeof(double) * N) == 0 && memcmp(x[0], x[2], sizeof(double) * N) == 0;
  const int ok = same_state && nres[0] == nres[1] && nres[0] == nres[2] && nres[0] > 0 &&
                 fabs(lml[0] - lml[1]) <= 1e-10 * fabs(lml[0]) && fabs(lml[0] - lml[2]) <= 1e-10 * fabs(lml[0]);
  printf("{\"N\": %lld, \"T\": %d, \"shards\": %d, \"devices\": %d, \"log_ml\": [%.12f, %.12f, %.12f], \"resamples\": [%d, %d, %d], "
         "\"traces_identical\": %s, \"ok\": %s}\n", (long long)N, T, G, ndev, lml[0], lml[1], lml[2], nres[0], nres[1], nres[2],
         same_state ? "true" : "false", ok ? "true" : "false");
  for (int k = 0; k < 3; k++) free(x[k]);
  free(z);
  return ok ? 0 : 1;
}
