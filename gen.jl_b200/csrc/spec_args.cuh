// spec_args.cuh -- launch arguments shared by the GB_MODEL_SPEC interpreter (kernels_spec.cuh) and the kernels
// NVRTC specialises from the same programs at run time (spec_jit.inc).
#pragma once
#include "kernels_step.cuh"

namespace gb {

constexpr int kSpecRegs = 32, kSpecMaxObs = 64;

struct SpecArgs {
  const void *state_in;
  void *state_out;
  void *logw, *incr;
  const int *anc;
  const RunCtl *ctl;
  const double *normals, *uniforms;   // PREDRAWN [nd][ld]
  long long n, ld, pid_offset;
  unsigned long long seed;
  unsigned int step;
  int op_begin, op_end;               // interpreter: slice of c_spec_ops to run (generate / step program)
  double *partials;
  unsigned int *ticket;
  WeightStats *stats;
  double obs[kSpecMaxObs];            // this call's observations
};

// A categorical table's offset into the parameter block is a RUN-TIME register value: [off, off + len) must lie inside the
// block (np entries), otherwise the choice is poisoned (NaN value / NaN weight) instead of reading out of bounds.
GB_D int spec_table_offset(double v, int len, int np) {
  const int o = (int)v;
  return (v == v && o >= 0 && o + len <= np) ? o : -1;
}

constexpr int kSpecMaxMv = 8;
// mvnormal.jl:30-33: mu + L * randn(k) with the Cholesky factor prepared at model creation
template <typename R> GB_D void spec_mvnormal_sample(R *x, const R *mu, const double *__restrict__ Lf, const R *eps, int k) {
  for (int i = 0; i < k; i++) {
    R s = mu[i];
    for (int c = 0; c <= i; c++) s += (R)Lf[i * k + c] * eps[c];
    x[i] = s;
  }
}
// mvnormal.jl:12-16: -(k log 2pi + logdet + |L^-1 (x - mu)|^2) / 2
template <typename R> GB_D R spec_mvnormal_logpdf(const R *x, const R *mu, const double *__restrict__ Lf, int k) {
  R z[kSpecMaxMv];
  R quad = R(0);
  for (int i = 0; i < k; i++) {
    R s = x[i] - mu[i];
    for (int p = 0; p < i; p++) s -= (R)Lf[i * k + p] * z[p];
    z[i] = s / (R)Lf[i * k + i];
    quad += z[i] * z[i];
  }
  return -(R(k) * Math<R>::log2pi() + (R)Lf[k * k] + quad) / R(2);
}

}  // namespace gb
