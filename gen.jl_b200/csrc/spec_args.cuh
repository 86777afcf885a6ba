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

}  // namespace gb
