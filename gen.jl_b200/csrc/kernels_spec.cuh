// kernels_spec.cuh -- GB_MODEL_SPEC: a declarative model specification for the supported class.
//
// Gen's static-IR functions are DAGs whose nodes are random choices of built-in distributions and Julia
// expressions (src/static_ir/dag.jl:1-46); the Julia expressions are opaque closures (dag.jl:14-19,
// dsl/static.jl:116-129), so the DAG cannot be lowered automatically.  What crosses the boundary instead is
// the same DAG written out as a tiny SSA program (one for `generate` at t = 0, one for the Unfold kernel's
// extend-by-one `update`): arithmetic nodes, `sample` nodes (unconstrained choices: value drawn, weight
// unchanged -- static_ir/generate.jl:31-38) and `observe` nodes (constrained choices: weight += logpdf --
// unfold/update.jl:64-68).  One generic kernel interprets it per particle: every thread executes the same
// op (uniform control flow, program and parameters in constant memory), registers live in local memory;
// the transcendental work dominates exactly as in the specialised kernels.
#pragma once
#include "spec_args.cuh"

namespace gb {

enum {
  OP_CONST = 0, OP_PREV = 1, OP_PARAM = 2, OP_OBS = 3, OP_TIME = 4,
  OP_ADD = 10, OP_SUB = 11, OP_MUL = 12, OP_DIV = 13, OP_NEG = 14, OP_SQRT = 15, OP_EXP = 16, OP_LOG = 17, OP_ATAN2 = 18,
  OP_ABS = 19,
  OP_SAMPLE_NORMAL = 20, OP_OBSERVE_NORMAL = 21, OP_SAMPLE_BERNOULLI = 22, OP_OBSERVE_BERNOULLI = 23,
  OP_SAMPLE_GAMMA = 24, OP_OBSERVE_GAMMA = 25, OP_SAMPLE_CATEGORICAL = 26, OP_OBSERVE_CATEGORICAL = 27,
  OP_STORE = 30, OP_WEIGHT = 31, OP_LOGPDF_NORMAL = 32,
  OP_SAMPLE_DIST = 40, OP_OBSERVE_DIST = 41, OP_LOGPDF_DIST = 42,  // the other scalar built-ins: imm = distribution id (+ 100 * x register)
  OP_SAMPLE_MVNORMAL = 43, OP_OBSERVE_MVNORMAL = 44                // k = (int)imm consecutive registers; b = offset of [L (k*k, row-major), logdet]
};
constexpr int kSpecMaxOps = 256, kSpecMaxParams = 2048, kSpecMaxState = 16;

struct SpecOp { int code, dst, a, b; double imm; };
__constant__ SpecOp c_spec_ops[kSpecMaxOps];
__constant__ double c_spec_params[kSpecMaxParams];


template <typename ST, bool INIT, bool PREDRAWN, int GATHER>
__global__ void __launch_bounds__(kBlock) spec_kernel(const SpecArgs a) {
  typedef typename Arith<ST>::type R;     // ST = storage type of the filter; the program is evaluated in fp64 (models.cuh)
  __shared__ double s_tab[kMathTabDoubles];
  pdl_trigger();
  const MathTables T = stage_math_tables(s_tab);
  __syncthreads();
  pdl_wait();
  if (!INIT && GATHER != 0) xchg_wait_done(a.stats);
  const ST *__restrict__ sin = static_cast<const ST *>(a.state_in);
  ST *__restrict__ sout = static_cast<ST *>(a.state_out);
  ST *__restrict__ logw = static_cast<ST *>(a.logw);
  ST *__restrict__ incr = static_cast<ST *>(a.incr);
  const bool gather = GATHER == 2 ? (a.ctl->do_resample != 0) : (GATHER == 1);
  double acc = -INFINITY;
  for (long long i = (long long)blockIdx.x * kBlock + threadIdx.x; i < a.n; i += (long long)gridDim.x * kBlock) {
    const long long src = (!INIT && gather) ? (long long)a.anc[i] : i;
    const unsigned long long pid = (unsigned long long)(a.pid_offset + i);
    R regs[kSpecRegs];
    R w = R(0);
    int n_norm = 0, n_unif = 0, n_gamma = 0;
    ST spare_n = ST(0), spare_u = ST(0);
    for (int k = a.op_begin; k < a.op_end; k++) {
      const SpecOp op = c_spec_ops[k];
      switch (op.code) {
        case OP_CONST: regs[op.dst] = (R)op.imm; break;
        case OP_PREV: regs[op.dst] = INIT ? R(0) : (R)__ldg(sin + (long long)op.a * a.ld + src); break;
        case OP_PARAM: regs[op.dst] = (R)c_spec_params[op.a]; break;
        case OP_OBS: regs[op.dst] = (R)a.obs[op.a]; break;
        case OP_TIME: regs[op.dst] = (R)a.step; break;
        case OP_ADD: regs[op.dst] = regs[op.a] + regs[op.b]; break;
        case OP_SUB: regs[op.dst] = regs[op.a] - regs[op.b]; break;
        case OP_MUL: regs[op.dst] = regs[op.a] * regs[op.b]; break;
        case OP_DIV: regs[op.dst] = regs[op.a] / regs[op.b]; break;
        case OP_NEG: regs[op.dst] = -regs[op.a]; break;
        case OP_ABS: regs[op.dst] = regs[op.a] < R(0) ? -regs[op.a] : regs[op.a]; break;
        case OP_SQRT: regs[op.dst] = Math<R>::sqrt_(regs[op.a]); break;
        case OP_EXP: regs[op.dst] = Math<R>::exp_t(regs[op.a]); break;
        case OP_LOG: regs[op.dst] = Math<R>::log_(regs[op.a]); break;
        case OP_ATAN2: regs[op.dst] = Math<R>::atan2_t(regs[op.a], regs[op.b], T); break;
        case OP_SAMPLE_NORMAL: {      // normal.jl:128: mu + std * randn()
          ST e;                        // noise in the storage precision
          if (PREDRAWN) e = (ST)__ldcs(a.normals + (long long)n_norm * a.ld + i);
          else if (n_norm & 1) e = spare_n;
          else philox_normal_pair(a.seed, pid, a.step, (unsigned)(n_norm >> 1), T, e, spare_n);
          n_norm++;
          regs[op.dst] = random_normal<R>(regs[op.a], regs[op.b], (R)e);
          break;
        }
        case OP_OBSERVE_NORMAL: w += logpdf_normal<R>(regs[op.dst], regs[op.a], regs[op.b]); break;   // normal.jl:56-59
        case OP_SAMPLE_BERNOULLI: case OP_SAMPLE_CATEGORICAL: case OP_SAMPLE_DIST: {
          ST u;
          if (PREDRAWN) u = (ST)__ldcs(a.uniforms + (long long)n_unif * a.ld + i);
          else if (n_unif & 1) u = spare_u;
          else philox_uniform_pair(a.seed, pid, a.step, (unsigned)(n_unif >> 1), u, spare_u);
          n_unif++;
          if (op.code == OP_SAMPLE_DIST) regs[op.dst] = (R)random_invcdf((int)op.imm, (double)regs[op.a], (double)regs[op.b], (double)u);
          else if (op.code == OP_SAMPLE_BERNOULLI) regs[op.dst] = random_bernoulli<R>(regs[op.a], (R)u) ? R(1) : R(0);   // bernoulli.jl:19
          else {                                                                                                     // categorical.jl:20-22
            const int off = spec_table_offset((double)regs[op.a], op.b, kSpecMaxParams);
            regs[op.dst] = off < 0 ? (R)NAN : (R)random_categorical<R>(c_spec_params + off, op.b, (R)u);
          }
          break;
        }
        case OP_OBSERVE_BERNOULLI: w += logpdf_bernoulli<R>(regs[op.dst] != R(0), regs[op.a]); break;             // bernoulli.jl:10-12
        case OP_OBSERVE_CATEGORICAL: {                                                                            // categorical.jl:10-12
          const long long x = (long long)regs[op.dst];
          const int off = spec_table_offset((double)regs[op.a], op.b, kSpecMaxParams);
          if (off < 0) w += (R)NAN;
          else w += (x > 0 && x <= op.b) ? Math<R>::log_((R)c_spec_params[off + (int)x - 1]) : Math<R>::neg_inf();
          break;
        }
        case OP_SAMPLE_GAMMA:                                                                                     // gamma.jl:29-31
          regs[op.dst] = random_gamma_philox<R>(regs[op.a], regs[op.b], a.seed ^ (0x632BE59BD9B4E019ull * (unsigned long long)(n_gamma + 1)),
                                               pid, a.step);
          n_gamma++;
          break;
        case OP_OBSERVE_GAMMA: w += logpdf_gamma<R>(regs[op.dst], regs[op.a], regs[op.b]); break;                 // gamma.jl:10-16
        case OP_STORE: st_stream(sout + (long long)op.dst * a.ld + i, (ST)regs[op.a]); break;
        case OP_WEIGHT: w += regs[op.a]; break;                                     // e.g. minus a proposal score
        case OP_LOGPDF_NORMAL: regs[op.dst] = logpdf_normal<R>(regs[(int)op.imm], regs[op.a], regs[op.b]); break;
        case OP_SAMPLE_MVNORMAL: {
          const int kk = (int)op.imm;
          R ev[kSpecMaxMv];
          for (int j = 0; j < kk; j++) {
            ST e;
            if (PREDRAWN) e = (ST)__ldcs(a.normals + (long long)n_norm * a.ld + i);
            else if (n_norm & 1) e = spare_n;
            else philox_normal_pair(a.seed, pid, a.step, (unsigned)(n_norm >> 1), T, e, spare_n);
            n_norm++;
            ev[j] = (R)e;
          }
          spec_mvnormal_sample<R>(regs + op.dst, regs + op.a, c_spec_params + op.b, ev, kk);
          break;
        }
        case OP_OBSERVE_MVNORMAL: w += spec_mvnormal_logpdf<R>(regs + op.dst, regs + op.a, c_spec_params + op.b, (int)op.imm); break;
        case OP_OBSERVE_DIST: w += (R)logpdf_scalar((int)op.imm, (double)regs[op.dst], (double)regs[op.a], (double)regs[op.b]); break;
        case OP_LOGPDF_DIST:
          regs[op.dst] = (R)logpdf_scalar((int)op.imm % 100, (double)regs[(int)op.imm / 100], (double)regs[op.a], (double)regs[op.b]);
          break;
        default: break;
      }
    }
    const R lw_old = (INIT || gather) ? R(0) : (R)logw[i];
    const ST lw_new = (ST)(INIT ? w : lw_old + w);             // particle_filter.jl:131,151 / :199,231
    logw[i] = lw_new;
    if (!INIT) st_stream(incr + i, (ST)w);
    acc = max_nan(acc, (double)lw_new);
  }
  publish_max(acc, a.partials, a.ticket, a.stats);
}

}  // namespace gb
