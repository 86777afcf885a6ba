// kernels_small.cuh -- the whole filter loop in ONE kernel for small particle counts (N <= 4096).
//
// A multi-kernel step cannot go below ~4 dependent launches (~22 us per step with programmatic dependent launch,
// profiles/r1g_small_n.md), which is what BASELINE config 1 (N = 1000, T = 100) pays per step no matter how little
// work there is.  Here one CTA runs   T x { maybe_resample!(systematic); particle_filter_step! }   (the loop of
// docs/src/tutorials/smc.md:674-689) with __syncthreads() between the phases: log weights live in registers, ancestors
// in shared memory, traces in the (L2-resident) double buffer.  Same arithmetic as the large-N kernels -- the same
// Transition<>::apply, Philox counters, exp_and_quantize and exact slots_below -- so states, ancestors and log weights
// are bit-identical to the multi-kernel path; only the order of the (sum e, sum e^2) reduction differs, i.e. lse / ESS /
// log_ml_est agree to rounding.
#pragma once
#include "kernels_step.cuh"

namespace gb {

constexpr int kSmallThreads = 512, kSmallItems = 8, kSmallMax = kSmallThreads * kSmallItems, kSmallHeavyCap = 32;

struct SmallRunArgs {
  ModelParams mp;
  const double *obs;       // DEVICE [T][nobs]
  int T, nobs;
  void *state[2];          // trace double buffer, S columns of stride ld
  int cur;                 // buffer holding the current traces at entry
  void *logw, *incr;
  int *anc;                // global copy of the last resample's ancestors (Gen's `parents`)
  long long n, ld;
  unsigned long long seed;
  unsigned int step;       // index of the step that produced the current weights
  int bits;
  RunCtl *ctl;             // in: log_ml_est, ess_threshold, log_n; out: n_resampled, log_ml_est
  WeightStats *stats;      // out: m = max of the final log weights
};

GB_D double small_block_sum(double v, double *s /* 16 */) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = v;
  __syncthreads();
  double r = 0.0;
  for (int k = 0; k < kSmallThreads / 32; k++) r += s[k];    // fixed order: every thread gets the same bits
  return r;
}
GB_D double small_block_max(double v, double *s /* 16 */) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = max_nan(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = v;
  __syncthreads();
  double r = -INFINITY;
  for (int k = 0; k < kSmallThreads / 32; k++) r = max_nan(r, s[k]);
  return r;
}
// exclusive prefix of one u64 per thread + block total
GB_D uint64_t small_block_scan(uint64_t v, uint64_t *s /* 17 */, uint64_t *total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint64_t inc = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint64_t o = shfl_up_u64(inc, d);
    if (lane >= d) inc += o;
  }
  __syncthreads();
  if (lane == 31) s[warp] = inc;
  __syncthreads();
  uint64_t base = 0, tot = 0;
  for (int k = 0; k < kSmallThreads / 32; k++) {
    if (k < warp) base += s[k];
    tot += s[k];
  }
  *total = tot;
  return base + inc - v;
}

template <typename R, int KIND>
__global__ void __launch_bounds__(kSmallThreads, 1) small_run_kernel(const SmallRunArgs a) {
  typedef typename Arith<R>::type A;     // storage R, arithmetic fp64 (models.cuh)
  typedef Transition<A, KIND, PROPOSAL_DEFAULT, false> Tr;
  constexpr int S = Tr::S, NN = Tr::NN, NU = Tr::NU, P = kSmallItems;
  __shared__ double s_tab[kMathTabDoubles];
  __shared__ int s_anc[kSmallMax];
  __shared__ int s_kend[kSmallThreads];
  __shared__ double s_d[kSmallThreads / 32];
  __shared__ uint64_t s_u[kSmallThreads / 32 + 1];
  __shared__ int3 s_heavy[kSmallMax / kSmallHeavyCap + 1];
  __shared__ int s_nheavy;
  __shared__ double s_wexp[64];
  const MathTables T = stage_math_tables(s_tab);
  const double *wt = stage_wexp_table(s_wexp);
  __syncthreads();
  const int n = (int)a.n;
  const int per = (n + kSmallThreads - 1) / kSmallThreads;       // particles per thread (blocked layout): <= P
  const int i0 = threadIdx.x * per;
  R *__restrict__ logw = static_cast<R *>(a.logw);
  R *__restrict__ incr = static_cast<R *>(a.incr);
  ModelParams mp = a.mp;
  const Hoisted<A> hz = Tr::hoist(mp);
  R lw[P];
#pragma unroll
  for (int j = 0; j < P; j++) lw[j] = (j < per && i0 + j < n) ? logw[i0 + j] : R(0);
  double log_ml_est = a.ctl->log_ml_est;
  const double thr = a.ctl->ess_threshold, log_n = a.ctl->log_n;
  int n_resampled = 0, cur = a.cur;
  double m = -INFINITY;
  for (int t = 0; t < a.T; t++) {
    const unsigned int wstep = a.step + (unsigned)t;               // index of the step that produced the weights
    // ---- maybe_resample! (particle_filter.jl:249-275): statistics, decision, systematic ancestors ----
    double mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < P; j++)
      if (j < per && i0 + j < n) mx = max_nan(mx, (double)lw[j]);
    m = small_block_max(mx, s_d);
    double s1 = 0.0, s2 = 0.0;
    uint64_t c[P], qt = 0;
#pragma unroll
    for (int j = 0; j < P; j++) {
      double e = 0.0;
      uint64_t q = 0;
      if (j < per && i0 + j < n) exp_and_quantize((double)lw[j] - m, a.bits, e, q, wt);
      s1 += e;
      s2 = fma(e, e, s2);
      qt += q;
      c[j] = qt;
    }
    const double t1 = small_block_sum(s1, s_d), t2 = small_block_sum(s2, s_d);
    uint64_t q_total;
    const uint64_t ex = small_block_scan(qt, s_u, &q_total);
    const double lse = (m == -INFINITY) ? -INFINITY : m + log(t1);  // inference.jl:5
    const double ess = (t1 * t1) / t2;                              // particle_filter.jl:3-6
    const bool resample = ess < thr;                                // :256 (strict; NaN => false), uniform over the CTA
    if (resample) {
      log_ml_est += lse - log_n;                                    // :263
      n_resampled++;
      SweepParams sp;
      sp.q_total = q_total;
      sp.q_inv = 1.0 / (double)q_total;
      sp.dscale = (uint64_t)n << 32;
      sp.ratio = (double)n / (double)q_total;
      sp.u0 = philox_block(a.seed, 0, wstep, PK_RESAMPLE, 0).x;
      int kend[P];
#pragma unroll
      for (int j = 0; j < P; j++) kend[j] = (int)slots_below(ex + c[j], sp);
      if (threadIdx.x == 0) s_nheavy = 0;
      s_kend[threadIdx.x] = kend[P - 1];
      __syncthreads();
      int kstart = threadIdx.x == 0 ? 0 : s_kend[threadIdx.x - 1];
#pragma unroll
      for (int j = 0; j < P; j++) {
        if (kend[j] - kstart > kSmallHeavyCap) s_heavy[atomicAdd(&s_nheavy, 1)] = make_int3(i0 + j, kstart, kend[j]);
        else
          for (int k = kstart; k < kend[j]; k++) s_anc[k] = i0 + j;
        kstart = kend[j];
      }
      __syncthreads();
      const int nh = s_nheavy;
      for (int e = 0; e < nh; e++) {
        const int3 h = s_heavy[e];
        for (int k = h.y + threadIdx.x; k < h.z; k += kSmallThreads) s_anc[k] = h.x;
      }
      __syncthreads();
      for (int k = threadIdx.x; k < n; k += kSmallThreads) a.anc[k] = s_anc[k];
    }
    // ---- particle_filter_step! (:217-240), reading through the ancestors when a resample was decided (:264-267) ----
    const R *__restrict__ sin = static_cast<const R *>(a.state[cur]);
    R *__restrict__ sout = static_cast<R *>(a.state[cur ^ 1]);
    mp.obs0 = a.obs[(long long)t * a.nobs];
    const unsigned int nstep = wstep + 1u;
#pragma unroll
    for (int j = 0; j < P; j++) {
      const int i = i0 + j;
      if (j < per && i < n) {
        const int src = resample ? s_anc[i] : i;
        A pv[S], nx[S], eps[NN > 0 ? NN : 1], us[NU > 0 ? NU : 1];
#pragma unroll
        for (int s = 0; s < S; s++) pv[s] = (A)sin[(long long)s * a.ld + src];
#pragma unroll
        for (int d = 0; d < NN; d += 2) {
          R e0, e1;
          philox_normal_pair(a.seed, (unsigned long long)i, nstep, (unsigned)(d / 2), T, e0, e1);
          eps[d] = (A)e0;
          if (d + 1 < NN) eps[d + 1] = (A)e1;
        }
#pragma unroll
        for (int d = 0; d < NU; d += 2) {
          R u0, u1;
          philox_uniform_pair(a.seed, (unsigned long long)i, nstep, (unsigned)(d / 2), u0, u1);
          us[d] = (A)u0;
          if (d + 1 < NU) us[d + 1] = (A)u1;
        }
        const A w = Tr::apply(mp, hz, T, pv, nx, eps, us);
#pragma unroll
        for (int s = 0; s < S; s++) sout[(long long)s * a.ld + i] = (R)nx[s];
        lw[j] = (R)((resample ? A(0) : (A)lw[j]) + w);              // :231 (after :266 when resampled)
        if (t == a.T - 1) incr[i] = (R)w;
      }
    }
    cur ^= 1;
    __syncthreads();     // the next iteration's gather reads traces other threads just wrote
  }
  double mx = -INFINITY;
#pragma unroll
  for (int j = 0; j < P; j++)
    if (j < per && i0 + j < n) {
      logw[i0 + j] = lw[j];
      mx = max_nan(mx, (double)lw[j]);
    }
  m = small_block_max(mx, s_d);
  if (threadIdx.x == 0) {
    a.stats->m = m;
    a.ctl->log_ml_est = log_ml_est;
    a.ctl->n_resampled = n_resampled;
    a.ctl->do_resample = 0;
  }
}

}  // namespace gb
