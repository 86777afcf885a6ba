// kernels_step.cuh -- K1/K2/K3: fused propose + weight + log-weight accumulation + online
// log-sum-exp / ESS partials, one launch per particle_filter_step! / initialize_particle_filter /
// importance_sampling call (particle_filter.jl:127-131, :149-151, :196-200, :224-232;
// importance.jl:28-31, :48-54; inference.jl:3-6; particle_filter.jl:3-12).
//
// HBM-bound streaming kernel: SoA columns, one 128-bit load/store per thread per column
// (V = 16/sizeof(R) consecutive particles per thread), grid = resident CTAs * 148 SMs, grid-stride
// over chunks.  The last CTA to finish folds the per-CTA (max, sum e, sum e^2) partials into the
// weight statistics, so there is no separate reduction launch.
#pragma once
#include "models.cuh"

namespace gb {

struct WeightStats {
  double m, s1, s2;   // shard-local online LSE triple
  double lse, ess;    // m + log s1 ; s1^2 / s2
};

template <typename R> struct VecOf;
template <> struct VecOf<double> { typedef double2 type; static constexpr int V = 2; };
template <> struct VecOf<float> { typedef float4 type; static constexpr int V = 4; };

template <typename R> GB_D void vload(const R *p, R (&out)[VecOf<R>::V]);
template <> GB_D void vload<double>(const double *p, double (&o)[2]) {
  double2 v = __ldcs(reinterpret_cast<const double2 *>(p));
  o[0] = v.x; o[1] = v.y;
}
template <> GB_D void vload<float>(const float *p, float (&o)[4]) {
  float4 v = __ldcs(reinterpret_cast<const float4 *>(p));
  o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
}
template <typename R> GB_D void vstore(R *p, const R (&in)[VecOf<R>::V]);
template <> GB_D void vstore<double>(double *p, const double (&i)[2]) {
  __stcs(reinterpret_cast<double2 *>(p), make_double2(i[0], i[1]));
}
template <> GB_D void vstore<float>(float *p, const float (&i)[4]) {
  __stcs(reinterpret_cast<float4 *>(p), make_float4(i[0], i[1], i[2], i[3]));
}
// plain (L2-allocating) store for data the very next kernel re-reads
template <typename R> GB_D void vstore_keep(R *p, const R (&in)[VecOf<R>::V]);
template <> GB_D void vstore_keep<double>(double *p, const double (&i)[2]) {
  *reinterpret_cast<double2 *>(p) = make_double2(i[0], i[1]);
}
template <> GB_D void vstore_keep<float>(float *p, const float (&i)[4]) {
  *reinterpret_cast<float4 *>(p) = make_float4(i[0], i[1], i[2], i[3]);
}

struct StepArgs {
  const void *state_in;   // S columns of stride ld (traces); unused when INIT
  void *state_out;        // S columns of stride ld (new_traces)
  void *logw;             // ld elements (in/out)
  void *incr;             // ld elements (log incremental weights), unused when INIT
  const int *anc;         // GATHER: local source index of every slot (deferred resample gather)
  const double *normals;  // PREDRAWN: [NN][ld] ; else NULL
  const double *uniforms; // PREDRAWN: [NU][ld]
  long long n, ld, pid_offset;
  unsigned long long seed;
  unsigned int step;
  Lse3 *partials;         // one per CTA
  unsigned int *ticket;
  WeightStats *stats;
  ModelParams mp;
};

GB_D Lse3 lse3_block_reduce(Lse3 v, Lse3 *smem /* kBlock/32 entries */) {
  v = lse3_warp_reduce(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) smem[warp] = v;
  __syncthreads();
  Lse3 r = lse3_empty();
  if (warp == 0) {
    if (lane < kBlock / 32) r = smem[lane];
    r = lse3_warp_reduce(r);
  }
  return r;  // valid in warp 0
}

GB_D Lse3 load_partial(const Lse3 *p) {   // L2 (coherent) read of another CTA's partial
  return Lse3{__ldcg(&p->m), __ldcg(&p->s1), __ldcg(&p->s2)};
}
GB_D void write_stats(WeightStats *st, const Lse3 &t) {
  st->m = t.m; st->s1 = t.s1; st->s2 = t.s2;
  // inference.jl:5: max == -Inf ? -Inf : max + log(sum(exp.(arr .- max)))
  st->lse = (t.m == -INFINITY) ? -INFINITY : t.m + log(t.s1);
  // particle_filter.jl:3-6 evaluates exp(-logsumexp(2 lnw)) == s1^2/s2 ; all -Inf => NaN there too
  st->ess = (t.s1 * t.s1) / t.s2;
}

// Fold this CTA's partial into global memory; the last CTA reduces all partials into `stats`.
GB_D void publish_partials(Lse3 acc, Lse3 *partials, unsigned int *ticket, WeightStats *stats) {
  __shared__ Lse3 s_red[kBlock / 32];
  __shared__ bool s_last;
  Lse3 r = lse3_block_reduce(acc, s_red);
  if (threadIdx.x == 0) {
    partials[blockIdx.x] = r;
    __threadfence();
    unsigned int t = atomicAdd(ticket, 1u);
    s_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  Lse3 a = lse3_empty();
  for (unsigned int i = threadIdx.x; i < gridDim.x; i += blockDim.x) {
    a = lse3_merge(a, load_partial(partials + i));
  }
  __syncthreads();
  Lse3 t = lse3_block_reduce(a, s_red);
  if (threadIdx.x == 0) {
    write_stats(stats, t);
    *ticket = 0;
  }
}

template <typename R, int KIND, int PK, bool INIT, bool PREDRAWN, bool GATHER>
__global__ void __launch_bounds__(kBlock) step_kernel(const StepArgs a) {
  typedef Transition<R, KIND, PK, INIT> Tr;
  constexpr int V = VecOf<R>::V, S = Tr::S, NN = Tr::NN, NU = Tr::NU;
  const R *__restrict__ sin = static_cast<const R *>(a.state_in);
  R *__restrict__ sout = static_cast<R *>(a.state_out);
  R *__restrict__ logw = static_cast<R *>(a.logw);
  R *__restrict__ incr = static_cast<R *>(a.incr);
  const Hoisted<R> hz = Tr::hoist(a.mp);
  Lse3 acc = lse3_empty();

  const long long nvec = (a.n + V - 1) / V;
  for (long long iv = (long long)blockIdx.x * kBlock + threadIdx.x; iv < nvec; iv += (long long)gridDim.x * kBlock) {
    const long long i0 = iv * V;
    R prev[S > 0 ? S : 1][V];
    R lw_old[V];
    if (!INIT) {
      if (GATHER) {
        // deferred maybe_resample! gather (particle_filter.jl:264-267): slot k continues the trace of
        // ancestor anc[k]; its log weight was reset to exactly 0.0
        int src[V];
        if constexpr (V == 2) { int2 t = __ldcs(reinterpret_cast<const int2 *>(a.anc + i0)); src[0] = t.x; src[1] = t.y; }
        else { int4 t = __ldcs(reinterpret_cast<const int4 *>(a.anc + i0)); src[0] = t.x; src[1] = t.y; src[2] = t.z; src[3] = t.w; }
#pragma unroll
        for (int s = 0; s < S; s++)
#pragma unroll
          for (int v = 0; v < V; v++) prev[s][v] = __ldg(sin + (long long)s * a.ld + src[v]);
#pragma unroll
        for (int v = 0; v < V; v++) lw_old[v] = R(0);
      } else {
#pragma unroll
        for (int s = 0; s < S; s++) vload<R>(sin + (long long)s * a.ld + i0, prev[s]);
        vload<R>(logw + i0, lw_old);
      }
    }
    R next[S][V], w[V];
#pragma unroll
    for (int v = 0; v < V; v++) {
      const long long i = i0 + v;
      R eps[NN > 0 ? NN : 1], us[NU > 0 ? NU : 1];
      if (PREDRAWN) {
#pragma unroll
        for (int d = 0; d < NN; d++) eps[d] = (R)__ldcs(a.normals + (long long)d * a.ld + i);
#pragma unroll
        for (int d = 0; d < NU; d++) us[d] = (R)__ldcs(a.uniforms + (long long)d * a.ld + i);
      } else {
        const unsigned long long pid = (unsigned long long)(a.pid_offset + i);
#pragma unroll
        for (int d = 0; d < NN; d += 2) {
          R e0, e1;
          philox_normal_pair(a.seed, pid, a.step, (unsigned)(d / 2), e0, e1);
          eps[d] = e0;
          if (d + 1 < NN) eps[d + 1] = e1;
        }
#pragma unroll
        for (int d = 0; d < NU; d += 2) {
          R u0, u1;
          philox_uniform_pair(a.seed, pid, a.step, (unsigned)(d / 2), u0, u1);
          us[d] = u0;
          if (d + 1 < NU) us[d + 1] = u1;
        }
      }
      R pv[S], nx[S];
#pragma unroll
      for (int s = 0; s < S; s++) pv[s] = INIT ? R(0) : prev[s][v];
      w[v] = Tr::apply(a.mp, hz, pv, nx, eps, us);
#pragma unroll
      for (int s = 0; s < S; s++) next[s][v] = nx[s];
    }
    R lw_new[V];
#pragma unroll
    for (int v = 0; v < V; v++) {
      lw_new[v] = INIT ? w[v] : lw_old[v] + w[v];     // particle_filter.jl:131,151 / :199,231
      if (i0 + v < a.n) lse3_push(acc, (double)lw_new[v]);
    }
#pragma unroll
    for (int s = 0; s < S; s++) vstore<R>(sout + (long long)s * a.ld + i0, next[s]);
    vstore_keep<R>(logw + i0, lw_new);
    if (!INIT) vstore<R>(incr + i0, w);
  }
  publish_partials(acc, a.partials, a.ticket, a.stats);
}

// Reduce an existing log-weight array (after gb_pf_set_log_weights, clone, raw-array entries).
template <typename R>
__global__ void __launch_bounds__(kBlock) lse_kernel(const R *__restrict__ logw, long long n, Lse3 *partials,
                                                     unsigned int *ticket, WeightStats *stats) {
  Lse3 acc = lse3_empty();
  for (long long i = (long long)blockIdx.x * kBlock + threadIdx.x; i < n; i += (long long)gridDim.x * kBlock)
    lse3_push(acc, (double)__ldg(logw + i));
  publish_partials(acc, partials, ticket, stats);
}

// ------------------------------------------------------------------------------------------
// cfg 4: Bayesian linear regression importance sampler (importance.jl:20-36 over
// w ~ Map(normal)(0,1), y_j ~ normal(x_j . w, sigma); map/generate.jl:10-37).  One thread per
// sample keeps w[64] in registers; X sits in constant memory so every multiply-add takes its matrix
// operand from the constant bank (no load instruction, no tensor cores: north_star).
// ------------------------------------------------------------------------------------------
constexpr int kBlrMax = 64;
__constant__ double c_blr_X[kBlrMax * kBlrMax];   // [n][64] zero padded
__constant__ double c_blr_y[kBlrMax];

struct BlrArgs {
  void *state_out;   // D columns of stride ld
  void *logw;
  const double *normals;  // PREDRAWN [D][ld]
  long long n, ld, pid_offset;
  unsigned long long seed;
  int D, nobs;
  double sigma;
  Lse3 *partials;
  unsigned int *ticket;
  WeightStats *stats;
};

template <typename R, bool PREDRAWN>
__global__ void __launch_bounds__(128) blr_generate_kernel(const BlrArgs a) {
  R *__restrict__ sout = static_cast<R *>(a.state_out);
  R *__restrict__ logw = static_cast<R *>(a.logw);
  const R sigma = (R)a.sigma, log_sigma = Math<R>::log_((R)a.sigma);
  Lse3 acc = lse3_empty();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += (long long)gridDim.x * blockDim.x) {
    R w[kBlrMax];
    if (PREDRAWN) {
#pragma unroll
      for (int d = 0; d < kBlrMax; d++) w[d] = d < a.D ? (R)__ldcs(a.normals + (long long)d * a.ld + i) : R(0);
    } else {
      const unsigned long long pid = (unsigned long long)(a.pid_offset + i);
#pragma unroll
      for (int d = 0; d < kBlrMax; d += 2) {
        R e0 = R(0), e1 = R(0);
        if (d < a.D) philox_normal_pair(a.seed, pid, 0u, (unsigned)(d / 2), e0, e1);
        w[d] = e0;
        w[d + 1] = (d + 1 < a.D) ? e1 : R(0);
      }
    }
    // normal(0, 1): mu + std * eps = 0 + 1 * eps (normal.jl:128) is exactly eps
#pragma unroll
    for (int d = 0; d < kBlrMax; d++)
      if (d < a.D) st_stream(sout + (long long)d * a.ld + i, w[d]);
    R lw = R(0);
#pragma unroll 1
    for (int j = 0; j < a.nobs; j++) {
      R m0 = R(0), m1 = R(0), m2 = R(0), m3 = R(0);
#pragma unroll
      for (int d = 0; d < kBlrMax; d += 4) {
        m0 = fma((R)c_blr_X[j * kBlrMax + d + 0], w[d + 0], m0);
        m1 = fma((R)c_blr_X[j * kBlrMax + d + 1], w[d + 1], m1);
        m2 = fma((R)c_blr_X[j * kBlrMax + d + 2], w[d + 2], m2);
        m3 = fma((R)c_blr_X[j * kBlrMax + d + 3], w[d + 3], m3);
      }
      lw += logpdf_normal_ls<R>((R)c_blr_y[j], (m0 + m1) + (m2 + m3), sigma, log_sigma);
    }
    logw[i] = lw;
    lse3_push(acc, (double)lw);
  }
  // publish_partials assumes kBlock threads per CTA for its shared scratch; 128 <= kBlock is fine
  __shared__ Lse3 s_red[kBlock / 32];
  __shared__ bool s_last;
  {
    Lse3 v = lse3_warp_reduce(acc);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) s_red[warp] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
      Lse3 r = lse3_empty();
      for (int k = 0; k < (int)(blockDim.x >> 5); k++) r = lse3_merge(r, s_red[k]);
      a.partials[blockIdx.x] = r;
      __threadfence();
      unsigned int t = atomicAdd(a.ticket, 1u);
      s_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    Lse3 r = lse3_empty();
    for (unsigned int k = threadIdx.x; k < gridDim.x; k += blockDim.x) r = lse3_merge(r, load_partial(a.partials + k));
    r = lse3_warp_reduce(r);
    __syncthreads();
    if (lane == 0) s_red[warp] = r;
    __syncthreads();
    if (threadIdx.x == 0) {
      Lse3 t = lse3_empty();
      for (int k = 0; k < (int)(blockDim.x >> 5); k++) t = lse3_merge(t, s_red[k]);
      write_stats(a.stats, t);
      *a.ticket = 0;
    }
  }
}

// importance.jl:34,57: log_normalized_weights = log_weights .- log_total_weight
template <typename R>
__global__ void __launch_bounds__(kBlock) normalize_kernel(const R *__restrict__ logw, long long n, const WeightStats *st,
                                                           double *__restrict__ out) {
  const double lse = st->lse;
  for (long long i = (long long)blockIdx.x * kBlock + threadIdx.x; i < n; i += (long long)gridDim.x * kBlock)
    out[i] = (double)logw[i] - lse;
}

}  // namespace gb
