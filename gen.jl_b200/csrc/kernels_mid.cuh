// kernels_mid.cuh -- the whole filter loop in ONE persistent kernel for mid-size particle counts (4096 < N <~ 10^7;
// BASELINE config 2 is the stochastic-volatility model at N = 10^6, T = 1000).
//
// At N = 10^6 a step of the multi-kernel run (statistics -> ancestors -> overflow -> step, chained with programmatic
// dependent launch) costs ~39 us, half of which is the four-launch floor (profiles/r1g_small_n.md): the working set
// (~40 MB) lives in the 126 MB L2 and every kernel is over in a few microseconds.  Here ONE co-resident grid (one launch
// for all T steps, cudaLaunchCooperativeKernel) runs  T x { maybe_resample!(systematic); particle_filter_step! }  (the
// loop of docs/src/tutorials/smc.md:674-689) with three grid barriers per step:
//
//   phase 1  statistics   every CTA, over the tiles it owns: e = exp(lw - m), q = quantised weight (kept in qbuf, which
//                         only this CTA reads back), CTA partial (sum e, sum e^2, sum q)          (inference.jl:3-6,
//                         particle_filter.jl:3-12)
//            -- barrier --
//   phase 2  ancestors    every CTA folds the G partials in the same fixed order (identical bits everywhere): lse, ESS,
//                         the maybe_resample! decision (:256) and the log-ML update (:263) need no designated CTA; the
//                         integer CDF mass before its first tile is the sum of the partials of the CTAs before it; then
//                         the large-N tile fill (tile_fill, kernels_resample.cuh) writes the ancestors of its tiles' slots
//            -- barrier --
//   phase 3  step         gather through the ancestors (:264-267) + propose + weight (:217-240) over an even share of
//                         the particles, CTA partial max
//            -- barrier --                                         (every CTA folds the G maxima: the next step's m)
//
// Data written by one CTA and read by another inside this kernel (log weights, ancestors, traces, partials) is read with
// ld.global.cg (L2, the point of coherence) -- the ld.global.nc / default-cached loads of the streaming kernels may be
// served from an SM's own L1, which is only invalidated at kernel boundaries.  Same arithmetic as the other paths (the same
// Transition<>::apply, Philox counters, exp_and_quantize, slots_estimate / slots_below): states, ancestors and log
// weights are bit-identical to the API loop; only the order of the (sum e, sum e^2) reduction differs, i.e. lse / ESS /
// log_ml_est agree to rounding (tests/test_gpu_parity.py::test_persistent_run_equals_api_loop).
#pragma once
#include "kernels_resample.cuh"

namespace gb {

struct MidPartial { double s1, s2; unsigned long long q; double mx; };   // one per CTA (32 bytes)

struct MidBarrier {
  unsigned long long *release;   // 1 word (its own 128-byte line)
  unsigned long long *arrive;    // [gridDim.x]
};
struct MidRunArgs {
  ModelParams mp;
  const double *obs;       // DEVICE [T][nobs]
  int T, nobs;
  void *state[2];          // trace double buffer, S columns of stride ld
  int cur;                 // buffer holding the current traces at entry
  void *logw, *incr;
  int *anc;                // the last resample's ancestors (Gen's `parents`)
  unsigned long long *qbuf;
  MidPartial *part;        // [gridDim.x]
  MidBarrier barrier;      // all words zero at launch
  long long n, ld, ntiles;
  unsigned long long seed;
  unsigned int step;       // index of the step that produced the current weights
  int bits;
  RunCtl *ctl;             // in: log_ml_est, ess_threshold, log_n; out: n_resampled, log_ml_est
  WeightStats *stats;      // in: m = max of the current log weights; out: max of the final ones
  unsigned long long *dbg; // GENB200_MID_DEBUG: [T][8] globaltimer stamps of the last CTA at the phase boundaries (nullptr: off)
};

GB_D unsigned long long ld_acquire_gpu(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
GB_D void st_release_gpu(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// Grid barrier, all threads of all CTAs (the grid is co-resident: cooperative launch).  One monotone arrival counter:
// thread 0 of every CTA adds one (after a release fence) and polls it with RELAXED loads until the epoch's count is
// reached, then one acquire fence.  Measured at G = 592 (GENB200_MID_DEBUG=1, profiles/r2x_persistent_run.md): polling
// with ld.acquire (a fence per probe) 3.5 us per barrier; per-CTA arrival words + a master CTA + a release word 4.0 us
// (two more fence + store + poll hops in the chain); this one the fastest of the three.
GB_D void mid_grid_barrier(const MidBarrier &bar, unsigned long long &epoch) {
  epoch += gridDim.x;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(bar.release, 1ull);
    while (*((volatile unsigned long long *)bar.release) < epoch) {}
    __threadfence();
  }
  __syncthreads();
}

GB_D unsigned long long mid_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
#define MID_STAMP(k) if (a.dbg && b == (int)gridDim.x - 1 && threadIdx.x == 0) a.dbg[(long long)t * 8 + (k)] = mid_now()

// MINB = CTAs per SM the kernel is compiled for: 4 (64 registers) for the 1-D models -- N = 10^6 is 489 tiles, and 4 x 148
// co-resident CTAs give every tile its own CTA, so phases 1 and 2 are ONE round of tiles instead of two -- 3 for bearings
template <typename R, int KIND, int MINB>
__global__ void __launch_bounds__(kBlock, MINB) mid_run_kernel(const MidRunArgs a) {
  typedef typename Arith<R>::type A;     // storage R, arithmetic fp64 (models.cuh)
  typedef Transition<A, KIND, PROPOSAL_DEFAULT, false> Tr;
  constexpr int S = Tr::S, NN = Tr::NN, NU = Tr::NU;
  __shared__ double s_tab[kMathTabDoubles];
  __shared__ double s_wexp[64];
  __shared__ TileSmem sm;
  __shared__ double s_d[3][kBlock / 32];
  __shared__ uint64_t s_u[2][kBlock / 32];
  const MathTables T = stage_math_tables(s_tab);
  const double *wt = stage_wexp_table(s_wexp);
  tile_smem_clear(sm);                   // (ends with a barrier)
  const int G = gridDim.x, b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long n = a.n;
  // tiles [tb0, tb1) and particles [p0, p1) of this CTA: contiguous, balanced shares
  const long long tb0 = a.ntiles * b / G, tb1 = a.ntiles * (b + 1) / G;
  const long long p0 = n * b / G, p1 = n * (b + 1) / G;
  R *__restrict__ logw = static_cast<R *>(a.logw);
  R *__restrict__ incr = static_cast<R *>(a.incr);
  ModelParams mp = a.mp;
  const Hoisted<A> hz = Tr::hoist(mp);
  double log_ml_est = a.ctl->log_ml_est;
  const double thr = a.ctl->ess_threshold, log_n = a.ctl->log_n;
  int n_resampled = 0, cur = a.cur;
  double m = __ldcg(&a.stats->m);
  unsigned long long bar_target = 0;      // epoch of the last barrier
  const AncOut ao{a.anc, 0, n, nullptr, 0, 0};
  for (int t = 0; t < a.T; t++) {
    const unsigned int wstep = a.step + (unsigned)t;               // index of the step that produced the weights
    // ---- phase 1: statistics of the tiles this CTA owns ----
    MID_STAMP(0);
    double s1 = 0.0, s2 = 0.0;
    uint64_t qt = 0;
    for (long long tile = tb0; tile < tb1; tile++) {
#pragma unroll
      for (int j = 0; j < kItems; j++) {
        const long long i = tile * kTile + (long long)j * kBlock + threadIdx.x;
        double e = 0.0;
        uint64_t q = 0;
        if (i < n) exp_and_quantize((double)__ldcg(logw + i) - m, a.bits, e, q, wt);
        s1 += e;
        s2 = fma(e, e, s2);
        qt += q;
        a.qbuf[i] = q;                                             // (arrays are padded to whole tiles)
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
      s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    qt = warp_sum_u64(qt);
    if (lane == 0) { s_d[0][warp] = s1; s_d[1][warp] = s2; s_u[0][warp] = qt; }
    __syncthreads();
    if (threadIdx.x == 0) {
      double x = 0.0, y = 0.0;
      uint64_t z = 0;
      for (int k = 0; k < kBlock / 32; k++) { x += s_d[0][k]; y += s_d[1][k]; z += s_u[0][k]; }
      a.part[b].s1 = x; a.part[b].s2 = y; a.part[b].q = z;
    }
    MID_STAMP(1);
    mid_grid_barrier(a.barrier, bar_target);
    MID_STAMP(2);
    // ---- phase 2: fold the partials (same order in every CTA), decide, ancestors of this CTA's tiles ----
    double t1 = 0.0, t2 = 0.0;
    uint64_t qall = 0, qbefore = 0;
    for (int c = threadIdx.x; c < G; c += kBlock) {
      t1 += __ldcg(&a.part[c].s1);
      t2 += __ldcg(&a.part[c].s2);
      const uint64_t qc = __ldcg(&a.part[c].q);
      qall += qc;
      if (c < b) qbefore += qc;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      t1 += __shfl_xor_sync(0xffffffffu, t1, o);
      t2 += __shfl_xor_sync(0xffffffffu, t2, o);
    }
    qall = warp_sum_u64(qall);
    qbefore = warp_sum_u64(qbefore);
    if (lane == 0) { s_d[0][warp] = t1; s_d[1][warp] = t2; s_u[0][warp] = qall; s_u[1][warp] = qbefore; }
    __syncthreads();
    t1 = t2 = 0.0; qall = qbefore = 0;
    for (int k = 0; k < kBlock / 32; k++) { t1 += s_d[0][k]; t2 += s_d[1][k]; qall += s_u[0][k]; qbefore += s_u[1][k]; }
    __syncthreads();                                                // (s_d / s_u are reused below)
    const double lse = (m == -INFINITY) ? -INFINITY : m + log(t1);  // inference.jl:5
    const double ess = (t1 * t1) / t2;                              // particle_filter.jl:3-6
    const bool resample = ess < thr && qall > 0;                    // :256 (strict; NaN => false), the same in every CTA
    if (resample) {
      log_ml_est += lse - log_n;                                    // :263
      n_resampled++;
      SweepParams sp;
      sp.q_total = qall;
      sp.q_inv = 1.0 / (double)qall;
      sp.dscale = (uint64_t)n << 32;
      sp.ratio = (double)n / (double)qall;
      sp.u0 = philox_block(a.seed, 0, wstep, PK_RESAMPLE, 0).x;
      uint64_t cb = qbefore;
      for (long long tile = tb0; tile < tb1; tile++) {
        long long slot0, slot_end;
        uint64_t tot;
        __syncthreads();
        tile_fill<true>(a.qbuf, sp, tile, nullptr, nullptr, 0, cb, 0, 0, 0, ao, 0, sm, &slot0, &slot_end, &tot);
        // a tile that owns more than kChunk slots: the remaining windows (degenerate weights; rare)
        for (long long lo = slot0 + kChunk; lo < slot_end; lo += kChunk) {
          const long long hi = lo + kChunk < slot_end ? lo + kChunk : slot_end;
          __syncthreads();
          tile_fill<false>(a.qbuf, sp, tile, nullptr, nullptr, 0, cb, slot0, lo, hi, ao, 0, sm, nullptr, nullptr, nullptr);
        }
        cb += tot;
      }
    }
    MID_STAMP(3);
    mid_grid_barrier(a.barrier, bar_target);
    MID_STAMP(4);
    // ---- phase 3: particle_filter_step! over this CTA's share, reading through the ancestors after a resample ----
    const R *__restrict__ sin = static_cast<const R *>(a.state[cur]);
    R *__restrict__ sout = static_cast<R *>(a.state[cur ^ 1]);
    mp.obs0 = a.obs[(long long)t * a.nobs];
    const unsigned int nstep = wstep + 1u;
    const bool last = t == a.T - 1;
    double acc = -INFINITY;
    // (software pipeline: every phase is latency-bound at a few particles per thread -- the ancestor index is loaded two
    // iterations ahead and the row + log weight one iteration ahead, so the two dependent L2 round trips of iteration k + 1
    // overlap the arithmetic of iteration k)
    long long i = p0 + threadIdx.x;
    long long src_n2 = 0;                                            // ancestor of iteration k + 2
    R pv_n[S > 0 ? S : 1];                                           // row of iteration k + 1
    R lw_n = R(0);
    auto load_src = [&](long long ix) { return resample ? (long long)__ldcg(a.anc + ix) : ix; };
    auto load_row = [&](long long ix, long long srcx) {
#pragma unroll
      for (int s = 0; s < S; s++) pv_n[s] = __ldcg(sin + (long long)s * a.ld + srcx);
      lw_n = resample ? R(0) : __ldcg(logw + ix);
    };
    if (i < p1) {
      const long long s0 = load_src(i);
      const long long s1x = i + kBlock < p1 ? load_src(i + kBlock) : 0;
      load_row(i, s0);
      src_n2 = s1x;
    }
    for (; i < p1; i += kBlock) {
      A pv[S], nx[S], eps[NN > 0 ? NN : 1], us[NU > 0 ? NU : 1];
#pragma unroll
      for (int s = 0; s < S; s++) pv[s] = (A)pv_n[s];
      const R lw_old = lw_n;
      if (i + kBlock < p1) {
        const long long srcx = src_n2;
        if (i + 2 * kBlock < p1) src_n2 = load_src(i + 2 * kBlock);
        load_row(i + kBlock, srcx);
      }
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
      const R lw_new = (R)((A)lw_old + w);                          // :231 (after :266 when resampled)
      logw[i] = lw_new;
      if (last) incr[i] = (R)w;
      acc = max_nan(acc, (double)lw_new);
    }
    cur ^= 1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc = max_nan(acc, __shfl_xor_sync(0xffffffffu, acc, o));
    if (lane == 0) s_d[2][warp] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      double r = -INFINITY;
      for (int k = 0; k < kBlock / 32; k++) r = max_nan(r, s_d[2][k]);
      a.part[b].mx = r;
    }
    MID_STAMP(5);
    mid_grid_barrier(a.barrier, bar_target);
    MID_STAMP(6);
    if (a.dbg) {                 // (debug only) cost of a barrier every CTA reaches at the same time
      mid_grid_barrier(a.barrier, bar_target);
      MID_STAMP(7);
    }
    double mm = -INFINITY;
    for (int c = threadIdx.x; c < G; c += kBlock) mm = max_nan(mm, __ldcg(&a.part[c].mx));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mm = max_nan(mm, __shfl_xor_sync(0xffffffffu, mm, o));
    if (lane == 0) s_d[2][warp] = mm;
    __syncthreads();
    mm = -INFINITY;
    for (int k = 0; k < kBlock / 32; k++) mm = max_nan(mm, s_d[2][k]);
    m = mm;
    __syncthreads();
  }
  if (b == 0 && threadIdx.x == 0) {
    a.stats->m = m;
    a.ctl->log_ml_est = log_ml_est;
    a.ctl->n_resampled = n_resampled;
    a.ctl->do_resample = 0;
  }
}

}  // namespace gb
