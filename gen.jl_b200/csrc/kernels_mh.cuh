// kernels_mh.cuh -- Metropolis-Hastings rejuvenation moves on every particle of the filter.
//
// The reference's users write   for i in 1:N  state.traces[i], _ = mh(state.traces[i], ...)   between
// particle_filter_step! and maybe_resample! (docs/src/tutorials/smc.md:459-463, :518-520).  The engine's store is
// Markov (current state + the buffer the last step read from), so the moves offered are the ones that touch the
// latent choices of the LATEST kernel application, whose parent state is still resident:
//   MH_RESIMULATE  mh(trace, select(<latest latents>))  (src/inference/mh.jl:16-27): the selected choices are
//                  re-drawn from the model's transition; regenerate's weight is the score change of the retained
//                  choices whose arguments changed (src/dynamic/regenerate.jl:36-50) = log p(y|x') - log p(y|x).
//   MH_DRIFT       mh(trace, proposal, args) with a Gaussian-drift proposal c' ~ normal(c, std) on every latest latent
//                  choice (mh.jl:37-58, tutorial perturbation_proposal smc.md:492-502): alpha = update weight
//                  (dynamic/update.jl:46-62: score' - score of the constrained latents and of the observation)
//                  - fwd_weight + bwd_weight.
// accept iff log(rand()) < alpha (mh.jl:20, :52).  Particle weights are untouched, as in the reference.
#pragma once
#include "kernels_step.cuh"

namespace gb {

enum { MH_RESIMULATE = 0, MH_DRIFT = 1 };
constexpr uint32_t kMhPairBase = 0x800000u;   // Philox draw-index block reserved for the moves: base + 16 * move + k

struct MhArgs {
  ModelParams mp;
  const void *parent;   // the buffer the last step read its inputs from
  void *cur;            // current traces (updated in place)
  const int *anc;       // parent row of particle i (the last step's fused gather); nullptr: row i
  const RunCtl *ctl;    // fused run: the gather decision was taken on the device (ctl->do_resample); nullptr: host-known
  const double *normals, *uniforms;   // PREDRAWN: [rows][ld]; the LAST uniform row is the accept draw
  double *alpha_out;    // optional: log acceptance ratios
  void *lik;            // mh_resim_lik_kernel: score of the constrained choices of every particle's latest step (in/out)
  long long n, ld, pid_offset;
  unsigned long long seed;
  unsigned int step, move;
  double drift_sd;
  unsigned long long *accepted;
};

template <typename R, int KIND, bool INIT, int MOVE, bool PREDRAWN>
__global__ void __launch_bounds__(kBlock) mh_kernel(const MhArgs a) {
  typedef typename Arith<R>::type A;     // storage R, arithmetic fp64 (models.cuh)
  using Tr = Transition<A, KIND, PROPOSAL_DEFAULT, INIT>;
  constexpr int S = Tr::S, NN = MOVE == MH_DRIFT ? Tr::ND : Tr::NN, NU = MOVE == MH_DRIFT ? 0 : Tr::NU;
  static_assert(NU <= 1, "mh_kernel draws at most one transition uniform (second word of the accept pair)");
  __shared__ double s_tab[kMathTabDoubles];
  const MathTables T = stage_math_tables(s_tab);
  __syncthreads();
  const Hoisted<A> h = Tr::hoist(a.mp);
  const R *__restrict__ par = static_cast<const R *>(a.parent);
  R *__restrict__ cur = static_cast<R *>(a.cur);
  const A dsd = (A)a.drift_sd, ldsd = Math<A>::log_((A)a.drift_sd);
  const uint32_t base = kMhPairBase + 16u * a.move;
  const bool via_anc = a.ctl ? (a.ctl->do_resample != 0) : (a.anc != nullptr);
  unsigned int nacc = 0;
  for (long long i = (long long)blockIdx.x * kBlock + threadIdx.x; i < a.n; i += (long long)gridDim.x * kBlock) {
    const unsigned long long pid = (unsigned long long)(a.pid_offset + i);
    A prev[S], st[S], next[S], eps[NN > 0 ? NN : 1], us[NU > 0 ? NU : 1];
    if (!INIT) {
      const long long src = via_anc ? (long long)a.anc[i] : i;
#pragma unroll
      for (int s = 0; s < S; s++) prev[s] = (A)__ldg(par + (long long)s * a.ld + src);
    } else {
#pragma unroll
      for (int s = 0; s < S; s++) prev[s] = A(0);
    }
#pragma unroll
    for (int s = 0; s < S; s++) st[s] = (A)cur[(long long)s * a.ld + i];
    A ua;
    if (PREDRAWN) {
#pragma unroll
      for (int d = 0; d < NN; d++) eps[d] = (A)(R)a.normals[(long long)d * a.ld + i];
#pragma unroll
      for (int d = 0; d < NU; d++) us[d] = (A)(R)a.uniforms[(long long)d * a.ld + i];
      ua = (A)(R)a.uniforms[(long long)NU * a.ld + i];
    } else {
#pragma unroll
      for (int d = 0; d < NN; d += 2) {
        R e0, e1;
        philox_normal_pair(a.seed, pid, a.step, base + (uint32_t)(d >> 1), T, e0, e1);
        eps[d] = (A)e0;
        if (d + 1 < NN) eps[d + 1] = (A)e1;
      }
      R u0, u1;
      philox_uniform_pair(a.seed, pid, a.step, base + 15u, u0, u1);
      ua = (A)u0;
      if (NU > 0) us[0] = (A)u1;
    }
    A alpha;
    if (MOVE == MH_RESIMULATE) {
      const A w_new = Tr::apply(a.mp, h, T, prev, next, eps, us);
      alpha = w_new - Tr::obs_logpdf(a.mp, h, T, st);                       // regenerate.jl:47-50
    } else {
      const A fwd_minus_bwd = Tr::drift(a.mp, h, prev, st, next, eps, dsd, ldsd);
      A w = Tr::latent_logpdf(a.mp, h, T, prev, next) - Tr::latent_logpdf(a.mp, h, T, prev, st);   // update.jl:46-50
      w += Tr::obs_logpdf(a.mp, h, T, next) - Tr::obs_logpdf(a.mp, h, T, st);
      alpha = w - fwd_minus_bwd;                                            // mh.jl:49
    }
    if (a.alpha_out) a.alpha_out[i] = (double)alpha;
    if (Math<A>::log_(ua) < alpha) {                                        // mh.jl:20, :52 (NaN => reject)
#pragma unroll
      for (int s = 0; s < S; s++) cur[(long long)s * a.ld + i] = (R)next[s];
      nacc++;
    }
  }
  nacc = __reduce_add_sync(0xffffffffu, nacc);
  if ((threadIdx.x & 31) == 0 && nacc) atomicAdd(a.accepted, (unsigned long long)nacc);
}

// Resimulation move for models that only offer Transition<>::apply (GB_MODEL_SPEC programs lowered at run time): the
// score of the retained constrained choices under the CURRENT latents is not recomputed but carried per particle in
// `lik` (initialised from the step's log incremental weights, updated on acceptance), so
//   alpha = (sum of observe scores at the re-drawn latents) - lik[i]            (mh.jl:16-27, regenerate.jl:47-50).
// Valid for programs whose weight is exactly that sum (no proposal-correction terms).
template <typename R, int KIND, bool INIT, bool PREDRAWN>
__global__ void __launch_bounds__(kBlock) mh_resim_lik_kernel(const MhArgs a) {
  typedef typename Arith<R>::type A;     // storage R, arithmetic fp64 (models.cuh)
  using Tr = Transition<A, KIND, PROPOSAL_DEFAULT, INIT>;
  constexpr int S = Tr::S, NN = Tr::NN, NU = Tr::NU;
  __shared__ double s_tab[kMathTabDoubles];
  const MathTables T = stage_math_tables(s_tab);
  __syncthreads();
  const Hoisted<A> h = Tr::hoist(a.mp);
  const R *__restrict__ par = static_cast<const R *>(a.parent);
  R *__restrict__ cur = static_cast<R *>(a.cur);
  R *__restrict__ lik = static_cast<R *>(a.lik);
  const uint32_t base = kMhPairBase + 16u * a.move;
  const bool via_anc = a.ctl ? (a.ctl->do_resample != 0) : (a.anc != nullptr);
  unsigned int nacc = 0;
  for (long long i = (long long)blockIdx.x * kBlock + threadIdx.x; i < a.n; i += (long long)gridDim.x * kBlock) {
    const unsigned long long pid = (unsigned long long)(a.pid_offset + i);
    A prev[S], next[S], eps[NN > 0 ? NN : 1], us[NU > 0 ? NU : 1];
    const long long src = (!INIT && via_anc) ? (long long)a.anc[i] : i;
#pragma unroll
    for (int s = 0; s < S; s++) prev[s] = INIT ? A(0) : (A)__ldg(par + (long long)s * a.ld + src);
    A ua;
    if (PREDRAWN) {
#pragma unroll
      for (int d = 0; d < NN; d++) eps[d] = (A)(R)a.normals[(long long)d * a.ld + i];
#pragma unroll
      for (int d = 0; d < NU; d++) us[d] = (A)(R)a.uniforms[(long long)d * a.ld + i];
      ua = (A)(R)a.uniforms[(long long)NU * a.ld + i];
    } else {
#pragma unroll
      for (int d = 0; d < NN; d += 2) {
        R e0, e1;
        philox_normal_pair(a.seed, pid, a.step, base + (uint32_t)(d >> 1), T, e0, e1);
        eps[d] = (A)e0;
        if (d + 1 < NN) eps[d + 1] = (A)e1;
      }
      // accept draw: first uniform of pair base + 15; the transition's own uniforms: pairs base + 8 ...
      R u0, u1;
      philox_uniform_pair(a.seed, pid, a.step, base + 15u, u0, u1);
      ua = (A)u0;
      if (NU > 0) us[0] = (A)u1;
#pragma unroll
      for (int d = 1; d < NU; d += 2) {
        R v0, v1;
        philox_uniform_pair(a.seed, pid, a.step, base + 8u + (uint32_t)(d >> 1), v0, v1);
        us[d] = (A)v0;
        if (d + 1 < NU) us[d + 1] = (A)v1;
      }
    }
    const A w_new = Tr::apply(a.mp, h, T, prev, next, eps, us);
    const A alpha = w_new - (A)lik[i];
    if (a.alpha_out) a.alpha_out[i] = (double)alpha;
    if (Math<A>::log_(ua) < alpha) {
#pragma unroll
      for (int s = 0; s < S; s++) cur[(long long)s * a.ld + i] = (R)next[s];
      lik[i] = (R)w_new;
      nacc++;
    }
  }
  nacc = __reduce_add_sync(0xffffffffu, nacc);
  if ((threadIdx.x & 31) == 0 && nacc) atomicAdd(a.accepted, (unsigned long long)nacc);
}

// ------------------------------------------------------------------------------------------
// Sliding-window rejuvenation (docs/src/tutorials/smc.md:492-541): one block Metropolis-Hastings move per particle on the
// latent choices of the time steps a..b <= the current step (the tutorial perturbs the velocities of the last five steps),
//     mh(trace, perturbation_proposal, (a, b))  with  {choice at t} ~ normal(current value, drift_std)      (mh.jl:37-58)
// The update weight (dynamic/update.jl:46-62) is the change of the log joint over everything the move touches: the latent
// choices of steps a..b under their parents, the choices of step b+1 under its (changed) parent, and the observations of
// ALL steps a..T whose states move with the perturbed choices (the bearings model accumulates positions from velocities).
// The engine's store is Markov, so the last W states of every particle's own lineage ride along in window columns
// (gb_pf_enable_window; gathered with the particle at every resample, like the copies Gen's persistent traces make): the
// kernel reloads the trajectory t = a-1 .. T, replays it with the proposed choices, and writes it back on acceptance.
// ------------------------------------------------------------------------------------------
constexpr int kWinMax = 8;                      // window depth W (states before the current one kept per particle)
constexpr uint32_t kMhWinBase = 0x900000u;      // Philox draw-index block of the window moves: base + 64 * move + pair
struct WinMhArgs {
  ModelParams mp;
  void *cur;               // current traces (lag 0), S columns of stride ld
  void *win;               // window columns [W][S][ld]: win[k] = the state k+1 steps before the current one
  const double *normals, *uniforms;   // PREDRAWN: normals [nb * ND][ld] (step-major), uniforms [1][ld] (accept draw)
  double *alpha_out;
  double obsw[kWinMax + 1];   // the (scalar) observations of steps a .. T
  long long n, ld, pid_offset;
  unsigned long long seed;
  unsigned int step, move;
  int L, nb, W;            // L = T - a + 1 steps replayed, nb = b - a + 1 of them perturbed
  double drift_sd;
  unsigned long long *accepted;
};

template <typename R, int KIND, bool PREDRAWN>
__global__ void __launch_bounds__(128) window_mh_kernel(const WinMhArgs a) {
  typedef typename Arith<R>::type A;
  using Tr = Transition<A, KIND, PROPOSAL_DEFAULT, false>;
  constexpr int S = Tr::S, ND = Tr::ND;
  __shared__ double s_tab[kMathTabDoubles];
  const MathTables T = stage_math_tables(s_tab);
  __syncthreads();
  ModelParams mp = a.mp;
  const Hoisted<A> h = Tr::hoist(mp);
  R *__restrict__ cur = static_cast<R *>(a.cur);
  R *__restrict__ win = static_cast<R *>(a.win);
  const A dsd = (A)a.drift_sd, ldsd = Math<A>::log_((A)a.drift_sd);
  const uint32_t base = kMhWinBase + 64u * a.move;
  unsigned int nacc = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += (long long)gridDim.x * blockDim.x) {
    const unsigned long long pid = (unsigned long long)(a.pid_offset + i);
    A told[kWinMax + 1][S], tnew[kWinMax + 1][S];      // trajectory t = a-1 .. T: index j <-> lag L - j
    for (int j = 0; j <= a.L; j++) {
      const int lag = a.L - j;
      for (int s = 0; s < S; s++)
        told[j][s] = lag == 0 ? (A)cur[(long long)s * a.ld + i] : (A)win[((long long)(lag - 1) * S + s) * a.ld + i];
    }
    for (int s = 0; s < S; s++) tnew[0][s] = told[0][s];
    A w = A(0), fb = A(0);
    for (int j = 1; j <= a.L; j++) {
      mp.obs0 = a.obsw[j - 1];
      A eps[ND > 0 ? ND : 1];
      if (j <= a.nb) {
        for (int d = 0; d < ND; d++) {
          const int idx = (j - 1) * ND + d;
          if (PREDRAWN) eps[d] = (A)(R)a.normals[(long long)idx * a.ld + i];
          else {
            R e0, e1;
            philox_normal_pair(a.seed, pid, a.step, base + (uint32_t)(idx >> 1), T, e0, e1);
            eps[d] = (A)((idx & 1) ? e1 : e0);
          }
        }
        fb += Tr::drift(mp, h, tnew[j - 1], told[j], tnew[j], eps, dsd, ldsd);         // forward - backward proposal score (mh.jl:44-48)
      } else {
        for (int d = 0; d < ND; d++) eps[d] = A(0);
        (void)Tr::drift(mp, h, tnew[j - 1], told[j], tnew[j], eps, A(0), A(0));       // replay: same choices, moved parent
      }
      w += Tr::latent_logpdf(mp, h, T, tnew[j - 1], tnew[j]) - Tr::latent_logpdf(mp, h, T, told[j - 1], told[j]);   // update.jl:46-50
      w += Tr::obs_logpdf(mp, h, T, tnew[j]) - Tr::obs_logpdf(mp, h, T, told[j]);
    }
    const A alpha = w - fb;                                                            // mh.jl:49
    A ua;
    if (PREDRAWN) ua = (A)(R)a.uniforms[i];
    else {
      R u0, u1;
      philox_uniform_pair(a.seed, pid, a.step, base + 63u, u0, u1);
      ua = (A)u0;
    }
    if (a.alpha_out) a.alpha_out[i] = (double)alpha;
    if (Math<A>::log_(ua) < alpha) {                                                   // mh.jl:52 (NaN => reject)
      for (int j = 1; j <= a.L; j++) {
        const int lag = a.L - j;
        for (int s = 0; s < S; s++) {
          if (lag == 0) cur[(long long)s * a.ld + i] = (R)tnew[j][s];
          else win[((long long)(lag - 1) * S + s) * a.ld + i] = (R)tnew[j][s];
        }
      }
      nacc++;
    }
  }
  nacc = __reduce_add_sync(0xffffffffu, nacc);
  if ((threadIdx.x & 31) == 0 && nacc) atomicAdd(a.accepted, (unsigned long long)nacc);
}

// window maintenance after a step / an eager gather: new_win[0] = parent's state, new_win[k] = parent's win[k-1]
// (shift = 1, after particle_filter_step!), or new_win[k] = win[k] of the ancestor (shift = 0, eager resample gather)
template <typename R>
__global__ void __launch_bounds__(kBlock) window_shift_kernel(const R *__restrict__ parent_state, const R *__restrict__ win_old,
                                                              R *__restrict__ win_new, const int *__restrict__ anc, int S, int W,
                                                              int shift, long long n, long long ld) {
  for (long long i = (long long)blockIdx.x * kBlock + threadIdx.x; i < n; i += (long long)gridDim.x * kBlock) {
    const long long src = anc ? (long long)anc[i] : i;
    for (int k = 0; k < W; k++)
      for (int s = 0; s < S; s++) {
        R v;
        if (shift && k == 0) v = parent_state[(long long)s * ld + src];
        else v = win_old[((long long)(k - shift) * S + s) * ld + src];
        win_new[((long long)k * S + s) * ld + i] = v;
      }
  }
}

}  // namespace gb
