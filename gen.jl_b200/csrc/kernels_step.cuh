// kernels_step.cuh -- K1/K2/K3 of the particle path.
//
//   step_kernel   : fused (deferred resample gather +) propose + weight + log-weight accumulation +
//                   running max, one launch per particle_filter_step! / initialize_particle_filter /
//                   importance_sampling call (particle_filter.jl:127-131, :149-151, :196-200, :224-232,
//                   :264-267; importance.jl:28-31, :48-54).
//   wstats_kernel : one pass over the log weights that yields everything maybe_resample! /
//                   log_ml_estimate need: sum exp(lw-m), sum exp(2(lw-m)) (inference.jl:3-6,
//                   particle_filter.jl:3-12) AND the per-tile totals of the integer-quantised weights the
//                   resamplers sweep over -- the exp is evaluated once and shared.
//
// Both are HBM-streaming kernels: SoA columns, 128-bit loads/stores, grid = resident CTAs x 148 SMs,
// grid-stride.  The last CTA to finish folds the per-CTA partials (ticket counter), so there is no
// separate reduction launch.  v0 of the step kernel was issue-bound at 487 instructions per particle
// (profiles/r1_step_kernel_v0.md); the exp moved into wstats (where the quantiser needs it anyway),
// the transcendental functions are the table-driven ones of gbmath.cuh.
#pragma once
#include "models.cuh"

namespace gb {

struct XchgCtl;
struct WeightStats {
  double m, s1, s2;   // reference max, sum exp(lw-m), sum exp(2(lw-m)) of this shard
  double lse, ess;    // m + log s1 ; s1^2 / s2
  unsigned long long q_total;   // sum of the quantised weights (wstats)
  XchgCtl *xc;        // sharded filters: the peer-memory exchange below (nullptr: single shard).  NOT part of the record.
  int incr_in_logw;   // device-decided steps (GATHER == 2): 1 = the last step kernel left its log incremental weights in logw
};
constexpr int kStatsRecordBytes = 48;   // {m, s1, s2, lse, ess, q_total}: what is copied to the host / to a clone

// ------------------------------------------------------------------------------------------
// Device-side exchange between the shards of one filter (one shard per GPU; NVLink / NVSwitch peer memory).
// maybe_resample! over G shards needs the global max of the log weights, the per-shard (sum e, sum e^2, integer CDF mass)
// and a barrier after the offspring rows have been written into their owners' memory.  Instead of three NCCL collectives
// per step (launch + protocol latency of ~35 us each: a third of the 8-GPU strong-scaling step), the producing kernel's
// last CTA peer-stores its 8..24 byte contribution into EVERY shard's mailbox, followed by a sequence flag, and the
// consuming kernel spins on the G flags of its own (local) mailbox:
//   phase A  publish_max (tail of the step / max kernel)  -> mx[src]            consumed at the start of wstats
//   phase B  wstats' last CTA                             -> (s1, s2, q)[src]   consumed by shard_plan_block (the same CTA's tail,
//                                                                               or shard_plan_kernel for synchronous callers)
//   phase C  overflow_sys_kernel's last CTA (after the     -> done flag          consumed at the top of the next step kernel
//            offspring export)
// Payload slots alternate with the parity of the sequence number; a shard can never be more than one phase ahead of a
// peer (each phase consumes what every peer produced in the previous one), so two slots are enough.  Every shard
// executes the same sequence of kernels (SPMD), so the expected sequence number is the shard's own post counter.
// Spins are bounded (timeout_cycles): a lost peer sets `error` instead of hanging the GPU.
// ------------------------------------------------------------------------------------------
constexpr int kMaxWorld = 64;
struct XchgMail {
  double mx[2][kMaxWorld];
  double s1[2][kMaxWorld], s2[2][kMaxWorld];
  unsigned long long q[2][kMaxWorld];
  unsigned long long mx_seq[kMaxWorld], rec_seq[kMaxWorld], done_seq[kMaxWorld];
};
struct XchgCtl {
  XchgMail *peer[kMaxWorld];      // every shard's mailbox as seen from THIS device (own included)
  XchgMail *mine;
  int world, rank;
  unsigned long long seq_a, seq_b, seq_c;   // sequence number of this shard's last post per phase
  long long timeout_cycles;
  int error;                      // sticky: a wait timed out
  unsigned int export_ticket;
};
GB_D unsigned long long ld_acquire_sys(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
GB_D void st_release_sys(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// flag traffic of the exchange: ONE system-scope fence orders a shard's payload stores before its G flag stores, and one
// after the spin orders the flag before the payload loads -- the flags themselves are relaxed accesses.  (A release store per
// peer and an acquire load per probe put a system-scope fence on every one of them: G fences per post, one per probe.)
GB_D void st_relaxed_sys(unsigned long long *p, unsigned long long v) {
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
GB_D unsigned long long ld_relaxed_sys(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
// wait until *flag >= seq (flag lives in this shard's own memory; a peer writes it)
GB_D void xchg_spin(const unsigned long long *flag, unsigned long long seq, XchgCtl *xc) {
  const long long t0 = clock64();
  while (ld_relaxed_sys(flag) < seq) {
    if (clock64() - t0 > xc->timeout_cycles) { xc->error = 1; break; }
    __nanosleep(20);
  }
  __threadfence_system();
}
// phase C wait, at the top of every kernel that reads a sharded filter's traces (call from all threads of the CTA, after
// pdl_wait): the offspring rows its peers exported in the resample before this step have landed in this shard's extension
// rows.  Every peer's "done" counter must have reached this shard's own export count -- all shards export in lock step, and
// when nothing was exported since the last wait the test passes at once.
GB_D void xchg_wait_done(WeightStats *stats) {
  XchgCtl *xc = stats->xc;
  if (!xc) return;
  const int world = xc->world;
  if (world > 1 && (int)threadIdx.x < world) xchg_spin(&xc->mine->done_seq[threadIdx.x], __ldcg(&xc->seq_c), xc);
  __syncthreads();
}
// phase A post (one thread): this shard's max of the log weights into every mailbox
GB_D void xchg_post_max(XchgCtl *xc, double m) {
  const unsigned long long seq = xc->seq_a + 1;
  xc->seq_a = seq;
  const int slot = (int)(seq & 1), me = xc->rank;
  for (int r = 0; r < xc->world; r++) xc->peer[r]->mx[slot][me] = m;
  __threadfence_system();
  for (int r = 0; r < xc->world; r++) st_relaxed_sys(&xc->peer[r]->mx_seq[me], seq);
}

// Device-side control block of the fused multi-step entry (gb_pf_run): the ESS test of maybe_resample!
// (particle_filter.jl:256), the running log-ML (:263) and the gathered/direct choice of the next step live
// on the device, so T steps run without a host round trip.
struct RunCtl {
  int do_resample;        // decision of the current iteration
  int n_resampled;        // resamples so far in this run
  double log_ml_est;      // ParticleFilterState.log_ml_est
  double ess_threshold;
  double log_n;           // log(num_particles)
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
// plain (L2-allocating) load for data another pass re-reads
template <typename R> GB_D void vload_keep(const R *p, R (&out)[VecOf<R>::V]);
template <> GB_D void vload_keep<double>(const double *p, double (&o)[2]) {
  double2 v = *reinterpret_cast<const double2 *>(p);
  o[0] = v.x; o[1] = v.y;
}
template <> GB_D void vload_keep<float>(const float *p, float (&o)[4]) {
  float4 v = *reinterpret_cast<const float4 *>(p);
  o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
}
// plain (L2-allocating) store for data the very next kernel re-reads
template <typename R> GB_D void vstore_keep(R *p, const R (&in)[VecOf<R>::V]);
template <> GB_D void vstore_keep<double>(double *p, const double (&i)[2]) {
  *reinterpret_cast<double2 *>(p) = make_double2(i[0], i[1]);
}
template <> GB_D void vstore_keep<float>(float *p, const float (&i)[4]) {
  *reinterpret_cast<float4 *>(p) = make_float4(i[0], i[1], i[2], i[3]);
}

// particles per thread of the step kernel (compile-time tuning knobs; 128-bit accesses for fp64 at the default; fp32
// filters run the same fp64 transition, so two particles per thread (64-bit accesses) keep them free of register spills)
#ifndef GB_STEP_V64
#define GB_STEP_V64 2
#endif
#ifndef GB_STEP_V32
#define GB_STEP_V32 2
#endif
template <typename R> struct StepV { static constexpr int V = GB_STEP_V32; };
template <> struct StepV<double> { static constexpr int V = GB_STEP_V64; };
template <typename R, int V> GB_D void svload(const R *p, R (&o)[V]) {
  if constexpr (V == VecOf<R>::V) vload<R>(p, o);
  else if constexpr (sizeof(R) == 4 && V == 2) { float2 t = __ldcs(reinterpret_cast<const float2 *>(p)); o[0] = t.x; o[1] = t.y; }
  else { for (int k = 0; k < V; k++) o[k] = __ldcs(p + k); }
}
template <typename R, int V> GB_D void svstore(R *p, const R (&o)[V]) {
  if constexpr (V == VecOf<R>::V) vstore<R>(p, o);
  else if constexpr (sizeof(R) == 4 && V == 2) __stcs(reinterpret_cast<float2 *>(p), make_float2(o[0], o[1]));
  else { for (int k = 0; k < V; k++) __stcs(p + k, o[k]); }
}
template <typename R, int V> GB_D void svstore_keep(R *p, const R (&o)[V]) {
  if constexpr (V == VecOf<R>::V) vstore_keep<R>(p, o);
  else if constexpr (sizeof(R) == 4 && V == 2) *reinterpret_cast<float2 *>(p) = make_float2(o[0], o[1]);
  else { for (int k = 0; k < V; k++) p[k] = o[k]; }
}

struct StepArgs {
  const void *state_in;   // S columns of stride ld (traces); unused when INIT
  void *state_out;        // S columns of stride ld (new_traces)
  void *logw;             // ld elements (in/out)
  void *incr;             // ld elements (log incremental weights), unused when INIT
  const int *anc;         // GATHER: local source index of every slot (deferred resample gather)
  const RunCtl *ctl;      // GATHER == 2: gathered iff ctl->do_resample (decided on the device)
  const double *normals;  // PREDRAWN: [NN][ld] ; else NULL
  const double *uniforms; // PREDRAWN: [NU][ld]
  long long n, ld, pid_offset;
  unsigned long long seed;
  unsigned int step;
  double *partials;       // one running max per CTA
  unsigned int *ticket;
  WeightStats *stats;     // stats->m receives the max of the new log weights
  // After a resample the log weights are reset to exactly 0.0 (particle_filter.jl:266), so the step's new log weights ARE its
  // log incremental weights (0.0 + w == w): with share_incr the gathered step stores them once (logw) and the host reads
  // gb_pf_get_log_increments from there -- 8 of the kernel's 84 bytes per particle.
  int share_incr;
  ModelParams mp;
};

template <typename T> GB_D void prefetch_l2(const T *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// maximum() with Julia's NaN propagation (inference.jl:4)
GB_D double max_nan(double a, double b) { return (a != a) ? a : ((b != b) ? b : (a > b ? a : b)); }

GB_D double block_max(double v, double *smem /* kBlock/32 */) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = max_nan(v, __shfl_xor_sync(0xffffffffu, v, o));
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) smem[warp] = v;
  __syncthreads();
  double r = -INFINITY;
  if (warp == 0) {
    if (lane < (int)(blockDim.x >> 5)) r = smem[lane];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) r = max_nan(r, __shfl_xor_sync(0xffffffffu, r, o));
  }
  return r;  // valid in warp 0
}

// Publish this CTA's max; the last CTA reduces all of them into stats->m.
GB_D void publish_max(double acc, double *partials, unsigned int *ticket, WeightStats *stats) {
  __shared__ double s_red[32];
  __shared__ bool s_last;
  double r = block_max(acc, s_red);
  if (threadIdx.x == 0) {
    partials[blockIdx.x] = r;
    __threadfence();
    unsigned int t = atomicAdd(ticket, 1u);
    s_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  double a = -INFINITY;
  for (unsigned int i = threadIdx.x; i < gridDim.x; i += blockDim.x) a = max_nan(a, __ldcg(partials + i));
  __syncthreads();
  a = block_max(a, s_red);
  if (threadIdx.x == 0) {
    stats->m = a;
    *ticket = 0;
    XchgCtl *xc = stats->xc;
    if (xc) xchg_post_max(xc, a);     // sharded filter: phase A of the peer-memory exchange
  }
}
// the same post for a max that is known without a pass over the weights (all zero after a resample, host-cached)
__global__ void xchg_post_max_kernel(WeightStats *stats) {
  if (threadIdx.x == 0 && stats->xc) xchg_post_max(stats->xc, stats->m);
}

#ifndef GB_STEP_MINB
#define GB_STEP_MINB 3
#endif
// Shared-memory staging of the deferred gather: every thread copies the ancestor rows of its NEXT iteration into its own
// shared-memory slots with asynchronous copies (cp.async: global -> shared without passing through registers, a deep
// queue per SM) while the current iteration computes, and reads them back one iteration later.  0: the r1 scheme (L2
// prefetch of the next rows, ld.global.nc at the point of use).
// Measured (profiles/README.md, r2i / r2ai / r2aj): with the 84-byte kernel of r2i the staged form was 1 % SLOWER (1.5515 vs
// 1.5345 ms at N = 1e8); once the gathered step stopped writing the increments twice (76 bytes) it is faster for every model
// and dtype -- N = 2e7: bearings 0.281 vs 0.291 ms (fp32 0.242 vs 0.259), SV 0.222 vs 0.227 (0.174 vs 0.183), LGSSM 0.185 vs
// 0.191 (0.144 vs 0.151); N = 1e8 bearings fp64 1.39 vs 1.41 ms.  ncu: the long-scoreboard stall (1.66 warps per issue slot)
// is gone, what is left is the fixed-latency dependency chain of the fp64 transition ("wait" 3.0) at 64 % issue-active.
#ifndef GB_STEP_STAGE
#define GB_STEP_STAGE 1
#endif
template <int BYTES> GB_D void cp_async(void *smem_dst, const void *gmem_src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(d), "l"(gmem_src), "n"(BYTES) : "memory");
}
GB_D void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> GB_D void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
// The body of the step kernel, specialised on whether the step reads through the ancestors (GATHER 1) or not (0).  DEVCTL:
// the choice was taken from the device flag by the kernel below (it is then recorded for the host, see share_incr).
template <typename R, int KIND, int PK, bool INIT, bool PREDRAWN, int GATHER /* 0 no, 1 yes */, bool DEVCTL>
GB_D void step_body(const StepArgs &a, const MathTables &T) {
  // R is the STORAGE type (traces, log weights, noise); the transition itself (propose + logpdf) is always evaluated in
  // fp64 on the widened values (models.cuh "Arith"), so a GB_F32 filter differs from the fp64 one only by the rounding of
  // what it stores.
  typedef typename Arith<R>::type A;
  typedef Transition<A, KIND, PK, INIT> Tr;
  constexpr int V = StepV<R>::V, S = Tr::S, NN = Tr::NN, NU = Tr::NU;
  // sharded filter: peers' offspring rows are in (no-op otherwise).  (A lazy variant -- only the warps whose slots another
  // shard's sweep produced wait, right before their first ancestor load -- hides nothing: the particles are dealt to the
  // threads statically, so a warp that starts late finishes late by the same amount, and the extra live registers made the
  // fp64 bearings instantiation spill: 2-GPU step kernel 0.756 -> 0.834 ms, profiles/README.md r2v.  Rotating the iteration
  // order of the affected warps so that their wait falls at the END of their work was tried too, r2an: the modular iteration
  // index costs the whole kernel 5-8 % on ONE GPU -- 1.39 -> 1.49 ms at N = 1e8 -- far more than the wait it would hide.)
  if (!INIT && (GATHER != 0 || DEVCTL)) xchg_wait_done(a.stats);
  const R *__restrict__ sin = static_cast<const R *>(a.state_in);
  R *__restrict__ sout = static_cast<R *>(a.state_out);
  R *__restrict__ logw = static_cast<R *>(a.logw);
  R *__restrict__ incr = static_cast<R *>(a.incr);
  const Hoisted<A> hz = Tr::hoist(a.mp);
  double acc = -INFINITY;

  const long long nvec = (a.n + V - 1) / V;
  const long long stride = (long long)gridDim.x * kBlock;
  long long iv = (long long)blockIdx.x * kBlock + threadIdx.x;
  // Software pipeline for the indirection anc -> state (two dependent DRAM round trips per iteration in
  // v0, 51% long-scoreboard stalls): the ancestor indices are loaded two iterations ahead into registers
  // and the ancestor rows of the next iteration are prefetched into L2 while this one computes.
  int src_cur[V], src_nxt[V];
  auto load_anc = [&](long long ivx, int (&dst)[V]) {
    if constexpr (V == 1) { dst[0] = __ldcs(a.anc + ivx); }
    else if constexpr (V == 2) { int2 t = __ldcs(reinterpret_cast<const int2 *>(a.anc + ivx * V)); dst[0] = t.x; dst[1] = t.y; }
    else { int4 t = __ldcs(reinterpret_cast<const int4 *>(a.anc + ivx * V)); dst[0] = t.x; dst[1] = t.y; dst[2] = t.z; dst[3] = t.w; }
  };
  constexpr bool gather = GATHER == 1;
  // staged gather: [stage][field][v][thread] -- a thread only ever touches its own column, so the asynchronous copies
  // need no barrier, just cp.async.wait_group
  constexpr bool STAGE = GB_STEP_STAGE && !INIT && GATHER != 0 && S > 0 && (2 * (S > 0 ? S : 1) * V * (int)sizeof(R) * kBlock <= 40 * 1024);
  __shared__ __align__(16) R s_stage[STAGE ? 2 : 1][STAGE ? S : 1][STAGE ? V : 1][STAGE ? kBlock : 1];
  auto stage_rows = [&](int st, const int (&src)[V]) {
#pragma unroll
    for (int s = 0; s < S; s++)
#pragma unroll
      for (int v = 0; v < V; v++) cp_async<(int)sizeof(R)>(&s_stage[STAGE ? st : 0][STAGE ? s : 0][STAGE ? v : 0][STAGE ? threadIdx.x : 0], sin + (long long)s * a.ld + src[v]);
    cp_async_commit();
  };
  int stage = 0;
  if (!INIT && gather) {
    if (iv < nvec) load_anc(iv, src_cur);
    if (iv + stride < nvec) load_anc(iv + stride, src_nxt);
    if (STAGE && iv < nvec) stage_rows(0, src_cur);
  }
  for (; iv < nvec; iv += stride) {
    const long long i0 = iv * V;
    R prev[S > 0 ? S : 1][V];
    R lw_old[V];
    if (!INIT) {
      if (gather) {
        // deferred maybe_resample! gather (particle_filter.jl:264-267): slot k continues the trace of
        // ancestor anc[k]; its log weight was reset to exactly 0.0
        if constexpr (STAGE) {
          const bool more = iv + stride < nvec;
          if (more) stage_rows(stage ^ 1, src_nxt);          // the next iteration's rows start flowing now
          if (more) cp_async_wait<1>(); else cp_async_wait<0>();   // this iteration's rows (committed one iteration ago) are in
#pragma unroll
          for (int s = 0; s < S; s++)
#pragma unroll
            for (int v = 0; v < V; v++) prev[s][v] = s_stage[stage][s][v][threadIdx.x];
          stage ^= 1;
          if (more) {
#pragma unroll
            for (int v = 0; v < V; v++) src_cur[v] = src_nxt[v];
            if (iv + 2 * stride < nvec) load_anc(iv + 2 * stride, src_nxt);
          }
        } else {
#pragma unroll
          for (int s = 0; s < S; s++)
#pragma unroll
            for (int v = 0; v < V; v++) prev[s][v] = __ldg(sin + (long long)s * a.ld + src_cur[v]);
          if (iv + stride < nvec) {
#pragma unroll
            for (int s = 0; s < S; s++)
#pragma unroll
              for (int v = 0; v < V; v++) prefetch_l2(sin + (long long)s * a.ld + src_nxt[v]);
#pragma unroll
            for (int v = 0; v < V; v++) src_cur[v] = src_nxt[v];
            if (iv + 2 * stride < nvec) load_anc(iv + 2 * stride, src_nxt);
          }
        }
#pragma unroll
        for (int v = 0; v < V; v++) lw_old[v] = R(0);
      } else {
#pragma unroll
        for (int s = 0; s < S; s++) svload<R, V>(sin + (long long)s * a.ld + i0, prev[s]);
        svload<R, V>(logw + i0, lw_old);
        if (iv + stride < nvec) {
#pragma unroll
          for (int s = 0; s < S; s++) prefetch_l2(sin + (long long)s * a.ld + i0 + stride * V);
          prefetch_l2(logw + i0 + stride * V);
        }
      }
    }
    R next[S][V], w[V], lw_new[V];
#pragma unroll
    for (int v = 0; v < V; v++) {
      const long long i = i0 + v;
      A eps[NN > 0 ? NN : 1], us[NU > 0 ? NU : 1];
      if (PREDRAWN) {
#pragma unroll
        for (int d = 0; d < NN; d++) eps[d] = (A)(R)__ldcs(a.normals + (long long)d * a.ld + i);
#pragma unroll
        for (int d = 0; d < NU; d++) us[d] = (A)(R)__ldcs(a.uniforms + (long long)d * a.ld + i);
      } else {
        const unsigned long long pid = (unsigned long long)(a.pid_offset + i);
#pragma unroll
        for (int d = 0; d < NN; d += 2) {
          R e0, e1;                  // the noise itself is drawn in the storage precision
          philox_normal_pair(a.seed, pid, a.step, (unsigned)(d / 2), T, e0, e1);
          eps[d] = (A)e0;
          if (d + 1 < NN) eps[d + 1] = (A)e1;
        }
#pragma unroll
        for (int d = 0; d < NU; d += 2) {
          R u0, u1;
          philox_uniform_pair(a.seed, pid, a.step, (unsigned)(d / 2), u0, u1);
          us[d] = (A)u0;
          if (d + 1 < NU) us[d + 1] = (A)u1;
        }
      }
      A pv[S], nx[S];
#pragma unroll
      for (int s = 0; s < S; s++) pv[s] = INIT ? A(0) : (A)prev[s][v];
      const A wa = Tr::apply(a.mp, hz, T, pv, nx, eps, us);
#pragma unroll
      for (int s = 0; s < S; s++) next[s][v] = (R)nx[s];
      w[v] = (R)wa;
      lw_new[v] = (R)(INIT ? wa : (A)lw_old[v] + wa);     // particle_filter.jl:131,151 / :199,231
      if (i < a.n) acc = max_nan(acc, (double)lw_new[v]);
    }
#pragma unroll
    for (int s = 0; s < S; s++) svstore<R, V>(sout + (long long)s * a.ld + i0, next[s]);
    svstore_keep<R, V>(logw + i0, lw_new);
    if (!INIT && !(gather && a.share_incr)) svstore<R, V>(incr + i0, w);
  }
  if (!INIT && DEVCTL && blockIdx.x == 0 && threadIdx.x == 0) a.stats->incr_in_logw = (gather && a.share_incr) ? 1 : 0;
  publish_max(acc, a.partials, a.ticket, a.stats);
}
// GATHER 2: gathered iff ctl->do_resample, decided on the device (fused / sharded runs).  The kernel branches ONCE into one of
// the two specialised bodies: with the flag as a run-time predicate inside a single body, the fp64 bearings kernel was 12 %
// slower than the GATHER 1 instantiation (N = 5e7: 0.752 vs 0.674 ms, tools/step_probe.py) -- the reason the sharded step
// kernel did not follow the single-GPU one when that got faster.
template <typename R, int KIND, int PK, bool INIT, bool PREDRAWN, int GATHER /* 0 no, 1 yes, 2 ctl->do_resample */>
__global__ void __launch_bounds__(kBlock, GB_STEP_MINB) step_kernel(const StepArgs a) {
  __shared__ double s_tab[kMathTabDoubles];
  pdl_trigger();
  const MathTables T = stage_math_tables(s_tab);   // constants only: overlaps the predecessor's tail
  __syncthreads();
  pdl_wait();
  if constexpr (GATHER == 2) {
    if (a.ctl->do_resample != 0) step_body<R, KIND, PK, INIT, PREDRAWN, 1, true>(a, T);
    else step_body<R, KIND, PK, INIT, PREDRAWN, 0, true>(a, T);
  } else {
    step_body<R, KIND, PK, INIT, PREDRAWN, GATHER, false>(a, T);
  }
}

// max of an existing log-weight array (after gb_pf_set_log_weights, clone, raw-array entries): 128-bit
// loads, four independent vectors in flight per thread (the array is padded to a multiple of 2048)
template <typename R>
__global__ void __launch_bounds__(kBlock) max_kernel(const R *__restrict__ logw, long long n, double *partials,
                                                     unsigned int *ticket, WeightStats *stats) {
  constexpr int V = VecOf<R>::V, U = 4;
  pdl_trigger();
  pdl_wait();
  double acc = -INFINITY;
  const long long nvec = (n + V - 1) / V;
  const long long stride = (long long)gridDim.x * kBlock;
  for (long long iv = (long long)blockIdx.x * kBlock + threadIdx.x; iv < nvec; iv += stride * U) {
    R v[U][V];
#pragma unroll
    for (int u = 0; u < U; u++)
      if (iv + u * stride < nvec) vload_keep<R>(logw + (iv + u * stride) * V, v[u]);
#pragma unroll
    for (int u = 0; u < U; u++)
      if (iv + u * stride < nvec) {
#pragma unroll
        for (int k = 0; k < V; k++)
          if ((iv + u * stride) * V + k < n) acc = max_nan(acc, (double)v[u][k]);
      }
  }
  publish_max(acc, partials, ticket, stats);
}

// Everything the ancestors pass needs to know about the global sweep; written by the super-tile scan.
struct ResamplePlan {
  SweepParams sp;
  uint64_t q_offset;     // integer CDF mass of the shards before this one
  long long slot_lo;     // first / one-past-last global output slot filled by this shard's particles
  long long slot_hi;
  long long n_global;
};

GB_D uint64_t shfl_up_u64(uint64_t v, int d) {
  uint32_t lo = __shfl_up_sync(0xffffffffu, (uint32_t)v, d);
  uint32_t hi = __shfl_up_sync(0xffffffffu, (uint32_t)(v >> 32), d);
  return ((uint64_t)hi << 32) | lo;
}

// Exclusive scan of the super-tile totals by ONE CTA (any multiple of 32 threads up to 1024) + the sweep plan.
// Runs as its own tiny kernel (superscan_kernel) or, in the fused multi-step run, in the tail of the statistics
// kernel's last CTA, which saves a launch per step.  Leaves super_q zeroed for the next statistics pass.
GB_D void superscan_block(long long nsuper, unsigned long long *__restrict__ super_q, unsigned long long *__restrict__ super_excl,
                          ResamplePlan *plan, uint64_t q_offset, uint64_t q_total_given, long long n_global, uint32_t u0,
                          unsigned int *overflow_count, int plan_preset, uint64_t *s_w /* 33 */) {
  const int T = blockDim.x;
  const long long per = (nsuper + T - 1) / T;
  const long long lo = (long long)threadIdx.x * per, hi = lo + per < nsuper ? lo + per : nsuper;
  uint64_t a = 0;
  for (long long i = lo; i < hi; i++) a += __ldcg(super_q + i);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = T >> 5;
  uint64_t inc = a;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    uint64_t o = shfl_up_u64(inc, d);
    if (lane >= d) inc += o;
  }
  if (lane == 31) s_w[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    uint64_t w = lane < nwarp ? s_w[lane] : 0, winc = w;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      uint64_t o = shfl_up_u64(winc, d);
      if (lane >= d) winc += o;
    }
    s_w[lane] = winc - w;
    if (lane == 31) s_w[32] = winc;
  }
  __syncthreads();
  uint64_t run = s_w[warp] + (inc - a);
  const uint64_t total = s_w[32];
  for (long long i = lo; i < hi; i++) {
    uint64_t v = __ldcg(super_q + i);
    super_excl[i] = run;
    super_q[i] = 0;              // leave the accumulator clean for the next statistics pass (no memset launch)
    run += v;
  }
  if (threadIdx.x == 0) {
    super_excl[nsuper] = total;
    if (!plan_preset) {     // (preset: shard_plan_kernel already wrote the global sweep parameters)
      SweepParams sp;
      sp.q_total = q_total_given ? q_total_given : total;
      sp.q_inv = 1.0 / (double)sp.q_total;
      sp.dscale = (uint64_t)n_global << 32;
      sp.ratio = (double)n_global / (double)sp.q_total;
      sp.u0 = u0;
      plan->sp = sp;
      plan->q_offset = q_offset;
      plan->n_global = n_global;
      plan->slot_lo = slots_below(q_offset, sp);
      plan->slot_hi = slots_below(q_offset + total, sp);
    }
    *overflow_count = 0;
  }
}

struct PeerMem {           // a peer shard's arrays (device pointers valid in THIS process)
  void *state[2];          // its two trace buffers
  int *anc;
  long long *ext_parents;
  long long ld, n_pad, ext_cap;
};
struct ShardPlanDev {
  PeerMem peer[kMaxWorld];
  long long send_lo[kMaxWorld], send_hi[kMaxWorld];   // global slots produced here that shard d owns
  long long ext_off[kMaxWorld];                       // first extension row at shard d for the rows sent from here
  int world, rank;
  int overflow;                                       // sticky: a capacity was exceeded, results are invalid
  WeightStats global;                                 // {m, S1, S2, lse, ess, Q} over ALL shards (what the host reads)
  unsigned long long q_shard[kMaxWorld];              // every shard's integer CDF mass (host-side capacity check)
};

// Waits for every shard's record (phase B), takes the maybe_resample! decision (particle_filter.jl:256; ess_threshold NaN: the
// run's own threshold in ctl), updates the log-ML estimate (:263), writes the sweep + exchange plan and, when a sweep follows
// (do_scan), the exclusive scan of this shard's super-tile totals.  One CTA (any multiple of 32 threads, >= kMaxWorld): its own
// kernel (shard_plan_kernel, synchronous callers) or the tail of the statistics kernel's last CTA (device-controlled run: one
// launch and one kernel boundary less per step).
GB_D void shard_plan_block(WeightStats *stats, const long long *__restrict__ offs, long long n_global, uint32_t u0,
                           double ess_threshold, int do_scan, long long spill_cap, RunCtl *ctl,
                           ResamplePlan *plan, ShardPlanDev *sh, long long nsuper,
                           unsigned long long *__restrict__ super_q,
                           unsigned long long *__restrict__ super_excl, unsigned int *overflow_count) {
  __shared__ SweepParams s_sp;
  __shared__ unsigned long long s_pref[kMaxWorld + 1];
  __shared__ long long s_lo[kMaxWorld], s_hi[kMaxWorld];
  __shared__ int s_do;
  __shared__ uint64_t s_w[33];
  // the scan of this shard's own super-tile totals needs nothing from the peers: it runs while their records are in flight
  if (do_scan) {
    superscan_block(nsuper, super_q, super_excl, plan, 0, 0, n_global, u0, overflow_count, 1, s_w);
    __syncthreads();
  }
  XchgCtl *xc = stats->xc;
  const int world = xc->world, rank = xc->rank;
  const int r = threadIdx.x;
  const unsigned long long seq = __ldcg(&xc->seq_b);
  if (r < world) xchg_spin(&xc->mine->rec_seq[r], seq, xc);
  __syncthreads();
  const int slot = (int)(seq & 1);
  // the G records are fetched by G threads at once (one L2 round trip instead of 3 G dependent ones on thread 0) ...
  __shared__ double s_s1[kMaxWorld], s_s2[kMaxWorld];
  __shared__ unsigned long long s_qk[kMaxWorld];
  if (r < world) {
    s_s1[r] = __ldcg(&xc->mine->s1[slot][r]);
    s_s2[r] = __ldcg(&xc->mine->s2[slot][r]);
    const unsigned long long qk = __ldcg(&xc->mine->q[slot][r]);
    s_qk[r] = qk;
    sh->q_shard[r] = qk;
  }
  __syncthreads();
  if (r == 0) {
    const double m = __ldcg(&stats->m);        // the GLOBAL max: wstats folded phase A and left it here
    double S1 = 0.0, S2 = 0.0;
    unsigned long long run = 0;
    for (int k = 0; k < world; k++) {          // ... and summed in fixed rank order: identical on every shard
      S1 += s_s1[k];
      S2 += s_s2[k];
      s_pref[k] = run;
      run += s_qk[k];
    }
    s_pref[world] = run;
    const double lse = (m == -INFINITY) ? -INFINITY : m + log(S1);   // inference.jl:5
    const double ess = (S1 * S1) / S2;                                // particle_filter.jl:3-6
    const double thr = ess_threshold == ess_threshold ? ess_threshold : ctl->ess_threshold;
    const int flag = (ess < thr && run > 0) ? 1 : 0;                  // :256
    ctl->do_resample = flag;
    if (flag) { ctl->log_ml_est += lse - ctl->log_n; ctl->n_resampled += 1; }   // :263
    s_do = flag;
    sh->global.m = m; sh->global.s1 = S1; sh->global.s2 = S2; sh->global.lse = lse; sh->global.ess = ess; sh->global.q_total = run;
    SweepParams sp;
    sp.q_total = run ? run : 1;
    sp.q_inv = 1.0 / (double)sp.q_total;
    sp.dscale = (uint64_t)n_global << 32;
    sp.ratio = (double)n_global / (double)sp.q_total;
    sp.u0 = u0;
    s_sp = sp;
  }
  __syncthreads();
  if (r < world) {
    s_lo[r] = slots_below(s_pref[r], s_sp);
    s_hi[r] = slots_below(s_pref[r + 1], s_sp);
  }
  __syncthreads();
  if (r == 0) {
    plan->sp = s_sp;
    plan->q_offset = s_pref[rank];
    plan->n_global = n_global;
    plan->slot_lo = s_lo[rank];
    plan->slot_hi = s_hi[rank];
    if (s_do && s_hi[rank] - s_lo[rank] > spill_cap) sh->overflow = 1;
  }
  if (r < world) {
    // rows going from this shard to shard r, and where they start in r's extension rows: after the rows r receives
    // from lower-ranked sources (every shard evaluates the same arithmetic, so all agree)
    long long a = s_lo[rank] > offs[r] ? s_lo[rank] : offs[r], b = s_hi[rank] < offs[r + 1] ? s_hi[rank] : offs[r + 1];
    if (b < a) b = a;
    long long before = 0, total = 0;
    for (int k = 0; k < world; k++) {
      if (k == r) continue;
      long long c = s_lo[k] > offs[r] ? s_lo[k] : offs[r], d = s_hi[k] < offs[r + 1] ? s_hi[k] : offs[r + 1];
      const long long cnt = d > c ? d - c : 0;
      if (k < rank) before += cnt;
      total += cnt;
    }
    if (r != rank && s_do && total > sh->peer[r].ext_cap) { sh->overflow = 1; b = a; }   // r cannot park its incoming rows
    sh->send_lo[r] = a; sh->send_hi[r] = b;
    sh->ext_off[r] = before;
  }
}
__global__ void __launch_bounds__(1024) shard_plan_kernel(WeightStats *stats, const long long *__restrict__ offs, long long n_global, uint32_t u0,
                                                          double ess_threshold, int do_scan, long long spill_cap, RunCtl *ctl,
                                                          ResamplePlan *plan, ShardPlanDev *sh, long long nsuper,
                                                          unsigned long long *__restrict__ super_q,
                                                          unsigned long long *__restrict__ super_excl, unsigned int *overflow_count) {
  pdl_trigger();
  pdl_wait();
  shard_plan_block(stats, offs, n_global, u0, ess_threshold, do_scan, spill_cap, ctl, plan, sh, nsuper, super_q, super_excl, overflow_count);
}

// the sweep plan computed in the tail of the statistics kernel (plan == nullptr: not fused)
struct FusedPlan {
  ResamplePlan *plan;
  unsigned long long *super_excl;
  unsigned int *overflow_count;
  long long n_global;
  uint32_t u0;
  // sharded filter (device-controlled run): the cross-shard plan instead of the local one
  ShardPlanDev *sh;
  const long long *offs;
  long long spill_cap;
  RunCtl *sh_ctl;
};

// ------------------------------------------------------------------------------------------
// wstats: e_i = exp(lw_i - m) once per particle ->
//   s1 = sum e, s2 = sum e^2            (logsumexp / ESS, fp64)
//   tile_q[t] = sum over tile t of floor(e_i 2^bits)   (integer CDF mass per 2048-particle tile)
//   super_q[t / 64] += tile_q[t]                       (per 131072-particle super-tile, integer atomics)
// m comes from device memory (stats->m, written by the step kernel) or, for sharded filters, from the
// host (the global max after the all-reduce).
// ------------------------------------------------------------------------------------------
constexpr int kSuper = 64;   // tiles per super-tile

template <typename R> GB_D void load_items(const R *__restrict__ logw, long long base, long long n, double (&lw)[kItems]);
template <> GB_D void load_items<double>(const double *__restrict__ logw, long long base, long long n, double (&lw)[kItems]) {
  const long long i0 = base + (long long)threadIdx.x * kItems;
  const bool full = base + kTile <= n;
#pragma unroll
  for (int j = 0; j < kItems; j += 2) {
    // arrays are padded to a multiple of kTile elements, so the vector load is always in bounds
    double2 v = *reinterpret_cast<const double2 *>(logw + i0 + j);
    lw[j] = (full || i0 + j < n) ? v.x : -INFINITY;
    lw[j + 1] = (full || i0 + j + 1 < n) ? v.y : -INFINITY;
  }
}
template <> GB_D void load_items<float>(const float *__restrict__ logw, long long base, long long n, double (&lw)[kItems]) {
  const long long i0 = base + (long long)threadIdx.x * kItems;
  const bool full = base + kTile <= n;
#pragma unroll
  for (int j = 0; j < kItems; j += 4) {
    float4 v = *reinterpret_cast<const float4 *>(logw + i0 + j);
    lw[j] = (full || i0 + j < n) ? (double)v.x : -INFINITY;
    lw[j + 1] = (full || i0 + j + 1 < n) ? (double)v.y : -INFINITY;
    lw[j + 2] = (full || i0 + j + 2 < n) ? (double)v.z : -INFINITY;
    lw[j + 3] = (full || i0 + j + 3 < n) ? (double)v.w : -INFINITY;
  }
}

// 64-bit warp sum
GB_D uint64_t warp_sum_u64(uint64_t v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    uint32_t lo = __shfl_xor_sync(0xffffffffu, (uint32_t)v, o), hi = __shfl_xor_sync(0xffffffffu, (uint32_t)(v >> 32), o);
    v += ((uint64_t)hi << 32) | lo;
  }
  return v;
}
GB_D uint64_t block_sum_u64(uint64_t v, uint64_t *smem /* kBlock/32 */) {
  v = warp_sum_u64(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) smem[warp] = v;
  __syncthreads();
  uint64_t r = 0;
#pragma unroll
  for (int k = 0; k < kBlock / 32; k++) r += smem[k];
  return r;  // valid in every thread
}

// Each WARP owns whole 2048-particle tiles (no CTA barrier and no atomics on the tile totals in the streaming loop):
// striped 128-bit loads in batches of four per lane, the next batch in flight while the current one is evaluated, one exp
// per particle, ONE 64-bit shuffle reduction of the quantised mass per tile.  (r1: 256-particle segments with a shuffle
// reduction + two atomics each and bounds selects on every element: 65 instructions per particle, 13% of them the selects.)
// WHOLE = false: the unit of work of a warp is one batch (256 fp64 / 512 fp32 particles) and the tile totals are
// accumulated with atomics (tile_q zeroed by the caller) -- mid-size filters, where whole tiles would leave SMs idle.
// (98 registers unconstrained = 2 CTAs/SM; 3 CTAs/SM measured fastest: 0.224 vs 0.236 ms at N = 1e8, 4 CTAs/SM 0.235)
#ifndef GB_WSTATS_MINB
#define GB_WSTATS_MINB 3
#endif
template <typename R, bool WHOLE>
__global__ void __launch_bounds__(kBlock, GB_WSTATS_MINB) wstats_kernel(const R *__restrict__ logw, long long n, double m_arg, int m_from_device,
                                                        int bits, unsigned long long *__restrict__ tile_q,
                                                        unsigned long long *__restrict__ super_q, unsigned long long *__restrict__ qout,
                                                        double2 *partials, unsigned int *ticket, WeightStats *stats, RunCtl *ctl,
                                                        const FusedPlan fp) {
  constexpr int V = VecOf<R>::V;
  constexpr int kB = 4;                              // loads per lane and batch
  constexpr int kPerTile = kTile / (32 * V * kB);    // batches per tile: 8 (fp64) / 4 (fp32)
  constexpr int kNB = WHOLE ? kPerTile : 1;          // batches per unit of work
  __shared__ uint64_t s_u64[kBlock / 32];
  __shared__ uint64_t s_u64_33[33];
  __shared__ double2 s_d2[kBlock / 32];
  __shared__ bool s_last;
  __shared__ double s_m;
  __shared__ double s_wexp[64];
  pdl_trigger();
  const double *wt = stage_wexp_table(s_wexp);       // constants only: overlaps the predecessor's tail
  __syncthreads();
  pdl_wait();
  double m = m_from_device ? __ldcg(&stats->m) : m_arg;
  XchgCtl *xc = m_from_device == 2 ? stats->xc : nullptr;
  if (xc) {
    // phase A of the exchange: every shard's max arrives in this shard's mailbox; the global max is their maximum,
    // taken in rank order by every CTA (identical bits everywhere)
    const unsigned long long seq = __ldcg(&xc->seq_a);
    const int world = xc->world;
    __shared__ double s_mx[kMaxWorld];
    if ((int)threadIdx.x < world) {            // (each waiting thread also fetches the value: one round trip, not G in a row)
      xchg_spin(&xc->mine->mx_seq[threadIdx.x], seq, xc);
      s_mx[threadIdx.x] = __ldcg(&xc->mine->mx[seq & 1][threadIdx.x]);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      double g = -INFINITY;
      for (int r = 0; r < world; r++) g = max_nan(g, s_mx[r]);
      s_m = g;
    }
    __syncthreads();
    m = s_m;
  }
  const long long ntiles = (n + kTile - 1) / kTile;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double s1 = 0.0, s2 = 0.0;
  // units of work: unit u covers batches [u * kNB, (u + 1) * kNB) of the flat batch sequence; batch g lives in tile
  // g / kPerTile at element offset ((g % kPerTile) * kB + l) * 32 * V + lane * V   (arrays are padded to whole tiles)
  const long long nunits = ntiles * (kPerTile / kNB);
  const long long ustride = (long long)gridDim.x * (kBlock / 32);
  long long unit = (long long)blockIdx.x * (kBlock / 32) + warp;
  int bi = 0;
  R cur[kB][V], nxt[kB][V];
  auto load_batch = [&](long long g, R (&dst)[kB][V]) {
    const long long off = (g / kPerTile) * kTile + (long long)(g % kPerTile) * (kB * 32 * V);
#pragma unroll
    for (int l = 0; l < kB; l++) vload_keep<R>(logw + off + (long long)(l * 32 + lane) * V, dst[l]);
  };
  if (unit < nunits) load_batch(unit * kNB, cur);
  uint64_t sq = 0;
  while (unit < nunits) {
    long long nunit = unit;
    int nbi = bi + 1;
    if (nbi == kNB) { nbi = 0; nunit += ustride; }
    if (nunit < nunits) load_batch(nunit * kNB + nbi, nxt);
    const long long g = unit * kNB + bi;
    const long long tile = g / kPerTile;
    const long long base = tile * kTile + (long long)(g % kPerTile) * (kB * 32 * V);
    // the quantised weights are kept (8 B per particle, 0 beyond n): the ancestors pass scans THEM instead of evaluating the
    // exponential a second time -- it was issue-bound with a third of its instructions in that exp (profiles/r2j)
    unsigned long long *qdst = qout + base + (long long)lane * V;
    if (base + kB * 32 * V <= n) {                   // warp-uniform: whole batch inside the array, no bounds selects
#pragma unroll
      for (int l = 0; l < kB; l++) {
        uint64_t qv[V];
#pragma unroll
        for (int k = 0; k < V; k++) {
          double e;
          exp_and_quantize((double)cur[l][k] - m, bits, e, qv[k], wt);
          s1 += e;
          s2 = fma(e, e, s2);
          sq += qv[k];
        }
#pragma unroll
        for (int k = 0; k < V; k += 2)
          *reinterpret_cast<ulonglong2 *>(qdst + (long long)l * 32 * V + k) = make_ulonglong2(qv[k], qv[k + 1]);
      }
    } else {
#pragma unroll
      for (int l = 0; l < kB; l++) {
        uint64_t qv[V];
#pragma unroll
        for (int k = 0; k < V; k++) {
          double e;
          const long long i = base + (long long)(l * 32 + lane) * V + k;
          exp_and_quantize(i < n ? (double)cur[l][k] - m : -INFINITY, bits, e, qv[k], wt);
          s1 += e;
          s2 = fma(e, e, s2);
          sq += qv[k];
        }
#pragma unroll
        for (int k = 0; k < V; k += 2)
          *reinterpret_cast<ulonglong2 *>(qdst + (long long)l * 32 * V + k) = make_ulonglong2(qv[k], qv[k + 1]);
      }
    }
    if (bi == kNB - 1) {                             // unit done: its integer CDF mass
      sq = warp_sum_u64(sq);
      if (lane == 0) {
        if (WHOLE) tile_q[tile] = (unsigned long long)sq;
        else atomicAdd(tile_q + tile, (unsigned long long)sq);
        atomicAdd(super_q + tile / kSuper, (unsigned long long)sq);
      }
      sq = 0;
    }
#pragma unroll
    for (int l = 0; l < kB; l++)
#pragma unroll
      for (int k = 0; k < V; k++) cur[l][k] = nxt[l][k];
    unit = nunit;
    bi = nbi;
  }
  // CTA partial sums (fixed order => deterministic for a fixed grid)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  }
  if (lane == 0) s_d2[warp] = make_double2(s1, s2);
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0;
    for (int k = 0; k < kBlock / 32; k++) { a += s_d2[k].x; b += s_d2[k].y; }
    partials[blockIdx.x] = make_double2(a, b);
    __threadfence();
    unsigned int t = atomicAdd(ticket, 1u);
    s_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  // last CTA: fold the partials and the super-tile totals
  double a = 0.0, b = 0.0;
  for (unsigned int i = threadIdx.x; i < gridDim.x; i += kBlock) {
    a += __ldcg(&partials[i].x);
    b += __ldcg(&partials[i].y);
  }
  const long long nsuper = (ntiles + kSuper - 1) / kSuper;
  uint64_t qs = 0;
  for (long long i = threadIdx.x; i < nsuper; i += kBlock) qs += __ldcg(super_q + i);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
  __syncthreads();
  if (lane == 0) s_d2[warp] = make_double2(a, b);
  const uint64_t qtot = block_sum_u64(qs, s_u64);
  if (threadIdx.x == 0) {
    double t1 = 0.0, t2 = 0.0;
    for (int k = 0; k < kBlock / 32; k++) { t1 += s_d2[k].x; t2 += s_d2[k].y; }
    stats->m = m;
    stats->s1 = t1;
    stats->s2 = t2;
    // inference.jl:5: max == -Inf ? -Inf : max + log(sum(exp.(arr .- max)))
    stats->lse = (m == -INFINITY) ? -INFINITY : m + log(t1);
    // particle_filter.jl:3-6: exp(-logsumexp(2 lnw)) == s1^2 / s2 ; all -Inf => NaN there too
    stats->ess = (t1 * t1) / t2;
    stats->q_total = qtot;
    if (xc) {    // phase B of the exchange: this shard's (sum e, sum e^2, integer CDF mass) into every mailbox
      const unsigned long long seq = xc->seq_b + 1;
      xc->seq_b = seq;
      const int slot = (int)(seq & 1), me = xc->rank;
      for (int r = 0; r < xc->world; r++) {
        XchgMail *pm = xc->peer[r];
        pm->s1[slot][me] = t1; pm->s2[slot][me] = t2; pm->q[slot][me] = qtot;
      }
      __threadfence_system();
      for (int r = 0; r < xc->world; r++) st_relaxed_sys(&xc->peer[r]->rec_seq[me], seq);
    }
    if (ctl) {   // maybe_resample! decision + log-ML update on the device (particle_filter.jl:256, :263)
      const int flag = stats->ess < ctl->ess_threshold ? 1 : 0;
      ctl->do_resample = flag;
      if (flag) {
        ctl->log_ml_est += stats->lse - ctl->log_n;
        ctl->n_resampled += 1;
      }
    }
    *ticket = 0;
  }
  if (fp.sh) {     // sharded device-controlled run: wait for the peers' records, ESS test, exchange plan, super-tile scan
    __syncthreads();
    shard_plan_block(stats, fp.offs, fp.n_global, fp.u0, NAN, 1, fp.spill_cap, fp.sh_ctl, fp.plan, fp.sh, nsuper, super_q, fp.super_excl,
                     fp.overflow_count);
  } else if (fp.plan) {   // fused multi-step run: the super-tile scan + sweep plan ride on this CTA (one launch less per step)
    __syncthreads();
    superscan_block(nsuper, super_q, fp.super_excl, fp.plan, 0, 0, fp.n_global, fp.u0, fp.overflow_count, 0, s_u64_33);
  }
}

// ------------------------------------------------------------------------------------------
// cfg 4: Bayesian linear regression importance sampler (importance.jl:20-36 over
// w ~ Map(normal)(0,1), y_j ~ normal(x_j . w, sigma); map/generate.jl:10-37).  One thread per sample.
// The thread never holds w: it keeps the 64 running means m_j = sum_d X[j][d] w[d] in registers and, in a
// ROLLED loop over pairs of latents, draws (w_2k, w_2k+1) (one Philox block -> one Box-Muller pair), streams
// them to the trace columns and folds them into all 64 means (128 independent FMAs per pair: no dependency
// stalls; the loop body is ~600 instructions, so it lives in the instruction cache -- the fully unrolled
// first version was 117 KB of straight-line code at 2 warps per scheduler and ran at half this speed).
// X sits in shared memory in the filter's precision, pair-major, and is read by warp-uniform (broadcast)
// 128-bit loads.  No tensor cores (north_star): CUDA-core FMA at the fp64 rate is the bound.
// ------------------------------------------------------------------------------------------
constexpr int kBlrMax = 64;
__constant__ double c_blr_y[kBlrMax];

struct BlrArgs {
  void *state_out;   // D columns of stride ld
  void *logw;
  const double *normals;  // PREDRAWN [D][ld]
  const double *Xt;       // DEVICE [32 pairs][64 rows][2]: X[j][2k], X[j][2k+1], zero padded
  long long n, ld, pid_offset;
  unsigned long long seed;
  int D, nobs;
  double sigma;
  double *partials;
  unsigned int *ticket;
  WeightStats *stats;
};

template <typename R> struct Pair2;
template <> struct Pair2<double> { typedef double2 type; };
template <> struct Pair2<float> { typedef float2 type; };

#ifndef GB_BLR_MINB
#define GB_BLR_MINB 3
#endif
template <typename R, bool PREDRAWN>
__global__ void __launch_bounds__(128, GB_BLR_MINB) blr_generate_kernel(const BlrArgs a) {
  typedef typename Pair2<R>::type R2;
  __shared__ double s_tab[kMathTabDoubles];
  __shared__ __align__(16) R s_X[kBlrMax * kBlrMax];
  const MathTables T = stage_math_tables(s_tab);
  for (int t = threadIdx.x; t < kBlrMax * kBlrMax; t += blockDim.x) s_X[t] = (R)a.Xt[t];
  __syncthreads();
  R *__restrict__ sout = static_cast<R *>(a.state_out);
  R *__restrict__ logw = static_cast<R *>(a.logw);
  const R inv_sigma = R(1) / (R)a.sigma, log_sigma = Math<R>::log_((R)a.sigma);
  const int npairs = (a.D + 1) / 2;
  double acc = -INFINITY;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += (long long)gridDim.x * blockDim.x) {
    const unsigned long long pid = (unsigned long long)(a.pid_offset + i);
    R m[kBlrMax];
#pragma unroll
    for (int j = 0; j < kBlrMax; j++) m[j] = R(0);
#pragma unroll 1
    for (int k = 0; k < npairs; k++) {
      const bool two = 2 * k + 1 < a.D;
      R e0, e1;
      if (PREDRAWN) {
        e0 = (R)__ldcs(a.normals + (long long)(2 * k) * a.ld + i);
        e1 = two ? (R)__ldcs(a.normals + (long long)(2 * k + 1) * a.ld + i) : R(0);
      } else {
        philox_normal_pair(a.seed, pid, 0u, (unsigned)k, T, e0, e1);
        if (!two) e1 = R(0);
      }
      // normal(0, 1): mu + std * eps = 0 + 1 * eps (normal.jl:128) is exactly eps
      st_stream(sout + (long long)(2 * k) * a.ld + i, e0);
      if (two) st_stream(sout + (long long)(2 * k + 1) * a.ld + i, e1);
      const R2 *__restrict__ xr = reinterpret_cast<const R2 *>(s_X + k * (2 * kBlrMax));
#pragma unroll
      for (int j = 0; j < kBlrMax; j++) {
        const R2 x = xr[j];
        m[j] = fma(x.x, e0, m[j]);
        m[j] = fma(x.y, e1, m[j]);
      }
    }
    R lw = R(0);
#pragma unroll
    for (int j = 0; j < kBlrMax; j++)
      if (j < a.nobs) lw += logpdf_normal_fast<R>((R)c_blr_y[j], m[j], inv_sigma, log_sigma);   // map/generate.jl:16-17, in order
    logw[i] = lw;
    acc = max_nan(acc, (double)lw);
  }
  publish_max(acc, a.partials, a.ticket, a.stats);
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
