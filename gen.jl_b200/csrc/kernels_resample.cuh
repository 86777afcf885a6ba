// kernels_resample.cuh -- K4/K5: the resampling half of maybe_resample! (particle_filter.jl:249-275).
//
// The reference draws N multinomial ancestors with Distributions.jl's alias sampler
// (particle_filter.jl:261-262).  Here all three resamplers (multinomial, systematic, residual) are
// sweeps over an exact INTEGER CDF of fixed-point weights q_i = floor(exp(lw_i - max) * 2^b)
// (SURVEY.md Appendix B), so ancestors are a pure function of (log weights, uniforms): independent of
// the tiling, of the launch geometry and of how particles are sharded over GPUs, and bit-identical to
// the sequential oracle.
//
// RESIDUAL == SYSTEMATIC.  Appendix B defines residual resampling as c_j = floor(N q_j / Q) deterministic copies plus
// a systematic sweep, with the same offset u0, over the residual masses r_j = N q_j - c_j Q.  In exact integers the two
// coincide: residual slot k lies before the end of particle j iff floor((k 2^32 + u0) Q / 2^32) < N C_j - Q Cc_j
// (C = inclusive CDF of q, Cc = running count of copies), i.e. iff ((k + Cc_j) 2^32 + u0) Q < 2^32 N C_j, i.e. iff the
// SYSTEMATIC threshold tau_{k + Cc_j} < C_j; and every slot k' < Cc_j satisfies tau_k' < C_j because Cc_j Q <= N C_j.
// So the offspring range of particle j ends at the same slot under both definitions (the "residual systematic
// resampling" identity of Bolic et al., here exact).  GB_RESAMPLE_RESIDUAL therefore runs the systematic pipeline; the
// oracle keeps the definitional two-stage restatement and the parity tests compare against THAT
// (tests/test_gpu_parity.py::test_resamplers_bit_exact, tests/test_oracle_golden.py::test_residual_equals_systematic).
//
// Passes (all HBM-streaming, tiles of kTile = 2048 particles, 256 threads x 8 items): see "Systematic resampling,
// production pipeline" below; the CDF materialisation + guided search right below serve the multinomial draws.
#pragma once
#include "kernels_step.cuh"

namespace gb {

// ---- 64-bit block primitives built on warp shuffles ----
GB_D uint64_t shfl_u64(uint64_t v, int src) {
  uint32_t lo = __shfl_sync(0xffffffffu, (uint32_t)v, src);
  uint32_t hi = __shfl_sync(0xffffffffu, (uint32_t)(v >> 32), src);
  return ((uint64_t)hi << 32) | lo;
}
// exclusive block scan of one u64 per thread (the CDF never exceeds 2^62); *total = block sum
GB_D uint64_t block_exclusive_scan(uint64_t v, uint64_t *s_warp /* kBlock/32 + 1 */, uint64_t *total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint64_t inc = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    uint64_t o = shfl_up_u64(inc, d);
    if (lane >= d) inc += o;
  }
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    uint64_t w = lane < kBlock / 32 ? s_warp[lane] : 0;
    uint64_t winc = w;
#pragma unroll
    for (int d = 1; d < kBlock / 32; d <<= 1) {
      uint64_t o = shfl_up_u64(winc, d);
      if (lane >= d) winc += o;
    }
    if (lane < kBlock / 32) s_warp[lane] = winc - w;
    if (lane == kBlock / 32 - 1) s_warp[kBlock / 32] = winc;
  }
  __syncthreads();
  *total = s_warp[kBlock / 32];
  return s_warp[warp] + (inc - v);
}

// the same with ONE barrier: every thread folds the kBlock/32 warp totals itself (broadcast shared-memory reads).  The
// caller must not reuse s_warp before its next barrier.
GB_D uint64_t block_exclusive_scan_1b(uint64_t v, uint64_t *s_warp /* kBlock/32 */, uint64_t *total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint64_t inc = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    uint64_t o = shfl_up_u64(inc, d);
    if (lane >= d) inc += o;
  }
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  uint64_t base = 0, tot = 0;
#pragma unroll
  for (int k = 0; k < kBlock / 32; k++) {
    const uint64_t w = s_warp[k];
    base += k < warp ? w : 0;
    tot += w;
  }
  *total = tot;
  return base + (inc - v);
}

// ---- inclusive integer CDF, materialised (multinomial draws, sample_unweighted_traces) ----
// tilescan_cdf : one CTA of 1024 threads turns the per-tile totals (wstats) into exclusive tile prefixes (ntiles + 1 entries)
// cdf_kernel   : one CTA per tile: block scan of the quantised weights on top of the tile prefix -> cdf_out[i]
__global__ void __launch_bounds__(1024) tilescan_cdf_kernel(long long ntiles, uint64_t *tile_q, ResamplePlan *plan, long long n_global) {
  __shared__ uint64_t s_w[33];
  const int T = 1024;
  const long long per = (ntiles + T - 1) / T;
  const long long lo = (long long)threadIdx.x * per, hi = lo + per < ntiles ? lo + per : ntiles;
  uint64_t a = 0;
  for (long long i = lo; i < hi; i++) a += tile_q[i];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint64_t inc = a;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    uint64_t o = shfl_up_u64(inc, d);
    if (lane >= d) inc += o;
  }
  if (lane == 31) s_w[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    uint64_t w = s_w[lane], winc = w;
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
    const uint64_t v = tile_q[i];
    tile_q[i] = run;
    run += v;
  }
  if (threadIdx.x == 0) {
    tile_q[ntiles] = total;
    SweepParams sp;
    sp.q_total = total;
    sp.q_inv = 1.0 / (double)total;
    sp.dscale = (uint64_t)n_global << 32;
    sp.ratio = (double)n_global / (double)total;
    sp.u0 = 0;
    plan->sp = sp;
    plan->q_offset = 0;
    plan->n_global = n_global;
    plan->slot_lo = 0;
    plan->slot_hi = n_global;
  }
}
__global__ void __launch_bounds__(kBlock) cdf_kernel(const unsigned long long *__restrict__ qbuf, const uint64_t *__restrict__ tile_q,
                                                     uint64_t *__restrict__ cdf_out) {
  __shared__ uint64_t s_warp64[kBlock / 32 + 1];
  const long long tile = blockIdx.x;
  const long long i0 = tile * kTile + (long long)threadIdx.x * kItems;
  uint64_t c[kItems];
  uint64_t t = 0;
#pragma unroll
  for (int j = 0; j < kItems; j++) {
    t += __ldcs(qbuf + i0 + j);
    c[j] = t;
  }
  uint64_t total;
  const uint64_t base = tile_q[tile] + block_exclusive_scan(t, s_warp64, &total);
#pragma unroll
  for (int j = 0; j < kItems; j++) cdf_out[i0 + j] = base + c[j];
}

// ---- multinomial (reference semantics: N iid categorical draws in slot order) ----
// tau_k = floor(U_k * Q_N / 2^64); ancestor = first particle whose inclusive CDF exceeds tau_k.
// Two-level search: tile table (L2 resident), then the tile's 2048 CDF entries.
__global__ void __launch_bounds__(kBlock) multinomial_kernel(const uint64_t *__restrict__ cdf, const uint64_t *__restrict__ tile_q,
                                                             long long ntiles, long long n, const ResamplePlan *__restrict__ plan,
                                                             const uint64_t *__restrict__ u64s, unsigned long long seed,
                                                             unsigned int step, long long nslots, int *__restrict__ anc,
                                                             unsigned int stream_kind) {
  const uint64_t qn = plan->sp.q_total;
  for (long long k = (long long)blockIdx.x * kBlock + threadIdx.x; k < nslots; k += (long long)gridDim.x * kBlock) {
    uint64_t u;
    if (u64s) u = u64s[k];
    else {
      U4 w = philox_block(seed, (uint64_t)k, step, stream_kind, 0);
      u = ((uint64_t)w.x << 32) | w.y;
    }
    const uint64_t tau = mulhi64(u, qn);
    // last tile whose exclusive prefix is <= tau
    long long l = 0, h = ntiles;
    while (h - l > 1) {
      long long mid = (l + h) >> 1;
      if (__ldg(tile_q + mid) <= tau) l = mid; else h = mid;
    }
    long long a = l * kTile, b = a + kTile - 1;
    if (b > n - 1) b = n - 1;
    while (a < b) {
      long long mid = (a + b) >> 1;
      if (__ldg(cdf + mid) > tau) b = mid; else a = mid + 1;
    }
    anc[k] = (int)a;
  }
}

// ---- K5: resampled trace gather (particle_filter.jl:264-267) ----
// Multinomial draws through a guide table.  guide[j] = smallest i with cdf[i] > floor(j Q / N) -- exactly the systematic
// ancestors for u0 = 0, which the sweep above produces at ~0.5 ms per 1e8.  A draw tau in [0, Q) lies in guide cell
// J = floor(tau N / Q); its ancestor (smallest i with cdf[i] > tau) is bracketed by guide[J] and guide[J + 1], so the
// search touches two neighbouring guide entries and a handful of CDF values instead of 27 dependent loads spread over
// an 800 MB array.  J comes from a double-precision product (off by at most one either way), hence the bracket
// [guide[j - 1 .. j + 3]].  Same result as multinomial_kernel for every tau (the bracket always contains the answer).
__global__ void __launch_bounds__(kBlock) multinomial_guided_kernel(const uint64_t *__restrict__ cdf, const int *__restrict__ guide,
                                                                    long long n, const ResamplePlan *__restrict__ plan,
                                                                    const uint64_t *__restrict__ u64s, unsigned long long seed,
                                                                    unsigned int step, long long nslots, int *__restrict__ anc,
                                                                    unsigned int stream_kind) {
  const uint64_t qn = plan->sp.q_total;
  const double ratio = (double)n / (double)qn;
  for (long long k = (long long)blockIdx.x * kBlock + threadIdx.x; k < nslots; k += (long long)gridDim.x * kBlock) {
    uint64_t u;
    if (u64s) u = u64s[k];
    else {
      U4 w = philox_block(seed, (uint64_t)k, step, stream_kind, 0);
      u = ((uint64_t)w.x << 32) | w.y;
    }
    const uint64_t tau = mulhi64(u, qn);
    long long j = (long long)((double)tau * ratio) - 1;
    j = j < 0 ? 0 : (j > n - 1 ? n - 1 : j);
    long long a = __ldg(guide + j);
    long long b = j + 3 < n ? (long long)__ldg(guide + j + 3) : n - 1;
    // the bracket [a, b] holds ~4 particles: its CDF entries are fetched with INDEPENDENT loads (one or two sectors in
    // flight at once) and counted, instead of a chain of three dependent probes -- the kernel is bound by the latency of
    // its random accesses, not by their bytes.  The answer is the first index in [a, b] whose CDF exceeds tau (b if none
    // before it does: the bracket always contains it), i.e. a + #{i in [a, b) : cdf[i] <= tau} because the CDF is monotone.
    if (b - a <= 8) {
      uint64_t c[8];
#pragma unroll
      for (int e = 0; e < 8; e++) c[e] = a + e < b ? __ldg(cdf + a + e) : ~0ull;
      int below = 0;
#pragma unroll
      for (int e = 0; e < 8; e++) below += c[e] <= tau ? 1 : 0;
      a += below;
    } else {
      while (a < b) {
        const long long mid = (a + b) >> 1;
        if (__ldg(cdf + mid) > tau) b = mid; else a = mid + 1;
      }
    }
    anc[k] = (int)a;
  }
}

template <typename R>
__global__ void __launch_bounds__(kBlock) gather_kernel(const R *__restrict__ sin, R *__restrict__ sout, long long ld, int S,
                                                        const int *__restrict__ anc, long long n, R *__restrict__ logw) {
  constexpr int V = VecOf<R>::V;
  const long long nvec = (n + V - 1) / V;
  for (long long iv = (long long)blockIdx.x * kBlock + threadIdx.x; iv < nvec; iv += (long long)gridDim.x * kBlock) {
    const long long i0 = iv * V;
    int src[V];
#pragma unroll
    for (int v = 0; v < V; v++) src[v] = (i0 + v < n) ? __ldcs(anc + i0 + v) : 0;
    for (int s = 0; s < S; s++) {
      R val[V];
#pragma unroll
      for (int v = 0; v < V; v++) val[v] = __ldg(sin + (long long)s * ld + src[v]);
      vstore<R>(sout + (long long)s * ld + i0, val);
    }
    if (logw) {
      R z[V];
#pragma unroll
      for (int v = 0; v < V; v++) z[v] = R(0);
      vstore_keep<R>(logw + i0, z);
    }
  }
}

// Exchange block of `count` offspring rows: [parents: count x int64 (global, 1-based)][rows: S x count x R]
template <typename R> GB_HD long long *block_parents(void *blk) { return reinterpret_cast<long long *>(blk); }
template <typename R> GB_HD R *block_rows(void *blk, long long count) { return reinterpret_cast<R *>(reinterpret_cast<long long *>(blk) + count); }

// gather for the sharded path: rows of `count` consecutive offspring slots (ancestors anc[0..count)) into a block
template <typename R>
__global__ void __launch_bounds__(kBlock) export_kernel(const R *__restrict__ sin, long long ld, int S, const int *__restrict__ anc,
                                                        long long count, void *blk, long long pid_offset) {
  long long *parents = block_parents<R>(blk);
  R *rows = block_rows<R>(blk, count);
  for (long long i = (long long)blockIdx.x * kBlock + threadIdx.x; i < count; i += (long long)gridDim.x * kBlock) {
    const int src = __ldcs(anc + i);
    for (int s = 0; s < S; s++) rows[(long long)s * count + i] = __ldg(sin + (long long)s * ld + src);
    parents[i] = pid_offset + src + 1;   // Gen's parents are 1-based
  }
}
template <typename R>
__global__ void __launch_bounds__(kBlock) import_kernel(R *__restrict__ sout, long long ld, int S, long long dst_lo, long long count,
                                                        const void *blk, long long *__restrict__ parents) {
  const long long *parents_in = block_parents<R>(const_cast<void *>(blk));
  const R *rows = block_rows<R>(const_cast<void *>(blk), count);
  for (long long i = (long long)blockIdx.x * kBlock + threadIdx.x; i < count; i += (long long)gridDim.x * kBlock) {
    for (int s = 0; s < S; s++) sout[(long long)s * ld + dst_lo + i] = rows[(long long)s * count + i];
    parents[dst_lo + i] = parents_in[i];
  }
}

// rows of the sampled particles: out[s][i] = state[s][idx[i]] as fp64; idx1[i] = global 1-based index
template <typename R>
__global__ void __launch_bounds__(kBlock) sample_rows_kernel(const R *__restrict__ sin, long long ld, int S, const int *__restrict__ idx,
                                                             long long count, double *__restrict__ rows, long long *__restrict__ idx1,
                                                             long long pid_offset) {
  for (long long i = (long long)blockIdx.x * kBlock + threadIdx.x; i < count; i += (long long)gridDim.x * kBlock) {
    const int src = idx[i];
    for (int s = 0; s < S; s++) rows[(long long)s * count + i] = (double)sin[(long long)s * ld + src];
    idx1[i] = pid_offset + src + 1;
  }
}

// trajectory reconstruction (get_traces with full history): out[i] = snapshot[field][idx[i]]; when a resample
// preceded this step, the lineage then moves to the ancestor: idx[i] = anc[idx[i]]
template <typename R>
__global__ void __launch_bounds__(kBlock) backtrace_kernel(const R *__restrict__ snap, const int *__restrict__ anc, int *__restrict__ idx,
                                                           long long n, double *__restrict__ out) {
  for (long long i = (long long)blockIdx.x * kBlock + threadIdx.x; i < n; i += (long long)gridDim.x * kBlock) {
    const int j = idx[i];
    out[i] = (double)snap[j];
    if (anc) idx[i] = anc[j];
  }
}
__global__ void __launch_bounds__(kBlock) iota_int_kernel(int *__restrict__ p, long long n) {
  for (long long i = (long long)blockIdx.x * kBlock + threadIdx.x; i < n; i += (long long)gridDim.x * kBlock) p[i] = (int)i;
}

template <typename R>
__global__ void __launch_bounds__(kBlock) fill_kernel(R *__restrict__ p, long long n, R v) {
  for (long long i = (long long)blockIdx.x * kBlock + threadIdx.x; i < n; i += (long long)gridDim.x * kBlock) p[i] = v;
}
__global__ void __launch_bounds__(kBlock) iota_parents_kernel(long long *__restrict__ p, long long n, long long base) {
  for (long long i = (long long)blockIdx.x * kBlock + threadIdx.x; i < n; i += (long long)gridDim.x * kBlock) p[i] = base + i + 1;
}
__global__ void __launch_bounds__(kBlock) parents_from_anc_kernel(const int *__restrict__ anc, long long *__restrict__ p, long long n,
                                                                  long long base, long long ext_begin,
                                                                  const long long *__restrict__ ext_parents) {
  for (long long i = (long long)blockIdx.x * kBlock + threadIdx.x; i < n; i += (long long)gridDim.x * kBlock) {
    const long long a = anc[i];
    p[i] = a >= ext_begin ? ext_parents[a - ext_begin] : base + a + 1;   // rows that migrated in carry their global parent
  }
}
// rows received from another shard -> extension rows [ext_row, ext_row + count) of the current trace
// buffer; the destination slots' ancestors point at them
template <typename R>
__global__ void __launch_bounds__(kBlock) import_ext_kernel(R *__restrict__ scur, long long ld, int S, long long dst_lo, long long count,
                                                            long long ext_row, const void *blk, int *__restrict__ anc,
                                                            long long *__restrict__ ext_parents) {
  const long long *parents_in = block_parents<R>(const_cast<void *>(blk));
  const R *rows = block_rows<R>(const_cast<void *>(blk), count);
  for (long long i = (long long)blockIdx.x * kBlock + threadIdx.x; i < count; i += (long long)gridDim.x * kBlock) {
    for (int s = 0; s < S; s++) scur[(long long)s * ld + ext_row + i] = rows[(long long)s * count + i];
    anc[dst_lo + i] = (int)(ext_row + i);
    ext_parents[i] = parents_in[i];
  }
}
template <typename R>
__global__ void __launch_bounds__(kBlock) convert_out_kernel(const R *__restrict__ in, double *__restrict__ out, long long n) {
  for (long long i = (long long)blockIdx.x * kBlock + threadIdx.x; i < n; i += (long long)gridDim.x * kBlock) out[i] = (double)in[i];
}
template <typename R>
__global__ void __launch_bounds__(kBlock) convert_in_kernel(const double *__restrict__ in, R *__restrict__ out, long long n) {
  for (long long i = (long long)blockIdx.x * kBlock + threadIdx.x; i < n; i += (long long)gridDim.x * kBlock) out[i] = (R)in[i];
}

// ==========================================================================================
// Systematic resampling, production pipeline (v1):
//   wstats (kernels_step.cuh)  per-tile + per-super-tile integer CDF mass, logsumexp/ESS sums
//   superscan                  <= 16384 super-tile totals: exclusive prefix + first slot (1 CTA, tiny)
//   ancestors_sys              one CTA per 2048-particle tile: tile prefix from the super prefix + <= 63
//                              tile totals, block scan of the quantised weights, offspring range per
//                              particle from a double-precision estimate with exact 128/64-bit fallback,
//                              head flags + max-scan in shared memory -> coalesced int32 ancestors
//   overflow_sys               tiles whose particles own more than kChunk output slots queue themselves;
//                              this kernel spreads the remaining slots over the whole GPU (exits at once
//                              when the queue is empty, which is the normal case)
// ==========================================================================================
struct OverflowEntry { long long tile, slot0, slot_lo, slot_hi; unsigned long long cbase; };

__global__ void __launch_bounds__(1024) superscan_kernel(long long nsuper, unsigned long long *__restrict__ super_q,
                                                         unsigned long long *__restrict__ super_excl,
                                                         ResamplePlan *plan, uint64_t q_offset, uint64_t q_total_given,
                                                         long long n_global, uint32_t u0, unsigned int *overflow_count,
                                                         int plan_preset) {
  __shared__ uint64_t s_w[33];
  pdl_trigger();
  pdl_wait();
  superscan_block(nsuper, super_q, super_excl, plan, q_offset, q_total_given, n_global, u0, overflow_count, plan_preset, s_w);
}

// Where the ancestors of output slot k go: slots this shard owns -> own[k - own_lo] (the shard's `anc`
// column, consumed by the deferred gather); slots owned by other shards -> spill[k - spill_base]
struct AncOut {
  int *own;
  long long own_lo, own_hi;
  int *spill;
  long long spill_base;     // >= 0: host-known first slot of the spill buffer; < 0: plan->slot_lo (device-side plan)
  long long spill_cap;      // entries in the spill buffer (writes beyond it are dropped: the plan flags the overflow)
};

constexpr int kBlk = 8;                             // output slots per fill block: two 128-bit stores per thread
constexpr int kEndSlots = kChunk + 2 * kBlk;        // histogram entries of one window (+ alignment slack)
constexpr int kEndWords = ((kEndSlots / 2 + kEndSlots / 8) + 3) / 4 * 4;   // two 16-bit entries per word + one padding word per four

// number of slots below CDF value C: double-precision estimate, exact only when the estimate is within
// 2^-17 slots of a boundary; `slow` is set instead of branching so callers can batch the rare fix-ups
GB_D int slots_estimate(uint64_t C, double ratio, double u0term, bool &slow) {
  const double y = fma((double)C, ratio, u0term);
  const double cc = ceil(y);
  const bool neg = y < -0x1p-17;
  // cc - y in (2^-17, 1 - 2^-17)  <=>  |cc - y - 1/2| < 1/2 - 2^-17
  slow = !(neg || fabs((cc - y) - 0.5) < 0.5 - 0x1p-17);
  return neg ? 0 : __double2int_rz(cc);
}

// Shared-memory scratch of one ancestors CTA
struct TileSmem {
  uint64_t warp64[kBlock / 32 + 1];
  // Histogram of the window: the number of particles whose offspring range ends at window slot x - sh, as a 16-bit count
  // (a tile has 2048 particles) -- two slots per word, so the window of kChunk slots costs 20 KB instead of 37 KB of shared
  // memory and a fifth CTA fits on the SM (the kernel is bound by the latency of its load -> scan -> histogram -> scan ->
  // store chain, i.e. by the number of tiles in flight).  One padding word per 4 words: consecutive lanes add at positions
  // ~8 slots = 4 words apart, which would be an 8-way bank conflict in a dense array; with the padding they -- and the
  // block-wise reads, stride 5 words -- spread over all 32 banks.
  __align__(16) unsigned int ends[kEndWords];
  int warp32[kBlock / 32 + 1];
  uint64_t cbase;
};
GB_D void ends_add(TileSmem &sm, int x) {
  const int w = x >> 1;
  atomicAdd(&sm.ends[w + (w >> 2)], 1u << ((x & 1) * 16));
}
GB_D void tile_smem_clear(TileSmem &sm) {   // the histogram must be all zero on entry of tile_fill
  uint4 *z = reinterpret_cast<uint4 *>(sm.ends);
  for (int i = threadIdx.x; i < kEndWords / 4; i += kBlock) z[i] = make_uint4(0u, 0u, 0u, 0u);
  __syncthreads();
}
// the 8 counts of fill block B (read and cleared)
GB_D void ends_take(TileSmem &sm, int B, int (&v)[kBlk]) {
  unsigned int *p = sm.ends + 5 * B;
#pragma unroll
  for (int e = 0; e < kBlk / 2; e++) {
    const unsigned int w = p[e];
    p[e] = 0u;
    v[2 * e] = (int)(w & 0xffffu);
    v[2 * e + 1] = (int)(w >> 16);
  }
}
// their sum (not cleared)
GB_D int ends_sum(const TileSmem &sm, int B) {
  const unsigned int *p = sm.ends + 5 * B;
  int t = 0;
#pragma unroll
  for (int e = 0; e < kBlk / 2; e++) t += (int)(p[e] & 0xffffu) + (int)(p[e] >> 16);
  return t;
}
// exclusive block scan of one int per thread
GB_D int block_exclusive_scan_i32(int v, int *s_warp /* kBlock/32 + 1 */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int o = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc += o;
  }
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    const int w = lane < kBlock / 32 ? s_warp[lane] : 0;
    int winc = w;
#pragma unroll
    for (int d = 1; d < kBlock / 32; d <<= 1) {
      const int o = __shfl_up_sync(0xffffffffu, winc, d);
      if (lane >= d) winc += o;
    }
    if (lane < kBlock / 32) s_warp[lane] = winc - w;
  }
  __syncthreads();
  return s_warp[warp] + (inc - v);
}

GB_D int block_exclusive_scan_i32_1b(int v, int *s_warp /* kBlock/32 */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int o = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc += o;
  }
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  int base = 0;
#pragma unroll
  for (int k = 0; k < kBlock / 32; k++) base += k < warp ? s_warp[k] : 0;
  return base + (inc - v);
}

// Scan one tile and fill the ancestors of the output slots [lo, hi) (hi - lo <= kChunk) that belong to it.
//   PREFIX = true : regular CTA.  cbase (integer CDF mass before the tile) is produced by warp 0 from the
//                   super-tile prefix while the other warps already quantise; [lo, hi) = the tile's first
//                   kChunk slots; returns the tile's slot range through slot0_out / slot_end_out.
//   PREFIX = false: overflow CTA.  cbase / slot0 / [lo, hi) are given.
// Fill (r2): the ancestor of window slot i is the number of particles of the tile whose offspring range ends at or before
// i.  So every particle adds ONE count at its end slot (a shared-memory reduction, whatever its number of offspring), a
// block-wide prefix sum over the window turns the counts into ancestors, and every thread writes its 8 consecutive
// slots -- aligned to the destination array -- with two 128-bit stores.  Work is proportional to particles + slots and
// every memory operation is independent of the others.  (The r1 fill let every particle write its own slots into a
// shared-memory window: the per-lane loops over 3+ offspring ran in lock step to the LONGEST range of the warp and were
// 24% of the kernel's instructions, profiles/r2f; a walk over sorted end slots per 8-slot block removed the loops but
// serialised on shared-memory latency and was slower still, 0.66 vs 0.48 ms.)  sm.ends is all zero on entry and on exit.
template <bool PREFIX>
GB_D void tile_fill(const unsigned long long *__restrict__ qbuf, const SweepParams &sp, long long tile,
                    const uint64_t *__restrict__ tile_q, const unsigned long long *__restrict__ super_excl, uint64_t q_offset,
                    uint64_t cbase_in, long long slot0_in, long long lo_in, long long hi_in, const AncOut &ao,
                    long long spill_base, TileSmem &sm, long long *slot0_out, long long *slot_end_out,
                    uint64_t *total_out = nullptr) {
  // the quantised weights of the tile as the statistics pass left them (q = 0 beyond n): no second exp per particle
  uint64_t qv[kItems];
  {
    const ulonglong2 *src = reinterpret_cast<const ulonglong2 *>(qbuf + tile * kTile + (long long)threadIdx.x * kItems);
#pragma unroll
    for (int j = 0; j < kItems; j += 2) {
      const ulonglong2 t2 = __ldcs(src + j / 2);
      qv[j] = t2.x; qv[j + 1] = t2.y;
    }
  }
  const double u0term = -(double)sp.u0 * 0x1p-32;
  if (PREFIX && threadIdx.x < 32) {
    if (tile_q) {
      // CDF mass before this tile: super-tile prefix + the (<= 63) tile totals before it inside the super-tile
      const long long sbase = (tile / kSuper) * kSuper;
      uint64_t v = 0;
      for (long long i = sbase + threadIdx.x; i < tile; i += 32) v += tile_q[i];
      v = warp_sum_u64(v);
      if (threadIdx.x == 0) sm.cbase = q_offset + super_excl[tile / kSuper] + v;
    } else if (threadIdx.x == 0) sm.cbase = cbase_in;     // (persistent run: the caller carries the prefix itself)
  }
  uint64_t c[kItems];
  uint64_t t = 0;
#pragma unroll
  for (int j = 0; j < kItems; j++) {
    t += qv[j];
    c[j] = t;
  }
  uint64_t total;
  const uint64_t ex = block_exclusive_scan_1b(t, sm.warp64, &total);   // one barrier: sm.cbase is visible after it
  if (total_out) *total_out = total;
  const uint64_t cbase = PREFIX ? sm.cbase : cbase_in;
  // offspring end of every particle; rare exact fix-ups batched after the unrolled loop
  int kend[kItems];
  unsigned slow = 0;
#pragma unroll
  for (int j = 0; j < kItems; j++) {
    bool sl;
    kend[j] = slots_estimate(cbase + ex + c[j], sp.ratio, u0term, sl);
    slow |= sl ? (1u << j) : 0u;
  }
  if (slow) {
#pragma unroll
    for (int j = 0; j < kItems; j++)
      if (slow & (1u << j)) kend[j] = (int)slots_below(cbase + ex + c[j], sp);
  }
  // the tile's slot range, by every thread itself (same inputs, same bits as the last particle's end / the first one's
  // start): saves a shared-memory exchange and its barrier
  long long slot0 = slot0_in, slot_end;
  {
    bool sl;
    int se = slots_estimate(cbase + total, sp.ratio, u0term, sl);
    if (sl) se = (int)slots_below(cbase + total, sp);
    slot_end = se;
    if (PREFIX) {
      int s0 = slots_estimate(cbase, sp.ratio, u0term, sl);
      if (sl) s0 = (int)slots_below(cbase, sp);
      slot0 = s0;
    }
  }
  if (PREFIX) { *slot0_out = slot0; *slot_end_out = slot_end; }
  const long long lo = PREFIX ? slot0 : lo_in;
  long long hi = PREFIX ? slot_end : hi_in;
  if (PREFIX && hi > lo + kChunk) hi = lo + kChunk;
  const int cnt = (int)(hi - lo);
  if (cnt <= 0) return;                       // uniform across the CTA
  // fill blocks are aligned to the destination: window slot i goes to own[d0 + i] and sits at histogram position
  // x = i + sh, sh = d0 mod 8; block B covers x in [8 B, 8 B + 8)
  const long long d0 = lo - ao.own_lo;
  const int sh = (int)(d0 & (kBlk - 1));
  const int ilo = (int)lo;
#pragma unroll
  for (int j = 0; j < kItems; j++) {
    int x = kend[j] - ilo;                    // the particle's range ends at window slot x: it precedes the owners of slots >= x
    x = x < 0 ? 0 : x;                        // (overflow pieces: ended before the window)
    if (x < cnt) ends_add(sm, x + sh);
  }
  __syncthreads();
  const int tile_base = (int)(tile * kTile);
  const int nblk = ((cnt - 1 + sh) >> 3) + 1;
  const int per = (nblk + kBlock - 1) / kBlock;            // consecutive blocks per thread: 1, unless the tile owns > 2048 slots
  const bool owned = lo >= ao.own_lo && hi <= ao.own_hi;   // the usual case (always, on a single shard): every slot is this shard's own
  int run;
  if (per == 1) {
    const int B = threadIdx.x;
    int v[kBlk] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (B < nblk) ends_take(sm, B, v);                      // (leaves the histogram clean)
#pragma unroll
    for (int e = 1; e < kBlk; e++) v[e] += v[e - 1];
    run = tile_base + block_exclusive_scan_i32_1b(v[kBlk - 1], sm.warp32);
    if (B < nblk) {
      const int s = (B << 3) - sh;                          // window slot of element 0 (< 0 only for block 0)
      if (owned && s >= 0 && s + kBlk <= cnt) {
        int4 *dst = reinterpret_cast<int4 *>(ao.own + (d0 + s));
        __stcs(dst, make_int4(run + v[0], run + v[1], run + v[2], run + v[3]));
        __stcs(dst + 1, make_int4(run + v[4], run + v[5], run + v[6], run + v[7]));
      } else {
#pragma unroll
        for (int e = 0; e < kBlk; e++) {
          const int i = s + e;
          if (i < 0 || i >= cnt) continue;
          const long long k = lo + i;
          if (k >= ao.own_lo && k < ao.own_hi) __stcs(ao.own + (k - ao.own_lo), run + v[e]);
          else if (k - spill_base >= 0 && k - spill_base < ao.spill_cap) __stcs(ao.spill + (k - spill_base), run + v[e]);
        }
      }
    }
    return;
  }
  // heavy tile (more than 2048 slots in the window): `per` consecutive blocks per thread, two passes over them
  const int B0 = threadIdx.x * per;
  int tot = 0;
  for (int q = 0; q < per; q++) {
    const int B = B0 + q;
    if (B >= nblk) break;
    tot += ends_sum(sm, B);
  }
  run = tile_base + block_exclusive_scan_i32(tot, sm.warp32);
  for (int q = 0; q < per; q++) {
    const int B = B0 + q;
    if (B >= nblk) break;
    int v[kBlk];
    ends_take(sm, B, v);
#pragma unroll
    for (int e = 1; e < kBlk; e++) v[e] += v[e - 1];
    const int s = (B << 3) - sh;
    if (owned && s >= 0 && s + kBlk <= cnt) {
      int4 *dst = reinterpret_cast<int4 *>(ao.own + (d0 + s));
      __stcs(dst, make_int4(run + v[0], run + v[1], run + v[2], run + v[3]));
      __stcs(dst + 1, make_int4(run + v[4], run + v[5], run + v[6], run + v[7]));
    } else {
#pragma unroll
      for (int e = 0; e < kBlk; e++) {
        const int i = s + e;
        if (i < 0 || i >= cnt) continue;
        const long long k = lo + i;
        if (k >= ao.own_lo && k < ao.own_hi) __stcs(ao.own + (k - ao.own_lo), run + v[e]);
        else if (k - spill_base >= 0 && k - spill_base < ao.spill_cap) __stcs(ao.spill + (k - spill_base), run + v[e]);
      }
    }
    run += v[kBlk - 1];
  }
}

#ifndef GB_ANC_MINB
#define GB_ANC_MINB 5
#endif
__global__ void __launch_bounds__(kBlock, GB_ANC_MINB) ancestors_sys_kernel(const unsigned long long *__restrict__ qbuf,
                                                               const uint64_t *__restrict__ tile_q,
                                                               const unsigned long long *__restrict__ super_excl,
                                                               const ResamplePlan *__restrict__ plan, const AncOut ao,
                                                               OverflowEntry *__restrict__ oflow,
                                                               unsigned int *__restrict__ oflow_count,
                                                               const RunCtl *__restrict__ ctl) {
  __shared__ TileSmem sm;
  pdl_trigger();
  tile_smem_clear(sm);             // (constant work: overlaps the predecessor's tail)
  pdl_wait();
  if (ctl && !ctl->do_resample) return;      // fused multi-step run: the decision lives on the device
  const SweepParams sp = plan->sp;
  const long long tile = blockIdx.x;
  long long slot0, slot_end;
  const long long spill_base = ao.spill_base >= 0 ? ao.spill_base : plan->slot_lo;
  tile_fill<true>(qbuf, sp, tile, tile_q, super_excl, plan->q_offset, 0, 0, 0, 0, ao, spill_base, sm, &slot0, &slot_end);
  if (slot_end > slot0 + kChunk && threadIdx.x == 0) {
    const unsigned int e = atomicAdd(oflow_count, 1u);
    oflow[e] = OverflowEntry{tile, slot0, slot0 + kChunk, slot_end, sm.cbase};
  }
}

// ==========================================================================================
// Sharded resample without a host round trip and without collectives (one shard per GPU):
//   * the global max / the per-shard sums travel through the peer-memory mailboxes of kernels_step.cuh (phases A, B);
//   * shard_plan_kernel turns them into the ESS test, the log-ML update and the exchange plan ON THE DEVICE (every shard
//     evaluates the same arithmetic on the same G records, so all agree without talking), then scans the super-tiles;
//   * offspring owned by another shard are written by the sweep's second kernel STRAIGHT INTO THAT SHARD'S MEMORY over NVLink
//     (peer pointers: same process or CUDA IPC): rows into the extension rows of its current trace buffer, the slot's
//     ancestor index and the global parent id next to them -- no staging buffer, no all-to-all, traffic = the spill;
//   * its last CTA posts a "done" flag to every peer (phase C); the next step kernel waits for all of them before it reads
//     the rows (xchg_wait_done at its top).
// ==========================================================================================
// Second kernel of the sweep.  (1) leaves the per-tile accumulator clean for the next statistics pass; (2) remaining slots
// of heavy tiles: work unit = (queue entry, kChunk-sized piece), dealt round-robin to the CTAs of the grid (normally the
// queue is empty); (3) sharded filters: the offspring rows of slots owned by OTHER shards go straight into the owners'
// memory (NVLink peer stores: rows into the extension rows of their current trace buffer, the slot's ancestor index and
// global parent id next to them), and the last CTA to finish posts this shard's "done" flag (phase C) to every peer.
// (3) used to be a kernel of its own: ~20 us of launch + fence latency per step, a twentieth of the 8-GPU strong-scaling step.
struct ExportArgs {
  const void *sin;                 // current traces (source rows); nullptr: not a sharded sweep
  long long ld;
  int S, cur;
  const int *spill;                // ancestors of the slots this shard produced for other shards
  const ShardPlanDev *sh;
  const long long *offs;
  long long pid_offset, spill_cap;
  WeightStats *stats;
  unsigned int *grid_barrier;      // monotone counter (only used when heavy tiles queued work: their slots may be exported)
};
template <typename R>
__global__ void __launch_bounds__(kBlock) overflow_sys_kernel(const unsigned long long *__restrict__ qbuf,
                                                              const ResamplePlan *__restrict__ plan, const AncOut ao,
                                                              const OverflowEntry *__restrict__ oflow,
                                                              const unsigned int *__restrict__ oflow_count,
                                                              unsigned long long *__restrict__ tile_q, long long ntiles,
                                                              const RunCtl *__restrict__ ctl, const ExportArgs ex) {
  __shared__ TileSmem sm;
  __shared__ bool s_last;
  pdl_trigger();
  tile_smem_clear(sm);
  pdl_wait();
  for (long long i = (long long)blockIdx.x * kBlock + threadIdx.x; i < ntiles; i += (long long)gridDim.x * kBlock) tile_q[i] = 0;
  if (ctl && !ctl->do_resample) return;        // uniform over the grid and over the shards: nobody posts, nobody waits
  const unsigned int ne = *oflow_count;
  if (ne) {
    const SweepParams sp = plan->sp;
    long long unit0 = 0;
    for (unsigned int e = 0; e < ne; e++) {
      const OverflowEntry en = oflow[e];
      const long long pieces = (en.slot_hi - en.slot_lo + kChunk - 1) / kChunk;
      // first piece of this entry handled by this CTA (units are dealt round-robin over the grid)
      long long first = ((long long)blockIdx.x - unit0 % gridDim.x + gridDim.x) % gridDim.x;
      for (long long pc = first; pc < pieces; pc += gridDim.x) {
        const long long lo = en.slot_lo + pc * kChunk;
        long long hi = lo + kChunk;
        if (hi > en.slot_hi) hi = en.slot_hi;
        __syncthreads();
        tile_fill<false>(qbuf, sp, en.tile, nullptr, nullptr, 0, en.cbase, en.slot0, lo, hi, ao,
                         ao.spill_base >= 0 ? ao.spill_base : plan->slot_lo, sm, nullptr, nullptr);
      }
      unit0 += pieces;
    }
  }
  if (!ex.sin) return;
  const ShardPlanDev *__restrict__ sh = ex.sh;
  XchgCtl *xc = ex.stats->xc;
  if (ne) {
    // heavy tiles wrote spill ancestors in THIS kernel: every CTA must be past that before any row is exported (the grid is
    // sized to be co-resident, so a counter barrier is safe; only taken in this rare case)
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      const unsigned int t = atomicAdd(ex.grid_barrier, 1u);
      const unsigned int target = (t / gridDim.x + 1u) * gridDim.x;
      while (*((volatile unsigned int *)ex.grid_barrier) < target) __nanosleep(100);
      __threadfence();
    }
    __syncthreads();
  }
  const R *__restrict__ sin = static_cast<const R *>(ex.sin);
  bool wrote = false;
  for (int d = 0; d < sh->world; d++) {
    if (d == sh->rank) continue;
    const long long cnt = sh->send_hi[d] - sh->send_lo[d];
    if (cnt <= 0) continue;
    const PeerMem pm = sh->peer[d];
    R *__restrict__ dst = static_cast<R *>(pm.state[ex.cur]);
    const long long base = sh->send_lo[d] - plan->slot_lo;       // index into this shard's spill buffer
    const long long dst_slot0 = sh->send_lo[d] - ex.offs[d];     // local slot at the destination
    const long long ext0 = pm.n_pad + sh->ext_off[d];
    for (long long i = (long long)blockIdx.x * kBlock + threadIdx.x; i < cnt; i += (long long)gridDim.x * kBlock) {
      const int src = (base + i < ex.spill_cap) ? ex.spill[base + i] : 0;
      for (int s = 0; s < ex.S; s++) dst[(long long)s * pm.ld + ext0 + i] = __ldg(sin + (long long)s * ex.ld + src);
      pm.anc[dst_slot0 + i] = (int)(ext0 + i);
      pm.ext_parents[sh->ext_off[d] + i] = ex.pid_offset + src + 1;   // Gen's parents are 1-based, global
      wrote = true;
    }
  }
  const int any = __syncthreads_or(wrote ? 1 : 0);
  if (threadIdx.x == 0) {
    // ONE system-scope fence per CTA, after the barrier: it is cumulative over the peer stores of every thread of the CTA
    // (they happen before it through the barrier), and orders them before the ticket and the flags below
    if (any) __threadfence_system();
    const unsigned int t = atomicAdd(&xc->export_ticket, 1u);
    s_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last || threadIdx.x != 0) return;
  __threadfence_system();
  xc->export_ticket = 0;
  const unsigned long long seq = xc->seq_c + 1;
  xc->seq_c = seq;
  for (int r = 0; r < xc->world; r++) st_relaxed_sys(&xc->peer[r]->done_seq[xc->rank], seq);
}
// phase C wait as a kernel of its own, for readers of the extension rows that are not step kernels (the eager gather)
__global__ void __launch_bounds__(kMaxWorld) xchg_wait_kernel(WeightStats *stats) {
  pdl_trigger();
  pdl_wait();
  xchg_wait_done(stats);
}

}  // namespace gb
