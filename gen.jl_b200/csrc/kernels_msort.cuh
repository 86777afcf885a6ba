// kernels_msort.cuh -- GB_RESAMPLE_MULTINOMIAL_SORTED: multinomial resampling as ONE streaming merge.
//
// The reference draws N iid categorical ancestors (particle_filter.jl:261-262); as N independent searches over the CDF
// that is bound by random 32-byte sectors (5.4 ms per 1e8 draws).  The order statistics of N iid
// uniforms can be generated directly, sorted, through exponential spacings -- U_(k) = (E_0 + ... + E_k) / (E_0 + ... + E_N)
// for iid exponential E_j (Devroye 1986, ch. V.3) -- and sorted thresholds turn the N searches into a merge of two sorted
// sequences, the same access pattern as the systematic sweep.  As a multiset the ancestors are distributed exactly like
// the reference's draws (up to the 2^-30 resolution of the spacings); they come out sorted by ancestor instead of in iid
// slot order, which a particle filter never looks at (traces are exchangeable after a resample) and which gives the gather
// fused into the next step its locality back.
//
// Exact integer spec (oracle: go_resample_multinomial_sorted).  Spacings e_j >= 1 (integers; Philox + the table-driven
// log, or pre-drawn by the caller), T_k = e_0 + ... + e_k, T = T_N, C = inclusive CDF of the quantised weights, Q = C_{N-1}:
//     a_k = min{ j : C_j * T > T_k * Q }              (128-bit products; T_k < T, so a_k exists)
//
// Passes (tiles of 2048, as everywhere):
//   msort_spacing    e_j (generated, or read) and the sum of every SLOT tile
//   msort_scan       one CTA: exclusive scan of the slot-tile sums (-> T) and of the PARTICLE-tile totals tile_q (-> Q)
//   msort_partition  one thread per slot-tile boundary X: the particle tile that contains threshold X (binary search over
//                    the 49 k tile prefixes, L2 resident)
//   msort_fill       one CTA per slot tile: T_k of its 2048 slots by a block scan, then for every particle tile in its
//                    range: block scan of the quantised weights into shared memory, binary search per slot
#pragma once
#include "kernels_resample.cuh"

namespace gb {

struct U128 { uint64_t hi, lo; };
GB_D U128 mul128(uint64_t a, uint64_t b) { return U128{__umul64hi(a, b), a * b}; }
GB_D bool gt128(const U128 &x, const U128 &y) { return x.hi > y.hi || (x.hi == y.hi && x.lo > y.lo); }

struct MsortPlan { uint64_t T, Q; };   // total of the spacings, total of the quantised weights

// exponential spacing of slot j, integer, >= 1 (mean 2^30: N <= 2^31 slots sum to < 2^62)
GB_D uint64_t spacing_from_philox(unsigned long long seed, unsigned long long j, unsigned int step, const MathTables &T) {
  const U4 w = philox_block(seed, j, step, PK_SORTED, 0);
  const double u = u52(w.x, w.y);                     // in (0, 1)
  return (uint64_t)(-gb_log(u, T) * 0x1p30) + 1ull;
}

// e[j] for j in [0, n] (GEN: generated and written; otherwise read), zero beyond; st_sum[s] = sum over slot tile s
template <bool GEN>
__global__ void __launch_bounds__(kBlock) msort_spacing_kernel(unsigned long long *__restrict__ e, long long n, long long nst,
                                                               unsigned long long seed, unsigned int step, long long j0,
                                                               unsigned long long *__restrict__ st_sum) {
  __shared__ double s_tab[kMathTabDoubles];
  __shared__ uint64_t s_w[kBlock / 32];
  const MathTables T = stage_math_tables(s_tab);
  __syncthreads();
  for (long long s = blockIdx.x; s < nst; s += gridDim.x) {
    uint64_t sum = 0;
#pragma unroll
    for (int i = 0; i < kItems; i++) {
      const long long j = s * kTile + (long long)i * kBlock + threadIdx.x;
      uint64_t v = 0;
      if (j <= n) v = GEN ? spacing_from_philox(seed, (unsigned long long)(j0 + j), step, T) : e[j];
      if (GEN || j > n) e[j] = v;
      sum += v;
    }
    sum = block_sum_u64(sum, s_w);
    if (threadIdx.x == 0 && st_sum) st_sum[s] = sum;
    __syncthreads();
  }
}

// exclusive scan of in[0..n) into out[0..n], out[n] = total; one CTA of 1024 threads (contiguous chunk per thread)
GB_D uint64_t msort_scan_array(const unsigned long long *__restrict__ in, unsigned long long *__restrict__ out, long long n, uint64_t *s_w /* 33 */) {
  const int Tn = blockDim.x;
  const long long per = (n + Tn - 1) / Tn;
  const long long lo = (long long)threadIdx.x * per, hi = lo + per < n ? lo + per : n;
  uint64_t a = 0;
  for (long long i = lo; i < hi; i++) a += __ldcg(in + i);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = Tn >> 5;
  uint64_t inc = a;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    uint64_t o = shfl_up_u64(inc, d);
    if (lane >= d) inc += o;
  }
  __syncthreads();
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
    const uint64_t v = __ldcg(in + i);
    out[i] = run;
    run += v;
  }
  if (threadIdx.x == 0) out[n] = total;
  return total;
}
__global__ void __launch_bounds__(1024) msort_scan_kernel(const unsigned long long *__restrict__ st_sum, unsigned long long *__restrict__ st_pref,
                                                          long long nst, const unsigned long long *__restrict__ tile_q,
                                                          unsigned long long *__restrict__ tile_cdf, long long ntiles, MsortPlan *plan) {
  __shared__ uint64_t s_w[33];
  const uint64_t T = msort_scan_array(st_sum, st_pref, nst, s_w);
  __syncthreads();
  const uint64_t Q = msort_scan_array(tile_q, tile_cdf, ntiles, s_w);
  if (threadIdx.x == 0) { plan->T = T; plan->Q = Q; }
}

// pt_lo[s] = the particle tile p with tile_cdf[p] * T <= st_pref[s] * Q < tile_cdf[p + 1] * T   (s = 0 .. nst)
__global__ void __launch_bounds__(kBlock) msort_partition_kernel(const unsigned long long *__restrict__ st_pref, long long nst,
                                                                 const unsigned long long *__restrict__ tile_cdf, long long ntiles,
                                                                 const MsortPlan *__restrict__ plan, int *__restrict__ pt_lo) {
  const uint64_t T = plan->T, Q = plan->Q;
  for (long long s = (long long)blockIdx.x * kBlock + threadIdx.x; s <= nst; s += (long long)gridDim.x * kBlock) {
    const U128 xq = mul128(st_pref[s], Q);
    long long lo = 0, hi = ntiles - 1;
    while (lo < hi) {
      const long long mid = (lo + hi + 1) >> 1;
      if (gt128(mul128(__ldg(tile_cdf + mid), T), xq)) hi = mid - 1; else lo = mid;
    }
    pt_lo[s] = (int)lo;
  }
}

// one CTA per slot tile s: ancestors of slots [2048 s, 2048 s + 2048) below n.
// The particle tile's CDF goes to shared memory ALREADY MULTIPLIED by T (128-bit, hi / lo arrays): a probe of the search is
// then two shared-memory loads and a compare instead of a 64 x 64 -> 128-bit multiply (the first version multiplied in every
// probe and searched every slot from scratch: 4.4 ms per 1e8 slots).  A thread's 8 slots are consecutive sorted thresholds,
// so after the first one (binary search) the next ancestor is a few particles further: linear probes, binary only as fallback.
constexpr int kMsPad = kTile + kTile / 8;            // one padding entry per 8: the blocked stores spread over the banks
GB_D int ms_idx(int i) { return i + (i >> 3); }
__global__ void __launch_bounds__(kBlock, 2) msort_fill_kernel(const unsigned long long *__restrict__ e, const unsigned long long *__restrict__ st_pref,
                                                               const unsigned long long *__restrict__ qbuf,
                                                               const unsigned long long *__restrict__ tile_cdf, long long ntiles,
                                                               const int *__restrict__ pt_lo, const MsortPlan *__restrict__ plan, long long n,
                                                               int *__restrict__ anc) {
  __shared__ uint64_t s_hi[kMsPad], s_lo[kMsPad];
  __shared__ uint64_t s_w1[kBlock / 32], s_w2[kBlock / 32];
  const long long s = blockIdx.x;
  const uint64_t T = plan->T, Q = plan->Q;
  const long long k0 = s * kTile + (long long)threadIdx.x * kItems;       // first slot of this thread
  // T_k of this thread's 8 consecutive slots
  uint64_t tk[kItems];
  {
    const ulonglong2 *src = reinterpret_cast<const ulonglong2 *>(e + k0);
    uint64_t run = 0;
#pragma unroll
    for (int i = 0; i < kItems; i += 2) {
      const ulonglong2 v = __ldcs(src + i / 2);
      run += v.x; tk[i] = run;
      run += v.y; tk[i + 1] = run;
    }
    uint64_t total;
    const uint64_t ex = block_exclusive_scan_1b(run, s_w1, &total) + __ldg(st_pref + s);
#pragma unroll
    for (int i = 0; i < kItems; i++) tk[i] += ex;
  }
  const int nvalid = n - k0 >= kItems ? kItems : (n - k0 > 0 ? (int)(n - k0) : 0);
  int a[kItems];
#pragma unroll
  for (int i = 0; i < kItems; i++) a[i] = 0;
  int i_cur = 0;
  const int p_lo = pt_lo[s];
  int p_hi = pt_lo[s + 1];
  if (p_hi > ntiles - 1) p_hi = (int)(ntiles - 1);
  for (int p = p_lo; p <= p_hi; p++) {
    // inclusive CDF of particle tile p, times T, in shared memory
    {
      const ulonglong2 *src = reinterpret_cast<const ulonglong2 *>(qbuf + (long long)p * kTile + (long long)threadIdx.x * kItems);
      uint64_t c[kItems], run = 0;
#pragma unroll
      for (int i = 0; i < kItems; i += 2) {
        const ulonglong2 v = __ldg(src + i / 2);
        run += v.x; c[i] = run;
        run += v.y; c[i + 1] = run;
      }
      uint64_t total;
      const uint64_t ex = block_exclusive_scan_1b(run, s_w2, &total) + __ldg(tile_cdf + p);
#pragma unroll
      for (int i = 0; i < kItems; i++) {
        const U128 pr = mul128(ex + c[i], T);
        const int ix = ms_idx(threadIdx.x * kItems + i);
        s_hi[ix] = pr.hi;
        s_lo[ix] = pr.lo;
      }
    }
    __syncthreads();
    const U128 end_t = U128{s_hi[ms_idx(kTile - 1)], s_lo[ms_idx(kTile - 1)]};     // the tile's last CDF value times T
    int pos = -1;                                             // (-1: no slot of this thread resolved in this tile yet)
    bool open = true;
#pragma unroll
    for (int i = 0; i < kItems; i++) {
      if (open && i >= i_cur && i < nvalid) {
        const U128 tq = mul128(tk[i], Q);
        if (gt128(end_t, tq)) {                              // the ancestor lies in this tile: first index whose C * T > T_k * Q
          int lo = pos < 0 ? 0 : pos, hi = kTile - 1;
          bool found = false;
          if (pos >= 0) {                                    // sorted thresholds: a few particles further than the last one
#pragma unroll
            for (int st = 0; st < 4; st++) {
              if (!found) {
                const int ix = ms_idx(lo);
                if (gt128(U128{s_hi[ix], s_lo[ix]}, tq)) found = true; else lo++;
              }
            }
          }
          if (!found) {
            while (lo < hi) {
              const int mid = (lo + hi) >> 1;
              const int ix = ms_idx(mid);
              if (gt128(U128{s_hi[ix], s_lo[ix]}, tq)) hi = mid; else lo = mid + 1;
            }
          }
          a[i] = p * kTile + lo;
          pos = lo;
          i_cur = i + 1;
        } else {
          open = false;                                      // sorted thresholds: the rest belongs to later tiles
        }
      }
    }
    __syncthreads();                                         // the shared arrays are rebuilt for the next tile
  }
  if (nvalid == kItems) {
    int4 *dst = reinterpret_cast<int4 *>(anc + k0);
    __stcs(dst, make_int4(a[0], a[1], a[2], a[3]));
    __stcs(dst + 1, make_int4(a[4], a[5], a[6], a[7]));
  } else {
#pragma unroll
    for (int i = 0; i < kItems; i++)
      if (i < nvalid) anc[k0 + i] = a[i];
  }
}

}  // namespace gb
