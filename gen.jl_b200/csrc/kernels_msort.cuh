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
//   msort_scan       two CTAs: exclusive scan of the slot-tile sums (-> T) / of the PARTICLE-tile totals tile_q (-> Q)
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

// exponential spacings, integers >= 1 (mean 2^30: N <= 2^31 slots sum to < 2^62).  One Philox block (counter = J / 2) serves
// the two slots J = 2m and 2m + 1: words (x, y) and (z, w).
GB_D uint64_t spacing_from_u52(double u, const MathTables &T) { return (uint64_t)(-gb_log(u, T) * 0x1p30) + 1ull; }
GB_D void spacing_pair(unsigned long long seed, unsigned long long m, unsigned int step, const MathTables &T, uint64_t &e0, uint64_t &e1) {
  const U4 w = philox_block(seed, m, step, PK_SORTED, 0);
  e0 = spacing_from_u52(u52(w.x, w.y), T);
  e1 = spacing_from_u52(u52(w.z, w.w), T);
}
GB_D uint64_t spacing_one(unsigned long long seed, unsigned long long J, unsigned int step, const MathTables &T) {
  const U4 w = philox_block(seed, J >> 1, step, PK_SORTED, 0);
  return (J & 1) ? spacing_from_u52(u52(w.z, w.w), T) : spacing_from_u52(u52(w.x, w.y), T);
}

// e[j] for j in [0, n] (GEN: generated for global slot j0 + j and written; otherwise read), zero beyond; st_sum[s] = sum
// over slot tile s
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
    if (GEN && (j0 & 1) == 0) {                     // two consecutive slots per thread and Philox block, 128-bit stores
#pragma unroll
      for (int i = 0; i < kItems / 2; i++) {
        const long long j = s * kTile + (long long)i * (2 * kBlock) + 2 * threadIdx.x;
        uint64_t v0 = 0, v1 = 0;
        if (j <= n) spacing_pair(seed, (unsigned long long)(j0 + j) >> 1, step, T, v0, v1);
        if (j + 1 > n) v1 = 0;
        *reinterpret_cast<ulonglong2 *>(e + j) = make_ulonglong2(v0, v1);
        sum += v0 + v1;
      }
    } else {
#pragma unroll
      for (int i = 0; i < kItems; i++) {
        const long long j = s * kTile + (long long)i * kBlock + threadIdx.x;
        uint64_t v = 0;
        if (j <= n) v = GEN ? spacing_one(seed, (unsigned long long)(j0 + j), step, T) : e[j];
        if (GEN || j > n) e[j] = v;
        sum += v;
      }
    }
    sum = block_sum_u64(sum, s_w);
    if (threadIdx.x == 0 && st_sum) st_sum[s] = sum;
    __syncthreads();
  }
}

// exclusive scan of in[0..n) into out[0..n], out[n] = total; one CTA of 1024 threads walking the array in chunks of 4096
// (four consecutive elements per thread: coalesced), carry in a register.  (First version: one contiguous chunk of n / 1024
// elements per thread -- every load instruction of a warp touched 32 different sectors, 0.2 ms for two arrays of 49 k entries.)
GB_D uint64_t msort_scan_array(const unsigned long long *__restrict__ in, unsigned long long *__restrict__ out, long long n, uint64_t *s_w /* 33 */) {
  const int Tn = blockDim.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = Tn >> 5;
  uint64_t carry = 0;
  for (long long base = 0; base < n; base += (long long)Tn * 4) {
    const long long i0 = base + (long long)threadIdx.x * 4;
    uint64_t v[4];
#pragma unroll
    for (int k = 0; k < 4; k++) v[k] = i0 + k < n ? __ldcg(in + i0 + k) : 0;
    const uint64_t a = v[0] + v[1] + v[2] + v[3];
    uint64_t inc = a;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      uint64_t o = shfl_up_u64(inc, d);
      if (lane >= d) inc += o;
    }
    __syncthreads();                                   // (s_w of the previous chunk has been read)
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
    uint64_t run = carry + s_w[warp] + (inc - a);
#pragma unroll
    for (int k = 0; k < 4; k++) {
      if (i0 + k < n) out[i0 + k] = run;
      run += v[k];
    }
    carry += s_w[32];
  }
  __syncthreads();
  if (threadIdx.x == 0) out[n] = carry;
  return carry;
}
__global__ void __launch_bounds__(1024) msort_scan_kernel(const unsigned long long *__restrict__ st_sum, unsigned long long *__restrict__ st_pref,
                                                          long long nst, const unsigned long long *__restrict__ tile_q,
                                                          unsigned long long *__restrict__ tile_cdf, long long ntiles, MsortPlan *plan) {
  __shared__ uint64_t s_w[33];
  if (blockIdx.x == 0) {                      // (two CTAs: the two scans are independent)
    const uint64_t T = msort_scan_array(st_sum, st_pref, nst, s_w);
    if (threadIdx.x == 0) plan->T = T;
  } else {
    const uint64_t Q = msort_scan_array(tile_q, tile_cdf, ntiles, s_w);
    if (threadIdx.x == 0) plan->Q = Q;
  }
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
// Every threshold is converted to CDF units once -- tau_k = floor(T_k * Q / T), one exact 128/64-bit division (div128_64,
// common.cuh: double-precision estimates + exact remainder fix-up) -- because for an integer C,  C * T > T_k * Q  <=>
// C > tau_k.  The particle tiles' CDF then sits in shared memory as plain 64-bit values and a probe of the search is one load
// and one compare.  (Versions before this one kept C * T as 128-bit products in shared memory: 16 multiplies per thread and
// pass, twice the shared memory; before that, one multiply per probe -- profiles/README.md r2as.)
// kMsTiles particle tiles per pass (a slot tile usually straddles two): one build phase and one search phase in which every
// thread has work, instead of one pass per tile with half of the CTA waiting at the barriers.  A thread's 8 slots are
// consecutive sorted thresholds, so after the first one (binary search) the next ancestor is a few particles further: linear
// probes, binary only as fallback.
#ifndef GB_MSORT_MINB
#define GB_MSORT_MINB 5
#endif
#ifndef GB_MSORT_TILES
#define GB_MSORT_TILES 2
#endif
constexpr int kMsTiles = GB_MSORT_TILES;                           // particle tiles per pass
constexpr int kMsEntries = kMsTiles * kTile;
constexpr int kMsPad = kMsEntries + kMsEntries / 8;                // one padding entry per 8: the blocked stores spread over the banks
GB_D int ms_idx(int i) { return i + (i >> 3); }
__global__ void __launch_bounds__(kBlock, GB_MSORT_MINB) msort_fill_kernel(const unsigned long long *__restrict__ e, const unsigned long long *__restrict__ st_pref,
                                                               const unsigned long long *__restrict__ qbuf,
                                                               const unsigned long long *__restrict__ tile_cdf, long long ntiles,
                                                               const int *__restrict__ pt_lo, const MsortPlan *__restrict__ plan, long long n,
                                                               int *__restrict__ anc) {
  extern __shared__ __align__(16) unsigned char ms_smem[];
  uint64_t *s_cs = reinterpret_cast<uint64_t *>(ms_smem);
  __shared__ uint64_t s_w1[kBlock / 32], s_w2[kMsTiles][kBlock / 32];
  const long long s = blockIdx.x;
  const uint64_t T = plan->T, Q = plan->Q;
  const double t_inv = 1.0 / (double)T;
  const long long k0 = s * kTile + (long long)threadIdx.x * kItems;       // first slot of this thread
  // tau_k of this thread's 8 consecutive slots
  uint64_t tau[kItems];
  {
    const ulonglong2 *src = reinterpret_cast<const ulonglong2 *>(e + k0);
    uint64_t run = 0;
#pragma unroll
    for (int i = 0; i < kItems; i += 2) {
      const ulonglong2 v = __ldcs(src + i / 2);
      run += v.x; tau[i] = run;
      run += v.y; tau[i + 1] = run;
    }
    uint64_t total;
    const uint64_t ex = block_exclusive_scan_1b(run, s_w1, &total) + __ldg(st_pref + s);
#pragma unroll
    for (int i = 0; i < kItems; i++) {
      const uint64_t tk = tau[i] + ex;                     // T_k <= T
      tau[i] = div128_64(mulhi64(tk, Q), tk * Q, T, t_inv);
    }
  }
  const int nvalid = n - k0 >= kItems ? kItems : (n - k0 > 0 ? (int)(n - k0) : 0);
  int a[kItems];
#pragma unroll
  for (int i = 0; i < kItems; i++) a[i] = 0;
  int i_cur = 0;
  const int p_lo = pt_lo[s];
  int p_hi = pt_lo[s + 1];
  if (p_hi > ntiles - 1) p_hi = (int)(ntiles - 1);
  for (int p = p_lo; p <= p_hi; p += kMsTiles) {
    const int nt = p_hi - p + 1 < kMsTiles ? p_hi - p + 1 : kMsTiles;     // particle tiles in this pass
    // inclusive CDF of the pass's particle tiles in shared memory
#pragma unroll
    for (int tt = 0; tt < kMsTiles; tt++) {
      if (tt < nt) {
        const ulonglong2 *src = reinterpret_cast<const ulonglong2 *>(qbuf + (long long)(p + tt) * kTile + (long long)threadIdx.x * kItems);
        uint64_t c[kItems], run = 0;
#pragma unroll
        for (int i = 0; i < kItems; i += 2) {
          const ulonglong2 v = __ldg(src + i / 2);
          run += v.x; c[i] = run;
          run += v.y; c[i + 1] = run;
        }
        uint64_t total;
        const uint64_t ex = block_exclusive_scan_1b(run, s_w2[tt], &total) + __ldg(tile_cdf + p + tt);
#pragma unroll
        for (int i = 0; i < kItems; i++) s_cs[ms_idx(tt * kTile + threadIdx.x * kItems + i)] = ex + c[i];
      }
    }
    __syncthreads();
    const int last = nt * kTile - 1;
    const uint64_t end_c = s_cs[ms_idx(last)];                // the pass's last CDF value
    int pos = -1;                                             // (-1: no slot of this thread resolved in this pass yet)
    bool open = true;
#pragma unroll
    for (int i = 0; i < kItems; i++) {
      if (open && i >= i_cur && i < nvalid) {
        const uint64_t t = tau[i];
        if (end_c > t) {                                     // the ancestor lies in this pass: first index whose C > tau_k
          int lo = pos < 0 ? 0 : pos, hi = last;
          bool found = false;
          if (pos >= 0) {                                    // sorted thresholds: a few particles further than the last one
            for (int st = 0; st < 4 && !found; st++) {
              if (s_cs[ms_idx(lo)] > t) found = true; else lo++;
            }
          }
          if (!found) {
            while (lo < hi) {
              const int mid = (lo + hi) >> 1;
              if (s_cs[ms_idx(mid)] > t) hi = mid; else lo = mid + 1;
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
    __syncthreads();                                         // the shared array is rebuilt for the next pass
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
