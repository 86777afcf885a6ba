// common.cuh -- shared device/host helpers of libgenb200 (sm_100a).
//
// Compiled with -fmad=false: Julia never contracts `mu + std * randn()` (normal.jl:128), so every
// a*b+c written in this tree stays two roundings; where a fused multiply-add is wanted it is written
// as fma() explicitly.
#pragma once
#ifndef __CUDACC_RTC__
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include <string.h>
#else
// NVRTC (run-time specialisation of GB_MODEL_SPEC programs, csrc/spec_jit.inc): no host headers
typedef unsigned int uint32_t;
typedef int int32_t;
typedef unsigned long long uint64_t;
typedef long long int64_t;
typedef unsigned char uint8_t;
typedef unsigned long size_t;
#define INFINITY __int_as_float(0x7f800000)
#define NAN __int_as_float(0x7fffffff)
#define M_PI 3.14159265358979323846
#endif

typedef unsigned __int128 u128;

#define GB_HD __host__ __device__ __forceinline__
#define GB_D __device__ __forceinline__

namespace gb {

// Programmatic dependent launch (sm_90+): a kernel launched with the programmatic-stream-serialization attribute may
// start while its predecessor in the stream drains; it must pdl_wait() before touching anything the predecessor wrote.
// pdl_trigger() lets the successor's CTAs be scheduled as this grid's CTAs retire.  Both are no-ops for plain launches.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

constexpr int kBlock = 256;          // threads per CTA for every streaming kernel
constexpr int kTile = 2048;          // particles per scan tile (kBlock * kItems)
constexpr int kItems = kTile / kBlock;
constexpr int kChunk = 4 * kTile;    // output slots one ancestors CTA fills before overflow CTAs take over
constexpr double kLog2Pi = 1.8378770664093454835606594728112;  // log(2*pi)

// ------------------------------------------------------------------------------------------
// Philox4x32-10.  Counter = (pid_lo, pid_hi, step, kind<<24 | idx), key = seed: a draw is a pure
// function of (seed, step, global particle id, draw index) and so independent of sharding and of the
// launch geometry.  Same integer spec as the oracle (oracle/gen_oracle.c go_philox4x32_10).
// ------------------------------------------------------------------------------------------
enum { PK_NORMAL = 1, PK_UNIFORM = 2, PK_RESAMPLE = 3, PK_MULTI = 4, PK_GAMMA = 5, PK_SAMPLE = 6, PK_SORTED = 7 };

struct U4 { uint32_t x, y, z, w; };

GB_HD uint32_t mulhi32(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
  return __umulhi(a, b);
#else
  return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}

GB_HD U4 philox4x32_10(U4 c, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
    uint32_t hi0 = mulhi32(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    uint32_t hi1 = mulhi32(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    U4 n;
    n.x = hi1 ^ c.y ^ k0;
    n.y = lo1;
    n.z = hi0 ^ c.w ^ k1;
    n.w = lo0;
    c = n;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return c;
}

GB_HD U4 philox_block(uint64_t seed, uint64_t pid, uint32_t step, uint32_t kind, uint32_t idx) {
  U4 c = {(uint32_t)pid, (uint32_t)(pid >> 32), step, (kind << 24) | (idx & 0xFFFFFFu)};
  return philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
}

// 52-bit uniform strictly inside (0,1); exact in fp64
GB_HD double u52(uint32_t hi, uint32_t lo) {
  uint64_t v = (((uint64_t)hi << 32) | lo) >> 12;
  return ((double)v + 0.5) * 0x1p-52;
}
// 24-bit uniform strictly inside (0,1); exact in fp32
GB_HD float u24(uint32_t w) { return ((float)(w >> 8) + 0.5f) * 0x1p-24f; }

GB_D void philox_uniform_pair(uint64_t seed, uint64_t pid, uint32_t step, uint32_t pair, double &a, double &b) {
  U4 w = philox_block(seed, pid, step, PK_UNIFORM, pair);
  a = u52(w.x, w.y);
  b = u52(w.z, w.w);
}
GB_D void philox_uniform_pair(uint64_t seed, uint64_t pid, uint32_t step, uint32_t pair, float &a, float &b) {
  double x, y;
  philox_uniform_pair(seed, pid, step, pair, x, y);
  a = (float)x;
  b = (float)y;
  // (float)u52 can round up to 1.0f; keep the draw inside [0,1)
  a = a >= 1.0f ? 0x1.fffffep-1f : a;
  b = b >= 1.0f ? 0x1.fffffep-1f : b;
}

// ------------------------------------------------------------------------------------------
// Integer-CDF weight quantisation (oracle go_dexp/go_quantize): helpers; the
// exponential itself is gbmath.cuh wexp_parts.
// ------------------------------------------------------------------------------------------
GB_HD double dmul(double a, double b) {
#ifdef __CUDA_ARCH__
  return __dmul_rn(a, b);
#else
  return a * b;
#endif
}
GB_HD double pow2i(int e) {
  uint64_t bits = (uint64_t)(e + 1023) << 52;
#ifdef __CUDA_ARCH__
  return __longlong_as_double((long long)bits);
#else
  double d;
  memcpy(&d, &bits, 8);
  return d;
#endif
}
#ifndef __CUDACC_RTC__
inline __host__ int quant_bits(int64_t n) {
  int lg = 0;
  while (((int64_t)1 << lg) < n) lg++;
  int b = 62 - lg;
  return b > 52 ? 52 : b;
}
#endif

// ------------------------------------------------------------------------------------------
// 128 / 64 -> 64 bit floor division via a double reciprocal and exact integer correction.
// Requires d >= 1 and floor(a / d) < 2^63.
// ------------------------------------------------------------------------------------------
GB_HD uint64_t mulhi64(uint64_t a, uint64_t b) {
#ifdef __CUDA_ARCH__
  return __umul64hi(a, b);
#else
  return (uint64_t)(((u128)a * b) >> 64);
#endif
}
GB_HD double u128_to_double(uint64_t hi, uint64_t lo) {
  // round-to-nearest of each half is enough: the estimate is corrected exactly below
  return (double)hi * 18446744073709551616.0 + (double)lo;
}
GB_HD uint64_t div128_64(uint64_t ahi, uint64_t alo, uint64_t d, double dinv) {
  // first estimate: relative error ~2^-50, i.e. up to ~2^13 units when the quotient is ~2^63
  double qd = u128_to_double(ahi, alo) * dinv;
  uint64_t q = qd >= 9223372036854775807.0 ? 0x7FFFFFFFFFFFFFFFull : (uint64_t)qd;
  // remainder r = a - q*d as a signed 128-bit number (|r| < 2^14 * d < 2^77)
  uint64_t plo = q * d, phi = mulhi64(q, d);
  uint64_t rlo = alo - plo;
  int64_t rhi = (int64_t)(ahi - phi - (alo < plo ? 1 : 0));
  // second estimate on the (small) remainder
  double rd = (double)rhi * 18446744073709551616.0 + (double)rlo;
  int64_t dq = (int64_t)(rd * dinv);
  q += (uint64_t)dq;
  // r -= dq * d  (dq*d fits in signed 128)
  {
    uint64_t m = dq < 0 ? (uint64_t)(-dq) : (uint64_t)dq;
    uint64_t mlo = m * d, mhi = mulhi64(m, d);
    if (dq >= 0) {
      uint64_t nlo = rlo - mlo;
      rhi = rhi - (int64_t)mhi - (rlo < mlo ? 1 : 0);
      rlo = nlo;
    } else {
      uint64_t nlo = rlo + mlo;
      rhi = rhi + (int64_t)mhi + (nlo < rlo ? 1 : 0);
      rlo = nlo;
    }
  }
  // final exact fix-up (a couple of iterations at most)
  while (rhi < 0) {
    uint64_t nlo = rlo + d;
    rhi += (nlo < rlo ? 1 : 0);
    rlo = nlo;
    q--;
  }
  while (rhi > 0 || rlo >= d) {
    uint64_t nlo = rlo - d;
    rhi -= (rlo < d ? 1 : 0);
    rlo = nlo;
    q++;
  }
  return q;
}

// Parameters of one systematic sweep over an integer CDF.
//   regular  (Appendix B "systematic"): tau_k = floor((k*2^32 + u0) * Q / (nslots * 2^32))
//   residual (Appendix B "residual")  : tau_k = floor((k*2^32 + u0) * Q / 2^32)   (CDF of N*q - c*Q)
struct SweepParams {
  uint64_t q_total;   // Q_N
  double q_inv;       // 1.0 / (double)Q_N
  uint64_t dscale;    // nslots << 32 (regular) ; unused for residual
  double ratio;       // (double)nslots / (double)Q_N: slots per unit of integer CDF mass
  uint32_t u0;
};
// number of slots k with tau_k < C  (regular sweep; C <= Q_N < 2^62, dscale < 2^63)
GB_HD int64_t slots_below(uint64_t C, const SweepParams &sp) {
  if (C == 0) return 0;
  // A = C * dscale - 1
  uint64_t alo = C * sp.dscale, ahi = mulhi64(C, sp.dscale);
  ahi -= (alo == 0 ? 1 : 0);
  alo -= 1;
  uint64_t X = div128_64(ahi, alo, sp.q_total, sp.q_inv);
  if (X < sp.u0) return 0;
  return (int64_t)((X - sp.u0) >> 32) + 1;
}
// Same count through a double-precision estimate, falling back to the exact division only when the
// estimate lands within 2^-17 slots of a slot boundary (probability ~1.5e-5 per call): slot coordinate
// y = (C*D/Q - u0) / 2^32, count = ceil(y) for y > 0.  |estimate - y| <= y * 2^-50 <= 2^-19.
GB_HD int64_t slots_below_fast(uint64_t C, const SweepParams &sp) {
  const double y = fma((double)C, sp.ratio, -(double)sp.u0 * 0x1p-32);
  const double c = ceil(y);
  if (y < -0x1p-17) return 0;
  if (c - y > 0x1p-17 && y - (c - 1.0) > 0x1p-17) return (int64_t)c;
  return slots_below(C, sp);
}
// residual sweep: C is a 128-bit CDF value (< 2^92): A = (C << 32) - 1
GB_HD int64_t slots_below_residual(uint64_t chi, uint64_t clo, const SweepParams &sp) {
  if ((chi | clo) == 0) return 0;
  uint64_t ahi = (chi << 32) | (clo >> 32), alo = clo << 32;
  ahi -= (alo == 0 ? 1 : 0);
  alo -= 1;
  uint64_t X = div128_64(ahi, alo, sp.q_total, sp.q_inv);
  if (X < sp.u0) return 0;
  return (int64_t)((X - sp.u0) >> 32) + 1;
}

// ------------------------------------------------------------------------------------------
// online log-sum-exp triple (m, s1 = sum exp(lw-m), s2 = sum exp(2(lw-m)))
// (inference.jl:3-6 and particle_filter.jl:3-12 in one pass: lse = m + log s1, ess = s1^2 / s2)
// ------------------------------------------------------------------------------------------
struct Lse3 { double m, s1, s2; };
GB_HD Lse3 lse3_empty() { return Lse3{-INFINITY, 0.0, 0.0}; }
GB_D void lse3_push(Lse3 &a, double lw) {
  if (lw == -INFINITY || lw != lw) {
    if (lw != lw) { a.m = lw; }  // NaN poisons the max exactly like Julia's maximum()
    return;
  }
  double d = lw - a.m;
  double e = exp(-fabs(d));   // a.m == -inf: d = +inf, e = 0
  double e2 = e * e;
  if (d > 0.0) {
    a.s1 = a.s1 * e + 1.0;
    a.s2 = a.s2 * e2 + 1.0;
    a.m = lw;
  } else {
    a.s1 += e;
    a.s2 += e2;
  }
}
GB_D Lse3 lse3_merge(const Lse3 &a, const Lse3 &b) {
  if (a.m != a.m) return a;
  if (b.m != b.m) return b;
  if (b.m == -INFINITY) return a;
  if (a.m == -INFINITY) return b;
  double m = fmax(a.m, b.m);
  double ea = exp(a.m - m), eb = exp(b.m - m);
  return Lse3{m, a.s1 * ea + b.s1 * eb, a.s2 * ea * ea + b.s2 * eb * eb};
}
GB_D Lse3 lse3_warp_reduce(Lse3 v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    Lse3 w;
    w.m = __shfl_xor_sync(0xffffffffu, v.m, o);
    w.s1 = __shfl_xor_sync(0xffffffffu, v.s1, o);
    w.s2 = __shfl_xor_sync(0xffffffffu, v.s2, o);
    v = lse3_merge(v, w);
  }
  return v;
}

// streaming (read-once / write-once) global accesses: keep the 126 MB L2 for data that is re-read
template <typename T> GB_D T ld_stream(const T *p) { return __ldcs(p); }
template <typename T> GB_D void st_stream(T *p, T v) { __stcs(p, v); }

}  // namespace gb
