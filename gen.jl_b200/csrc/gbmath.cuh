// gbmath.cuh -- the engine's own fp64 elementary functions for the particle hot loop.
//
// Why not CUDA's libm: ncu on the first step kernel (profiles/r1_step_kernel_v0.md) showed the kernel
// ISSUE-bound, not fp64-pipe-bound: 487 instructions per particle, 22% of them UMOVs that rebuild the
// 64-bit immediates of libm's polynomials, plus special-case handling the hot loop can never reach
// (its arguments are Philox uniforms in (0,1) and finite states).  The functions below
//   * take every coefficient from __constant__ memory, which fp64 instructions read as a direct
//     operand (no UMOV pairs, no load),
//   * use small lookup tables (staged in shared memory by the calling kernel) to shorten the
//     polynomials: log via 91 breakpoints, atan2 via 65,
//   * consist only of integer ops, fma, IEEE +,*,/ and sqrt, so the host build of the same code gives
//     bit-identical results (tests/test_abi_cpu.py checks accuracy against mpmath on the CPU).
// Accuracy: <= 2 ulp (measured: tests), i.e. indistinguishable from libm at the 1e-10 weight tolerance.
#pragma once
#include "common.cuh"
#include "gbmath_tables.inc"

namespace gb {

// ---- coefficient storage: __constant__ on the device, static const on the host ----
#define GB_DECL_COEFFS(name, n, ...)                       \
  __constant__ double c_##name[n] = {__VA_ARGS__};         \
  static const double h_##name[n] = {__VA_ARGS__};
#ifdef __CUDA_ARCH__
#define KC(name, i) (c_##name[i])
#else
#define KC(name, i) (h_##name[i])
#endif

// sin / cos kernels on [-pi/4, pi/4] (classic fdlibm minimax coefficients)
GB_DECL_COEFFS(sin, 6, -1.66666666666666324348e-01, 8.33333333332248946124e-03, -1.98412698298579493134e-04,
               2.75573137070700676789e-06, -2.50507602534068634195e-08, 1.58969099521155010221e-10)
GB_DECL_COEFFS(cos, 6, 4.16666666666666019037e-02, -1.38888888888741095749e-03, 2.48015872894767294178e-05,
               -2.75573143513906633035e-07, 2.08757232129817482790e-09, -1.13596475577881948265e-11)
// log1p(r), |r| <= 1/256: r - r^2/2 + r^3/3 - ... + r^7/7
GB_DECL_COEFFS(log1p, 6, -0.5, 1.0 / 3.0, -0.25, 0.2, -1.0 / 6.0, 1.0 / 7.0)
// atan(t), |t| <= 1/128: t - t^3/3 + t^5/5 - t^7/7 + t^9/9
GB_DECL_COEFFS(atan, 4, -1.0 / 3.0, 0.2, -1.0 / 7.0, 1.0 / 9.0)
// exp(r), |r| <= ln2/2: Taylor to r^13 (general gb_exp; the weight quantiser has its own table-driven form below)
GB_DECL_COEFFS(exp, 12, 1.0 / 6227020800.0, 1.0 / 479001600.0, 1.0 / 39916800.0, 1.0 / 3628800.0, 1.0 / 362880.0,
               1.0 / 40320.0, 1.0 / 5040.0, 1.0 / 720.0, 1.0 / 120.0, 1.0 / 24.0, 1.0 / 6.0, 0.5)
GB_DECL_COEFFS(misc, 10, GB_LN2_HI, GB_LN2_LO, GB_PI_HI, GB_PI_LO, GB_PIO2_HI, GB_PIO2_LO,
               1.57079632679489655800e+00 /* pi/2 for the sincos argument */, 1.4426950408889634074 /* log2(e) */,
               -6.93147180369123816490e-01 /* -ln2 hi (exp reduction) */, -1.90821492927058770002e-10 /* -ln2 lo */)
enum { MI_LN2_HI = 0, MI_LN2_LO = 1, MI_PI_HI = 2, MI_PI_LO = 3, MI_PIO2_HI = 4, MI_PIO2_LO = 5, MI_PIO2 = 6, MI_LOG2E = 7,
       MI_NLN2_HI = 8, MI_NLN2_LO = 9 };

// lookup tables (source of truth in constant memory / host memory; kernels stage them into smem)
__constant__ double c_log_tab[2 * GB_LOG_TAB_N] = {GB_LOG_TAB_VALUES};
static const double h_log_tab[2 * GB_LOG_TAB_N] = {GB_LOG_TAB_VALUES};
__constant__ double c_atan_tab[GB_ATAN_TAB_N] = {GB_ATAN_TAB_VALUES};
static const double h_atan_tab[GB_ATAN_TAB_N] = {GB_ATAN_TAB_VALUES};

struct MathTables {
  const double *log_tab;   // 2 * GB_LOG_TAB_N doubles: {1/c_i, log c_i}
  const double *atan_tab;  // GB_ATAN_TAB_N doubles: atan(i/64)
};
constexpr int kMathTabDoubles = 2 * GB_LOG_TAB_N + GB_ATAN_TAB_N;

// stage the tables into shared memory (call from every thread of the CTA, then __syncthreads())
GB_D MathTables stage_math_tables(double *smem /* kMathTabDoubles */) {
  for (int i = threadIdx.x; i < 2 * GB_LOG_TAB_N; i += blockDim.x) smem[i] = c_log_tab[i];
  for (int i = threadIdx.x; i < GB_ATAN_TAB_N; i += blockDim.x) smem[2 * GB_LOG_TAB_N + i] = c_atan_tab[i];
  return MathTables{smem, smem + 2 * GB_LOG_TAB_N};
}
#ifndef __CUDACC_RTC__
inline __host__ MathTables host_math_tables() { return MathTables{h_log_tab, h_atan_tab}; }
#endif

GB_HD uint64_t dbits(double x) {
#ifdef __CUDA_ARCH__
  return (uint64_t)__double_as_longlong(x);
#else
  uint64_t b;
  memcpy(&b, &x, 8);
  return b;
#endif
}
GB_HD double bits2d(uint64_t b) {
#ifdef __CUDA_ARCH__
  return __longlong_as_double((long long)b);
#else
  double x;
  memcpy(&x, &b, 8);
  return x;
#endif
}

// 52-bit uniform in (0,1) from two Philox words without an integer->double conversion:
// [1,2) bit pattern minus 1, plus half a grid step: (v + 0.5) * 2^-52 exactly (== common.cuh u52)
GB_HD double u52_fast(uint32_t hi, uint32_t lo) {
  uint64_t v = (((uint64_t)hi << 32) | lo) >> 12;
  return (bits2d(0x3FF0000000000000ull | v) - 1.0) + 0x1p-53;
}

// log(x) for positive normal x.  x = 2^e * m, m in [sqrt(1/2), sqrt(2)); breakpoint c_i = 1 + i/128
// nearest to m; r = m / c_i - 1 (one fma with the tabulated reciprocal); log x = e ln2 + log c_i + log1p(r).
GB_HD double gb_log(double x, const MathTables &T) {
  const uint64_t b = dbits(x);
  const uint32_t hi = (uint32_t)(b >> 32);
  int e = (int)(hi >> 20) - 1023;
  const uint32_t mant = hi & 0xFFFFFu;
  int idx;
  uint32_t nhi;
  if (mant >= 0x6A09Fu) {          // m' >= ~sqrt(2): use m = m'/2 in [sqrt(1/2), 1)
    e += 1;
    nhi = mant | 0x3FE00000u;
    idx = (int)((mant + 0x2000u) >> 14) - 64;    // round(64 f) - 64  in [-37, 0]
  } else {
    nhi = mant | 0x3FF00000u;
    idx = (int)((mant + 0x1000u) >> 13);         // round(128 f)      in [0, 53]
  }
  const double m = bits2d(((uint64_t)nhi << 32) | (b & 0xFFFFFFFFull));
  const double inv_c = T.log_tab[2 * (idx - GB_LOG_TAB_LO)], log_c = T.log_tab[2 * (idx - GB_LOG_TAB_LO) + 1];
  const double r = fma(m, inv_c, -1.0);
  double p = KC(log1p, 5);
  p = fma(p, r, KC(log1p, 4));
  p = fma(p, r, KC(log1p, 3));
  p = fma(p, r, KC(log1p, 2));
  p = fma(p, r, KC(log1p, 1));
  p = fma(p, r, KC(log1p, 0));
  p = p * r;                       // log1p(r) = r + r * p
  const double ed = (double)e;
  // (e ln2_hi + log c) is exact-ish (ln2_hi has 21 significant bits); small terms added last
  const double big = fma(ed, KC(misc, MI_LN2_HI), log_c);
  const double small = fma(ed, KC(misc, MI_LN2_LO), fma(r, p, r));
  return big + small;
}

// sin(2 pi u), cos(2 pi u) for u in [0, 1]
GB_HD void gb_sincos2pi(double u, double &s, double &c) {
  const double a = 4.0 * u;                 // exact
  const double k = rint(a);                 // quadrant 0..4
  const double f = a - k;                   // exact, in [-1/2, 1/2]
  const double x = f * KC(misc, MI_PIO2);    // angle in [-pi/4, pi/4]
  const double z = x * x;
  double ps = KC(sin, 5);
  ps = fma(ps, z, KC(sin, 4));
  ps = fma(ps, z, KC(sin, 3));
  ps = fma(ps, z, KC(sin, 2));
  ps = fma(ps, z, KC(sin, 1));
  ps = fma(ps, z, KC(sin, 0));
  const double sx = fma(x * z, ps, x);
  double pc = KC(cos, 5);
  pc = fma(pc, z, KC(cos, 4));
  pc = fma(pc, z, KC(cos, 3));
  pc = fma(pc, z, KC(cos, 2));
  pc = fma(pc, z, KC(cos, 1));
  pc = fma(pc, z, KC(cos, 0));
  const double cx = fma(z * z, pc, fma(z, -0.5, 1.0));
  const int q = (int)k & 3;
  const double ss = (q & 1) ? cx : sx, cc = (q & 1) ? sx : cx;
  s = (q & 2) ? -ss : ss;
  c = ((q + 1) & 2) ? -cc : cc;
}

// exp(x) for |x| < 700
GB_HD double gb_exp_parts(double x, int &kout) {
  const double kd = rint(dmul(x, KC(misc, MI_LOG2E)));
  double r = fma(kd, KC(misc, MI_NLN2_HI), x);
  r = fma(kd, KC(misc, MI_NLN2_LO), r);
  double p = KC(exp, 0);
#pragma unroll
  for (int i = 1; i < 12; i++) p = fma(p, r, KC(exp, i));
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  kout = (int)kd;
  return p;
}
GB_HD double gb_exp(double x) {
  if (!(x > -700.0)) return x != x ? x : 0.0;
  if (x > 700.0) return INFINITY;
  int k;
  const double p = gb_exp_parts(x, k);
  return p * pow2i(k);
}

// ---- the weight exponential of the integer-CDF quantiser (shared bit for bit with the oracle) ----
// exp(x) for -708 < x <= 0:  x = j ln2/64 + r,  j = rint(x 64/ln2),  |r| <= ln2/128:
//     exp(x) = 2^(j >> 6) * T[j & 63] * (1 + r + r^2/2 + r^3/6 + r^4/24 + r^5/120)          (truncation < 4e-17)
// T = 2^(i/64) correctly rounded (64 doubles, staged in shared memory: every lane reads its own entry).  Six fp64
// operations in a dependency chain four deep, against the 14-deep Horner chain of the general gb_exp: the statistics
// kernel spent half of its stall samples waiting on that chain (profiles/r2j).  Only rint, fma, * and exact
// power-of-two scaling, so host, device and oracle produce identical bits.
__constant__ double c_wexp_tab[64] = {GB_WEXP_TAB_VALUES};
static const double h_wexp_tab[64] = {GB_WEXP_TAB_VALUES};
GB_DECL_COEFFS(wexp, 7, GB_WEXP_SCALE, GB_WEXP_NL_HI, GB_WEXP_NL_LO, 1.0 / 120.0, 1.0 / 24.0, 1.0 / 6.0, 0.5)
GB_D const double *stage_wexp_table(double *smem /* 64 */) {   // call from every thread of the CTA, then __syncthreads()
  for (int i = threadIdx.x; i < 64; i += blockDim.x) smem[i] = c_wexp_tab[i];
  return smem;
}
GB_HD double wexp_parts(double x, int &kout, const double *__restrict__ tab) {
  const double jd = rint(dmul(x, KC(wexp, 0)));
  const int j = (int)jd;
  double r = fma(jd, KC(wexp, 1), x);
  r = fma(jd, KC(wexp, 2), r);
  const double r2 = dmul(r, r);
  const double a = fma(r, KC(wexp, 3), KC(wexp, 4));
  const double b = fma(r, KC(wexp, 5), KC(wexp, 6));
  const double c = fma(r2, a, b);
  const double poly = fma(r2, c, r);
  const double t = tab[j & 63];
  kout = j >> 6;
  return fma(t, poly, t);
}
// Arguments of the weight kernels are x = lw - max: -708 < x <= -0.0 or x == +0.0 for every particle that carries weight.
// Anything else (NaN, -Inf, below the representable floor, or x > 0 when a caller passes a reference below the max) is
// rare and takes a branch instead of being folded into every evaluation as selects (the selects were 12 of the 65
// instructions per particle of the statistics kernel).  Integer test on the high word: for x <= -0.0 the magnitude
// ordering is the unsigned ordering of the bits.
GB_HD bool wx_regular(double x) {
  const uint32_t hi = (uint32_t)(dbits(x) >> 32);
  return ((hi ^ 0x80000000u) < 0x40862000u /* high word of 708.0 (low word 0): -708 < x <= -0.0 */) || x == 0.0;
}
// floor(exp(x) 2^b) for x = lw - max: bit-identical to the oracle's go_quantize()
GB_HD uint64_t quantize_k(double x, int b, const double *__restrict__ wt) {
  int k;
  if (wx_regular(x)) {
    const double p = wexp_parts(x, k, wt);
    const int e = k + b;
    const uint64_t q = (uint64_t)dmul(p, pow2i(e < 0 ? 0 : e));
    return e < 0 ? 0 : q;
  }
  if (!(x > -708.0)) return 0;                // NaN, -Inf, below the representable floor
  const double p = wexp_parts(0.0, k, wt);    // x > 0: clamped to 0
  return (uint64_t)dmul(p, pow2i(k + b));
}
// exp(x) and floor(exp(x) 2^b) from ONE evaluation (weight statistics pass)
GB_HD void exp_and_quantize(double x, int b, double &ex, uint64_t &q, const double *__restrict__ wt) {
  int k;
  if (wx_regular(x)) {
    const double p = wexp_parts(x, k, wt);
    const int e = k + b;
    const uint64_t qq = (uint64_t)dmul(p, pow2i(e < 0 ? 0 : e));
    q = e < 0 ? 0 : qq;
    ex = p * pow2i(k);
    return;
  }
  if (!(x > -708.0)) { q = 0; ex = 0.0; return; }
  const double p = wexp_parts(0.0, k, wt);
  q = (uint64_t)dmul(p, pow2i(k + b));
  ex = p * pow2i(k);
}

// atan2(y, x) for finite arguments whose exponents are within +-1000 (anything else: libm)
GB_HD double gb_atan2(double y, double x, const MathTables &T) {
  const double ay = fabs(y), ax = fabs(x);
  const bool swap = ay > ax;
  const double a = swap ? ax : ay, b = swap ? ay : ax;     // 0 <= a <= b
  const uint32_t bhi = (uint32_t)(dbits(b) >> 32);
  const uint32_t bexp = bhi >> 20;
  if (bexp < 24u || bexp > 2022u || !(a == a)) return atan2(y, x);   // zero / denormal / huge / inf / nan
  // index of the breakpoint t_i = i/64 nearest to a/b, from a single-precision quotient of the scaled values
  const uint64_t scale = (uint64_t)(2046u - bexp) << 52;              // 2^-(exponent of b)
  const float af = (float)(a * bits2d(scale)), bf = (float)(b * bits2d(scale));   // bf in [1,2), af in [0,2)
#ifdef __CUDA_ARCH__
  const int i = __float2int_rn(__fdividef(af, bf) * 64.0f);
#else
  const int i = (int)rintf((af / bf) * 64.0f);
#endif
  const double ti = (double)i * 0.015625;
  // tan(theta - theta_i) = (a - t_i b) / (b + t_i a), |.| <= ~1/128
  const double t = fma(-ti, b, a) / fma(ti, a, b);
  const double z = t * t;
  double p = KC(atan, 3);
  p = fma(p, z, KC(atan, 2));
  p = fma(p, z, KC(atan, 1));
  p = fma(p, z, KC(atan, 0));
  double r = T.atan_tab[i] + fma(t * z, p, t);               // in [0, pi/4]
  if (swap) r = (KC(misc, MI_PIO2_HI) - r) + KC(misc, MI_PIO2_LO);
  if (x < 0.0) r = (KC(misc, MI_PI_HI) - r) + KC(misc, MI_PI_LO);
  return y < 0.0 ? -r : r;
}

// fp64 noise: one Philox block -> one Box-Muller pair (spec shared with the oracle's philox_normal_pair:
// same 52-bit uniforms; r = sqrt(-2 log u1), (cos, sin)(2 pi u2))
GB_D void philox_normal_pair(uint64_t seed, uint64_t pid, uint32_t step, uint32_t pair, const MathTables &T, double &n0,
                             double &n1) {
  U4 w = philox_block(seed, pid, step, PK_NORMAL, pair);
  const double u1 = u52_fast(w.x, w.y), u2 = u52_fast(w.z, w.w);
  const double r = sqrt(-2.0 * gb_log(u1, T));
  double s, c;
  gb_sincos2pi(u2, s, c);
  n0 = r * c;
  n1 = r * s;
}
// fp32 noise: the same Philox block, transformed in single precision
GB_D void philox_normal_pair(uint64_t seed, uint64_t pid, uint32_t step, uint32_t pair, const MathTables &, float &n0,
                             float &n1) {
  U4 w = philox_block(seed, pid, step, PK_NORMAL, pair);
  float u1 = u24(w.x), u2 = u24(w.z);
  float r = sqrtf(-2.0f * __logf(u1));
  float s, c;
  sincospif(2.0f * u2, &s, &c);
  n0 = r * c;
  n1 = r * s;
}

}  // namespace gb
