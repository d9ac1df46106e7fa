// Double-precision ln / exp for the PEAQ kernels (sm_100a).
//
// The reference calls libm's pow / exp / log / log10 (earmodel.c:890-907, fftearmodel.c:647-675,
// modpatt.c:234, movs.c:725-738, :1240-1260, :1396-1403).  The kernels evaluate x^y as
// exp(y ln x), and every such call goes through the two functions below: ~6 000 evaluations per
// PEAQ frame.  CUDA's own log()/exp() cost 82 and 47 SASS instructions per call on sm_100a (ncu
// source page of round 1), most of it special-case handling; the versions here take 36 and 24 on
// their fast path and branch to the library functions for everything the fast path does not cover
// (zero, denormal, negative, infinite, NaN arguments; exp arguments beyond +-708), so the result
// classes of the library functions are preserved exactly.
//
// Accuracy of the fast paths (tests/test_fast_math.py, against long double libm over 2^22 arguments
// per function): max relative error 2.3e-16 (exp), 3.3e-16 (ln; absolute 2.2e-16 ln 2 near 1) --
// the same 1..2 ulp class as the library functions, against a parity bar of 1e-6 on the MOVs.
//
//   ln x : x = 2^e m, m in [sqrt(1/2), sqrt(2));  f = (m-1)/(m+1)  (division by Newton iterations
//          on MUFU.RCP64H, residual-corrected);  ln m = 2 f + 2 f s g(s), s = f^2, g = minimax
//          polynomial of (atanh(sqrt s)/sqrt s - 1)/s of degree 6 (error 1.6e-16 of g, 5e-18 of ln m)
//   e^x  : k = rint(x log2 e), r = x - k ln 2 (two-constant Cody-Waite, exact first step);
//          e^r = 1 + r + r^2 q(r), q of degree 9 (Chebyshev fit, error 1.3e-17 of e^r);  2^k by
//          adding k to the exponent field (result normal for |x| < 708)
// The file also compiles as host C++ (PEAQ_MATH_HOST) for the accuracy test; host and device run
// the same sequence of IEEE operations except for the reciprocal seed.
#pragma once

#if defined(PEAQ_MATH_HOST)
#include <cmath>
#include <cstdint>
#include <cstring>
#define PEAQ_MATH_FN inline
namespace peaq {
inline int pm_hi(double x) { uint64_t u; std::memcpy(&u, &x, 8); return (int)(u >> 32); }
inline int pm_lo(double x) { uint64_t u; std::memcpy(&u, &x, 8); return (int)(u & 0xffffffffu); }
inline double pm_make(int hi, int lo) {
  uint64_t u = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo;
  double x; std::memcpy(&x, &u, 8); return x;
}
// stand-ins for MUFU.RCP64H / RSQ64H: full exponent range, ~20 good mantissa bits
inline double pm_trunc20(double v) { int e; double m = std::frexp(v, &e); return std::ldexp(std::floor(m * 1048576.0) / 1048576.0, e); }
inline double pm_rcp_seed(double x) { return pm_trunc20(1.0 / x); }
inline double pm_rsqrt_seed(double x) { return pm_trunc20(1.0 / std::sqrt(x)); }
inline double pm_fma(double a, double b, double c) { return std::fma(a, b, c); }
inline double pm_slow_log(double x) { return std::log(x); }
inline double pm_slow_exp(double x) { return std::exp(x); }
}  // namespace peaq
#else
#include <cuda_runtime.h>
#define PEAQ_MATH_FN __device__ __forceinline__
namespace peaq {
__device__ __forceinline__ int pm_hi(double x) { return __double2hiint(x); }
__device__ __forceinline__ int pm_lo(double x) { return __double2loint(x); }
__device__ __forceinline__ double pm_make(int hi, int lo) { return __hiloint2double(hi, lo); }
__device__ __forceinline__ double pm_rcp_seed(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  return y;
}
__device__ __forceinline__ double pm_rsqrt_seed(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  return y;
}
__device__ __forceinline__ double pm_fma(double a, double b, double c) { return fma(a, b, c); }
static __device__ __noinline__ double pm_slow_log(double x) { return log(x); }
static __device__ __noinline__ double pm_slow_exp(double x) { return exp(x); }
}  // namespace peaq
#endif

namespace peaq {

#if defined(PEAQ_LIBM_MATH)   // development switch: the CUDA library functions (bit-identical to round 1)
PEAQ_MATH_FN double peaq_log(double x) { return log(x); }
PEAQ_MATH_FN double peaq_exp(double x) { return exp(x); }
PEAQ_MATH_FN double peaq_log10(double x) { return log10(x); }
#else
// ln x
PEAQ_MATH_FN double peaq_log(double x) {
  int hi = pm_hi(x);
  // fast path: positive, normal, finite
  if ((unsigned)(hi - 0x00100000) >= 0x7fe00000u) return pm_slow_log(x);
  int e = (hi >> 20) - 1023;
  hi = (hi & 0x000fffff) | 0x3ff00000;         // m in [1, 2)
  if (hi >= 0x3ff6a09f) {                      // m > ~sqrt 2: halve it
    hi -= 0x00100000;
    e += 1;
  }
  const double m = pm_make(hi, pm_lo(x));
  const double num = m - 1.0;                  // exact
  const double den = m + 1.0;
  double r = pm_rcp_seed(den);
  r = pm_fma(pm_fma(-den, r, 1.0), r, r);
  r = pm_fma(pm_fma(-den, r, 1.0), r, r);
  double f = num * r;
  f = pm_fma(pm_fma(-f, den, num), r, f);      // residual correction: f = num / den to ~0.5 ulp
  const double s = f * f;
  double g = 0.073115919867018450491;
  g = pm_fma(g, s, 0.076655829258400204968);
  g = pm_fma(g, s, 0.090914531170014184225);
  g = pm_fma(g, s, 0.11111105446758570067);
  g = pm_fma(g, s, 0.14285714313751241874);
  g = pm_fma(g, s, 0.1999999999994800026);
  g = pm_fma(g, s, 0.33333333333333349075);
  const double u = f + f;
  const double w = (u * s) * g;
  const double ed = (double)e;
  const double hi_part = pm_fma(ed, 0.6931471803691238, u);           // e * ln2_hi (32 bits) is exact
  const double lo_part = pm_fma(ed, 1.9082149292705877e-10, w);
  return hi_part + lo_part;
}

// e^x
PEAQ_MATH_FN double peaq_exp(double x) {
  if (!(fabs(x) < 708.0)) {
    // far underflow is common on this path (0.5^((e/s)^b) of the detection probability, movs.c:1256-1258;
    // the beta of the noise loudness, movs.c:731): exact zero without the library call
    if (x < -746.0) return 0.0;
    return pm_slow_exp(x);
  }
  const double magic = 6755399441055744.0;      // 1.5 * 2^52: rint through the adder
  const double t = pm_fma(x, 1.4426950408889634, magic);
  const int k = pm_lo(t);
  const double kd = t - magic;
  double r = pm_fma(kd, -0.6931471803691238, x);
  r = pm_fma(kd, -1.9082149292705877e-10, r);
  double q = 2.5100385495510319078e-8;
  q = pm_fma(q, r, 2.7620088445409748161e-7);
  q = pm_fma(q, r, 2.7557268459997064772e-6);
  q = pm_fma(q, r, 0.000024801521295954375131);
  q = pm_fma(q, r, 0.00019841269863053616878);
  q = pm_fma(q, r, 0.0013888888917213716901);
  q = pm_fma(q, r, 0.0083333333333300618325);
  q = pm_fma(q, r, 0.041666666666624127873);
  q = pm_fma(q, r, 0.16666666666666667453);
  q = pm_fma(q, r, 0.50000000000000010221);
  const double p = 1.0 + pm_fma(r * r, q, r);   // in [0.70, 1.42]
  return pm_make(pm_hi(p) + (k << 20), pm_lo(p));
}

// log10 x = ln x / ln 10
PEAQ_MATH_FN double peaq_log10(double x) { return peaq_log(x) * 0.4342944819032518; }
#endif

// ---- branch-free variants for arguments the caller knows to be ordinary ----------------------
// The library's division and square root (and the functions above) branch to a slow path for
// special operands; a branch ends the instruction scheduler's window, so two independent quotients
// or logarithms in a row run one after the other -- measured on B200: 127 cycles per DDIV and 94 per
// DSQRT whether one or four independent ones are in flight, against 8.2 cycles per dependent
// DFMA.  The recurrent kernels are bound by exactly these dependent chains.  The variants below
// have no branch (independent ones interleave) and a third of the instructions.  They are NOT
// correctly rounded: error <= 1 ulp (tests/test_fast_math.py), against a parity bar of 1e-6.

// a / b for normal finite b != 0 with |b| in [2^-1000, 2^1000] and a finite quotient
PEAQ_MATH_FN double peaq_div(double a, double b) {
  double r = pm_rcp_seed(b);
  r = pm_fma(pm_fma(-b, r, 1.0), r, r);
  r = pm_fma(pm_fma(-b, r, 1.0), r, r);
  const double q = a * r;
  return pm_fma(pm_fma(-q, b, a), r, q);
}

// sqrt x for finite x >= 0 (denormal x gives 0: an absolute error below 1.5e-154)
PEAQ_MATH_FN double peaq_sqrt(double x) {
  double y = pm_rsqrt_seed(x);                 // ~2^-20
  double h = 0.5 * y;
  double e = pm_fma(-x * y, h, 0.5);           // 0.5 - 0.5 x y^2
  y = pm_fma(y, e, y);
  h = 0.5 * y;
  e = pm_fma(-x * y, h, 0.5);
  y = pm_fma(y, e, y);                         // 1/sqrt(x) to ~1 ulp
  const double s = x * y;
  const double res = pm_fma(pm_fma(-s, s, x), 0.5 * y, s);
  return x >= 2.2250738585072014e-308 ? res : (x == x ? 0.0 : x);   // 0 and denormals -> 0 (their seed is infinite); NaN stays
}

// ln x for normal finite x > 0 (no special cases, no branch)
PEAQ_MATH_FN double peaq_log_pos(double x) {
#if defined(PEAQ_LIBM_MATH)
  return log(x);
#else
  int hi = pm_hi(x);
  int e = (hi >> 20) - 1023;
  hi = (hi & 0x000fffff) | 0x3ff00000;
  const int big = hi >= 0x3ff6a09f;
  hi -= big << 20;
  e += big;
  const double m = pm_make(hi, pm_lo(x));
  const double num = m - 1.0;
  const double den = m + 1.0;
  double r = pm_rcp_seed(den);
  r = pm_fma(pm_fma(-den, r, 1.0), r, r);
  r = pm_fma(pm_fma(-den, r, 1.0), r, r);
  double f = num * r;
  f = pm_fma(pm_fma(-f, den, num), r, f);
  const double s = f * f;
  // Estrin: two half-length chains instead of one
  const double s2 = s * s;
  const double g_lo = pm_fma(pm_fma(0.14285714313751241874, s, 0.1999999999994800026), s, 0.33333333333333349075);
  const double g_hi = pm_fma(pm_fma(pm_fma(0.073115919867018450491, s, 0.076655829258400204968), s,
                                    0.090914531170014184225), s, 0.11111105446758570067);
  const double g = pm_fma(g_hi, s2 * s, g_lo);
  const double u = f + f;
  const double w = (u * s) * g;
  const double ed = (double)e;
  return pm_fma(ed, 0.6931471803691238, u) + pm_fma(ed, 1.9082149292705877e-10, w);
#endif
}

// e^x with the argument clamped to [-708, 708] (e^-708 = 3e-308 stands in for smaller results;
// NaN stays NaN); no branch
PEAQ_MATH_FN double peaq_exp_clamped(double x) {
#if defined(PEAQ_LIBM_MATH)
  return exp(x);
#else
  const double xc = fmin(fmax(x, -708.0), 708.0);
  const double magic = 6755399441055744.0;
  const double t = pm_fma(xc, 1.4426950408889634, magic);
  const int k = pm_lo(t);
  const double kd = t - magic;
  double r = pm_fma(kd, -0.6931471803691238, xc);
  r = pm_fma(kd, -1.9082149292705877e-10, r);
  // Estrin in r^2: even and odd halves of q
  const double r2 = r * r;
  const double qe = pm_fma(pm_fma(pm_fma(pm_fma(2.7620088445409748161e-7, r2, 0.000024801521295954375131), r2,
                                         0.0013888888917213716901), r2, 0.041666666666624127873), r2,
                           0.50000000000000010221);
  const double qo = pm_fma(pm_fma(pm_fma(pm_fma(2.5100385495510319078e-8, r2, 2.7557268459997064772e-6), r2,
                                         0.00019841269863053616878), r2, 0.0083333333333300618325), r2,
                           0.16666666666666667453);
  const double q = pm_fma(qo, r, qe);
  const double p = 1.0 + pm_fma(r2, q, r);
  const double res = pm_make(pm_hi(p) + (k << 20), pm_lo(p));
  return x != x ? x : res;
#endif
}

}  // namespace peaq
