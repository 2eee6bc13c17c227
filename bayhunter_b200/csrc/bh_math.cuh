// bh_math.cuh -- straight-line fp64 elementary functions for the secular and
// reflectivity kernels.
//
// The CUDA library versions of exp / sincos / sqrt / division are correct for
// every input (NaN, denormals, |x| > 1e5, ...) and pay for it with slow-path
// subroutines and branches; ncu showed them to be ~50 % of the instructions of
// the dispersion kernel.  The arguments here are always finite, normal and of
// modest size, so these versions are branch-free DFMA sequences around one
// MUFU seed.  Accuracy: <= ~1-2 ulp (not correctly rounded); the hot path only
// needs ~1e-13 relative agreement with the reference's libm, the one place
// where exactness matters (the final +-1.0 saturation of the secular value)
// keeps an IEEE division in swd_core.cuh.
//
// All polynomial / reduction constants live in ONE __constant__ table so that
// every DFMA takes its coefficient straight from the constant bank
// (DFMA R, R, R, c[3][off]).  Written as literals, ptxas re-materialises each
// 64-bit constant with two UMOV/IMAD.MOV per use, which was ~1/3 of all
// instructions issued by the dispersion kernel (profiles/r01_swd_v1_summary.txt).
//
// Host builds (tests/host_sim) map everything to libm.
#pragma once
#include "bh_common.cuh"

namespace bh {
namespace fm {

#if defined(__CUDACC__)
enum {
  K_MAGIC = 0,      // 1.5 * 2^52
  K_LOG2E, K_LN2_HI, K_LN2_LO,
  K_Q7, K_Q6, K_Q5, K_Q4, K_Q3, K_Q2, K_Q1,
  K_TWO_OVER_PI, K_PIO2_1, K_PIO2_2, K_PIO2_3,
  K_S6, K_S5, K_S4, K_S3, K_S2, K_S1,
  K_C6, K_C5, K_C4, K_C3, K_C2, K_C1,
  K_COUNT
};
static __constant__ double kTab[K_COUNT] = {
    6755399441055744.0,
    1.4426950408889634, 0.6931471805599453, 2.3190468138462996e-17,
    // exp(r) = 1 + r + r^2 Q(r) on |r| <= ln2/2: Q = degree-9 interpolant of (e^r - 1 - r)/r^2 at the Chebyshev
    // nodes of the interval (300-bit fit; max relative error of the sum 3.9e-17 = 0.35 ulp including the
    // truncation of q9, q8 to their high words and q0 = 0.5; tests/test_math_constants.py).  Coefficients q7 .. q1:
    2.7557268378684192e-06, 2.480152119021773e-05, 1.9841269863105968e-04, 1.3888888917281794e-03,
    8.333333333330051e-03, 4.166666666662399e-02, 1.6666666666666669e-01,
    0.6366197723675814, 1.5707963267948966, 6.123233995736766e-17, -1.4973849048591698e-33,
    1.58969099521155010221e-10, -2.50507602534068634195e-08, 2.75573137070700676789e-06,
    -1.98412698298579493134e-04, 8.33333333332248946124e-03, -1.66666666666666324348e-01,
    -1.13596475577881948265e-11, 2.08757232129817482790e-09, -2.75573143513906633035e-07,
    2.48015872894767294178e-05, -1.38888888888741095749e-03, 4.16666666666666019037e-02,
};
#define BH_K(i) (::bh::fm::kTab[::bh::fm::i])
#endif

// The highest-order polynomial coefficients only need ~20 mantissa bits (their terms are
// < 1e-11 of the result), so they are written as doubles whose low word is zero: ptxas
// encodes those as 32-bit immediates of the DFMA, which keeps the table short enough
// for the remaining constants to stay in uniform registers (with all 33 constants the
// dispersion kernel spilled and refilled ~16 uniform registers per layer).
#if defined(__CUDA_ARCH__)
#define BH_KHI(hi) __hiloint2double((int)(hi), 0)
#define BH_K_Q9 BH_KHI(0x3e5af38d)    /* 2.51004e-08 */
#define BH_K_Q8 BH_KHI(0x3e92891a)    /* 2.76201e-07 */
#define BH_K_S6 BH_KHI(0x3de5d93a)
#define BH_K_C6 BH_KHI(0xbda8faea)
#endif

#if defined(__CUDA_ARCH__)
__device__ __forceinline__ double hi_lo(int hi, int lo) { return __hiloint2double(hi, lo); }
#endif

// 1/x, x finite, normal, non-zero
BH_HD double rcp(double x) {
#if defined(__CUDA_ARCH__)
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  return r;
#else
  return 1.0 / x;
#endif
}

// a/b (faithfully rounded), b as for rcp
BH_HD double div(double a, double b) {
#if defined(__CUDA_ARCH__)
  double r = rcp(b);
  double q = a * r;
  double rem = fma(-b, q, a);
  return fma(rem, r, q);
#else
  return a / b;
#endif
}

// sqrt(x) and 1/sqrt(x) together, x > 0 finite normal.  x == 0 yields NaNs that
// the callers mask with a select.
BH_HD void sqrt_rsqrt(double x, double* s, double* rs) {
#if defined(__CUDA_ARCH__)
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  // one coupled Newton step on the 2^-21 seed (-> 2^-40), then one residual correction each
  // for sqrt and 1/sqrt (sqrt <= 0.5 ulp + rounding, 1/sqrt <= 1.5 ulp)
  double g = x * y, h = 0.5 * y;
  double r = fma(-g, h, 0.5);
  g = fma(g, r, g); h = fma(h, r, h);
  double d = fma(-g, g, x);
  g = fma(d, h, g);
  h = h + h;
  r = fma(-g, h, 1.0);
  *s = g;
  *rs = fma(h, r, h);
#else
  *s = sqrt(x);
  *rs = 1.0 / *s;
#endif
}

// exp(x) for x in [-700, 0.7].  Out-of-range arguments give garbage (never a
// trap); the callers select the result away in that case.
BH_HD double exp_small(double x) {
#if defined(__CUDA_ARCH__)
  double t = fma(x, BH_K(K_LOG2E), BH_K(K_MAGIC));       // round(x * log2 e) in the low word
  int n = __double2loint(t);
  double fn = t - BH_K(K_MAGIC);
  double r = fma(-fn, BH_K(K_LN2_HI), x);
  r = fma(-fn, BH_K(K_LN2_LO), r);                       // |r| <= 0.3466
  double p = fma(BH_K_Q9, r, BH_K_Q8);                   // 1 + r + r^2 Q(r): 11 FMAs (Taylor to 1/13!: 13)
  p = fma(p, r, BH_K(K_Q7));
  p = fma(p, r, BH_K(K_Q6));
  p = fma(p, r, BH_K(K_Q5));
  p = fma(p, r, BH_K(K_Q4));
  p = fma(p, r, BH_K(K_Q3));
  p = fma(p, r, BH_K(K_Q2));
  p = fma(p, r, BH_K(K_Q1));
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  return hi_lo(__double2hiint(p) + (int)((unsigned)n << 20), __double2loint(p));
#else
  return exp(x);
#endif
}

// sin(x), cos(x) for |x| < ~1e5: 2-term Cody-Waite reduction (exact with FMA)
// + the classic minimax kernels on [-pi/4, pi/4]
BH_HD void sincos_cw(double x, double* sn, double* cs) {
#if defined(__CUDA_ARCH__)
  double t = fma(x, BH_K(K_TWO_OVER_PI), BH_K(K_MAGIC));
  int q = __double2loint(t);
  double fn = t - BH_K(K_MAGIC);
  double r = fma(-fn, BH_K(K_PIO2_1), x);
  r = fma(-fn, BH_K(K_PIO2_2), r);                       // a third term (1.5e-33 fn) is below fp64 resolution of r
  double z = r * r;
  double ps = BH_K_S6;
  ps = fma(ps, z, BH_K(K_S5));
  ps = fma(ps, z, BH_K(K_S4));
  ps = fma(ps, z, BH_K(K_S3));
  ps = fma(ps, z, BH_K(K_S2));
  ps = fma(ps, z, BH_K(K_S1));
  double s = fma(r * z, ps, r);
  double pc = BH_K_C6;
  pc = fma(pc, z, BH_K(K_C5));
  pc = fma(pc, z, BH_K(K_C4));
  pc = fma(pc, z, BH_K(K_C3));
  pc = fma(pc, z, BH_K(K_C2));
  pc = fma(pc, z, BH_K(K_C1));
  double c = fma(z, fma(z, pc, -0.5), 1.0);              // <= 0.82 ulp (the compensated form: 0.73)
  // quadrant: swap on bit 0, negate sin on bit 1, cos on bit 0 ^ bit 1 -- the
  // sign flips are integer XORs on the high words
  double a = (q & 1) ? c : s;
  double b = (q & 1) ? s : c;
  int sa = (q & 2) << 30;
  int sb = ((q + 1) & 2) << 30;
  *sn = hi_lo(__double2hiint(a) ^ sa, __double2loint(a));
  *cs = hi_lo(__double2hiint(b) ^ sb, __double2loint(b));
#else
  *sn = sin(x);
  *cs = cos(x);
#endif
}

// max(|a|, |b|) without the NaN plumbing of fmax
BH_HD double absmax(double a, double b) {
  double x = fabs(a), y = fabs(b);
  return x > y ? x : y;
}

// 2^-k with k = the binary exponent of max(|v0..v4|): an exact power-of-two
// scale that brings the largest component into [1, 2).  Integer compares on the
// high words only; zero / denormal maxima give 2^1023 (harmless: 0 stays 0).
BH_HD double pow2_rescale5(double v0, double v1, double v2, double v3, double v4) {
#if defined(__CUDA_ARCH__)
  int m0 = __double2hiint(v0) & 0x7fffffff, m1 = __double2hiint(v1) & 0x7fffffff;
  int m2 = __double2hiint(v2) & 0x7fffffff, m3 = __double2hiint(v3) & 0x7fffffff;
  int m4 = __double2hiint(v4) & 0x7fffffff;
  int m = max(max(max(m0, m1), max(m2, m3)), m4);
  return hi_lo(0x7fe00000 - (m & 0x7ff00000), 0);
#else
  double t = absmax(absmax(absmax(v0, v1), absmax(v2, v3)), v4);
  if (!(t > 0.0)) return 1.0;
  int e;
  frexp(t, &e);             // t = f * 2^e, f in [0.5, 1)
  return ldexp(1.0, 1 - e);
#endif
}
BH_HD double pow2_rescale2(double v0, double v1) {
#if defined(__CUDA_ARCH__)
  int m0 = __double2hiint(v0) & 0x7fffffff, m1 = __double2hiint(v1) & 0x7fffffff;
  int m = max(m0, m1);
  return hi_lo(0x7fe00000 - (m & 0x7ff00000), 0);
#else
  double t = absmax(v0, v1);
  if (!(t > 0.0)) return 1.0;
  int e;
  frexp(t, &e);
  return ldexp(1.0, 1 - e);
#endif
}

}  // namespace fm
}  // namespace bh
