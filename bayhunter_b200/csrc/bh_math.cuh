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
// Host builds (tests/host_sim) map everything to libm.
#pragma once
#include "bh_common.cuh"

namespace bh {
namespace fm {

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
  double g = x * y, h = 0.5 * y;
  double r = fma(-g, h, 0.5);
  g = fma(g, r, g); h = fma(h, r, h);
  r = fma(-g, h, 0.5);
  g = fma(g, r, g); h = fma(h, r, h);
  double d = fma(-g, g, x);
  g = fma(d, h, g);
  *s = g;
  *rs = h + h;
#else
  *s = sqrt(x);
  *rs = 1.0 / *s;
#endif
}

// exp(x) for x in [-700, 0.7]
BH_HD double exp_small(double x) {
#if defined(__CUDA_ARCH__)
  const double MAGIC = 6755399441055744.0;               // 1.5 * 2^52
  double t = fma(x, 1.4426950408889634, MAGIC);          // round(x * log2 e) in the low word
  int n = __double2loint(t);
  double fn = t - MAGIC;
  double r = fma(-fn, 0.6931471805599453, x);
  r = fma(-fn, 2.3190468138462996e-17, r);                // |r| <= 0.3466
  double p = 1.6059043836821613e-10;                      // 1/13!
  p = fma(p, r, 2.08767569878681e-09);                    // 1/12!
  p = fma(p, r, 2.505210838544172e-08);                   // 1/11!
  p = fma(p, r, 2.755731922398589e-07);                   // 1/10!
  p = fma(p, r, 2.7557319223985893e-06);                  // 1/9!
  p = fma(p, r, 2.48015873015873e-05);                    // 1/8!
  p = fma(p, r, 1.984126984126984e-04);                   // 1/7!
  p = fma(p, r, 1.388888888888889e-03);                   // 1/6!
  p = fma(p, r, 8.333333333333333e-03);                   // 1/5!
  p = fma(p, r, 4.1666666666666664e-02);                  // 1/4!
  p = fma(p, r, 1.6666666666666666e-01);                  // 1/3!
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  return hi_lo(__double2hiint(p) + (int)((unsigned)n << 20), __double2loint(p));
#else
  return exp(x);
#endif
}

// sin(x), cos(x) for |x| < ~1e5: 3-term Cody-Waite reduction (exact with FMA)
// + the classic minimax kernels on [-pi/4, pi/4]
BH_HD void sincos_cw(double x, double* sn, double* cs) {
#if defined(__CUDA_ARCH__)
  const double MAGIC = 6755399441055744.0;
  double t = fma(x, 0.6366197723675814, MAGIC);
  int q = __double2loint(t);
  double fn = t - MAGIC;
  double r = fma(-fn, 1.5707963267948966, x);
  r = fma(-fn, 6.123233995736766e-17, r);
  r = fma(-fn, -1.4973849048591698e-33, r);
  double z = r * r;
  double ps = 1.58969099521155010221e-10;
  ps = fma(ps, z, -2.50507602534068634195e-08);
  ps = fma(ps, z, 2.75573137070700676789e-06);
  ps = fma(ps, z, -1.98412698298579493134e-04);
  ps = fma(ps, z, 8.33333333332248946124e-03);
  ps = fma(ps, z, -1.66666666666666324348e-01);
  double s = fma(r * z, ps, r);
  double pc = -1.13596475577881948265e-11;
  pc = fma(pc, z, 2.08757232129817482790e-09);
  pc = fma(pc, z, -2.75573143513906633035e-07);
  pc = fma(pc, z, 2.48015872894767294178e-05);
  pc = fma(pc, z, -1.38888888888741095749e-03);
  pc = fma(pc, z, 4.16666666666666019037e-02);
  double hz = 0.5 * z;
  double w = 1.0 - hz;
  double c = w + (((1.0 - w) - hz) + z * z * pc);
  double a = (q & 1) ? c : s;
  double b = (q & 1) ? s : c;
  *sn = (q & 2) ? -a : a;
  *cs = ((q + 1) & 2) ? -b : b;
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

}  // namespace fm
}  // namespace bh
