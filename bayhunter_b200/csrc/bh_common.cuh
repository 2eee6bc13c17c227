// bh_common.cuh -- shared host/device helpers for the BayHunter B200 engine.
//
// Every numerical routine of the hot path is written as a BH_HD inline so the
// same source compiles (a) into the sm_100a kernels and (b) into a host-side
// lock-step simulation used only by tests/ (tests/host_sim) to check the
// control logic against the oracle without a GPU.  The host build is NOT a
// fallback: nothing in the product path calls it.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define BH_HD __host__ __device__ __forceinline__
#define BH_D __device__ __forceinline__
#else
#define BH_HD inline
#define BH_D inline
#endif

namespace bh {

// ---- IEEE round-to-nearest fp32/fp64 primitives that must never be fused ----
// (the fp32 start value and group-velocity formula of SURF96
//  have to be reproduced operation by operation.)
BH_HD float fmul(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fmul_rn(a, b);
#else
  volatile float r = a * b; return r;
#endif
}
BH_HD float fadd(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fadd_rn(a, b);
#else
  volatile float r = a + b; return r;
#endif
}
BH_HD float fsub(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fsub_rn(a, b);
#else
  volatile float r = a - b; return r;
#endif
}
BH_HD float fdiv(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fdiv_rn(a, b);
#else
  volatile float r = a / b; return r;
#endif
}
BH_HD float fsqrt(float a) {
#if defined(__CUDA_ARCH__)
  return __fsqrt_rn(a);
#else
  return sqrtf(a);
#endif
}
BH_HD double dmul(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __dmul_rn(a, b);
#else
  volatile double r = a * b; return r;
#endif
}
BH_HD double dadd(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __dadd_rn(a, b);
#else
  volatile double r = a + b; return r;
#endif
}

BH_HD double sign1(double x) { return copysign(1.0, x); }

BH_HD void sincos_d(double x, double* s, double* c) {
#if defined(__CUDA_ARCH__)
  sincos(x, s, c);
#else
  *s = sin(x); *c = cos(x);
#endif
}

struct f4 { float x, y, z, w; };  // host-side stand-in layout of float4

// ---- minimal complex<double> for the receiver-function path ----
struct cd {
  double re, im;
};
BH_HD cd mk(double re, double im) { cd r; r.re = re; r.im = im; return r; }
BH_HD cd operator+(cd a, cd b) { return mk(a.re + b.re, a.im + b.im); }
BH_HD cd operator-(cd a, cd b) { return mk(a.re - b.re, a.im - b.im); }
BH_HD cd operator-(cd a) { return mk(-a.re, -a.im); }
BH_HD cd operator*(cd a, cd b) { return mk(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re); }
BH_HD cd operator*(double s, cd a) { return mk(s * a.re, s * a.im); }
BH_HD cd operator*(cd a, double s) { return mk(s * a.re, s * a.im); }
BH_HD cd cconj(cd a) { return mk(a.re, -a.im); }
BH_HD double cnorm(cd a) { return a.re * a.re + a.im * a.im; }
BH_HD cd crecip(cd a) {
  double d = 1.0 / cnorm(a);
  return mk(a.re * d, -a.im * d);
}
BH_HD cd operator/(cd a, cd b) { return a * crecip(b); }
// principal square root (branch cut on the negative real axis, Im >= 0 for
// Im(z) = +0 exactly as std::sqrt(std::complex))
BH_HD cd csqrt_p(cd z) {
  double x = z.re, y = z.im;
  if (x == 0.0 && y == 0.0) return mk(0.0, y);
  double r = sqrt(x * x + y * y);
  double t = sqrt(0.5 * (r + fabs(x)));
  if (x >= 0.0) return mk(t, y / (2.0 * t));
  return mk(fabs(y) / (2.0 * t), copysign(t, y));
}
BH_HD cd cexp_d(cd z) {
  double s, c;
  sincos_d(z.im, &s, &c);
  double e = exp(z.re);
  return mk(e * c, e * s);
}

struct cm2 {  // complex 2x2
  cd a11, a12, a21, a22;
};
BH_HD cm2 mmul(const cm2& x, const cm2& y) {
  cm2 r;
  r.a11 = x.a11 * y.a11 + x.a12 * y.a21;
  r.a12 = x.a11 * y.a12 + x.a12 * y.a22;
  r.a21 = x.a21 * y.a11 + x.a22 * y.a21;
  r.a22 = x.a21 * y.a12 + x.a22 * y.a22;
  return r;
}

}  // namespace bh
