// swd_core.cuh -- surface-wave dispersion core (Rayleigh/Love secular functions
// and the SURF96 root search) written from scratch for the B200 engine.
//
// Behavioural reference: src/extensions/surfdisp96.f of BayHunter
//   secular functions  dltar1 (:710-769), dltar4 (:773-871), var (:874-991),
//                      dnka (:1024-1068), normc (:995-1020)
//   root search        surfdisp96 period loop (:223-311), getsol (:390-482),
//                      nevill (:557-674), half (:676-686), gtsolh (:367-388)
//
// Design (not a port): the reference interleaves control flow and secular-
// function evaluations in nested loops.  Here the search is an explicit state
// machine ("Search") that only ever *requests* phase-velocity candidates and
// *consumes* secular values, so that all 32 lanes of a warp always execute the
// expensive secular function together, and so that spare lanes can evaluate the
// next bracket candidates c1+dc, c1+2dc, ... speculatively.  The sequence of
// candidates consumed is exactly the reference's, independent of how many are
// evaluated per round, hence results do not depend on the lane allocation.
#pragma once
#include "bh_common.cuh"
#include "bh_math.cuh"

namespace bh {

constexpr int SWD_MAX_PERIODS = 60;   // NP, surfdisp96.f:62
constexpr int SWD_MAX_LAYERS = 100;   // NL, surfdisp96.f:60

// One layer row of the REAL*4 model the reference sees after the f2py cast
// (surfdisp96.f:82): x = thickness d, y = vp (a), z = vs (b), w = rho.
#if defined(__CUDACC__)
typedef float4 LayerRow;
#else
typedef f4 LayerRow;
#endif

// ---------------------------------------------------------------------------
// Love secular function: Haskell 2-vector from the half-space to the surface.
// rows[l*stride], l = 0..L-1, l = L-1 is the half-space.
// ---------------------------------------------------------------------------
// ltop = llw - 1: 1 when the top layer is water (surfdisp96.f:134-135), which Love
// waves simply skip (:730).
BH_HD double secular_love_reforder(const LayerRow* rows, int stride, int L, double wvno, double omega,
                                   int ltop = 0) {
  LayerRow hs = rows[(L - 1) * stride];
  double beta1 = (double)hs.z;
  double rho1 = (double)hs.w;
  double xkb = omega / beta1;
  double rb = sqrt((wvno + xkb) * fabs(wvno - xkb));
  double e1 = rho1 * rb;
  double e2 = 1.0 / (beta1 * beta1);
  for (int l = L - 2; l >= ltop; --l) {
    LayerRow r = rows[l * stride];
    double d = (double)r.x;
    beta1 = (double)r.z;
    rho1 = (double)r.w;
    double xmu = rho1 * beta1 * beta1;
    xkb = omega / beta1;
    rb = sqrt((wvno + xkb) * fabs(wvno - xkb));
    double q = d * rb;
    double cosq, y, z;
    if (wvno < xkb) {
      double sinq;
      sincos_d(q, &sinq, &cosq);
      y = sinq / rb;
      z = -rb * sinq;
    } else if (wvno == xkb) {
      cosq = 1.0;
      y = d;
      z = 0.0;
    } else {
      double fac = (q < 16.0) ? exp(-2.0 * q) : 0.0;
      cosq = (1.0 + fac) * 0.5;
      double sinq = (1.0 - fac) * 0.5;
      y = sinq / rb;
      z = rb * sinq;
    }
    double e10 = e1 * cosq + e2 * xmu * z;
    double e20 = e1 * y / xmu + e2 * cosq;
    double xnor = fmax(fabs(e10), fabs(e20));
    if (xnor < 1.0e-40) xnor = 1.0;
    // true divisions: when |e10| is the maximum the reference gets exactly
    // +-1.0, and nevill's sign/ratio tests (:619,:628) sit on that tie.
    e1 = e10 / xnor;
    e2 = e20 / xnor;
  }
  return e1;
}

// ---------------------------------------------------------------------------
// Rayleigh secular function: Dunkin 5-component compound vector propagated
// from the half-space up; per layer the eigenfunction products of `var` and
// the compound matrix of `dnka` are formed in registers.
// ---------------------------------------------------------------------------
// ltop = 1: the top layer is water and enters through the fluid boundary condition (:850-867).
BH_HD double secular_rayleigh_reforder(const LayerRow* rows, int stride, int L, double wvno, double omga,
                                       int ltop = 0) {
  double omega = (omga < 1.0e-4) ? 1.0e-4 : omga;
  double wvno2 = wvno * wvno;
  double e0, e1, e2, e3, e4;
  {
    LayerRow hs = rows[(L - 1) * stride];
    double a = (double)hs.y, b = (double)hs.z, rho1 = (double)hs.w;
    double xka = omega / a, xkb = omega / b;
    double ra = sqrt((wvno + xka) * fabs(wvno - xka));
    double rb = sqrt((wvno + xkb) * fabs(wvno - xkb));
    double t = b / omega;
    double gammk = 2.0 * t * t;
    double gam = gammk * wvno2;
    double gamm1 = gam - 1.0;
    e0 = rho1 * rho1 * (gamm1 * gamm1 - gam * gammk * ra * rb);
    e1 = -rho1 * ra;
    e2 = rho1 * (gamm1 - gammk * ra * rb);
    e3 = rho1 * rb;
    e4 = wvno2 - ra * rb;
  }
  for (int l = L - 2; l >= ltop; --l) {
    LayerRow r = rows[l * stride];
    double dpth = (double)r.x, a = (double)r.y, b = (double)r.z, rho = (double)r.w;
    double xka = omega / a, xkb = omega / b;
    double t = b / omega;
    double gammk = 2.0 * t * t;
    double gam = gammk * wvno2;
    double ra = sqrt((wvno + xka) * fabs(wvno - xka));
    double rb = sqrt((wvno + xkb) * fabs(wvno - xkb));
    double p = ra * dpth, q = rb * dpth;
    // --- var (:921-990) ---
    double pex = 0.0, sex = 0.0;
    double cosp, w, x, cosq, y, z;
    if (wvno < xka) {
      double sinp;
      sincos_d(p, &sinp, &cosp);
      w = sinp / ra;
      x = -ra * sinp;
    } else if (wvno == xka) {
      cosp = 1.0; w = dpth; x = 0.0;
    } else {
      pex = p;
      double fac = (p < 16.0) ? exp(-2.0 * p) : 0.0;
      cosp = (1.0 + fac) * 0.5;
      double sinp = (1.0 - fac) * 0.5;
      w = sinp / ra;
      x = ra * sinp;
    }
    if (wvno < xkb) {
      double sinq;
      sincos_d(q, &sinq, &cosq);
      y = sinq / rb;
      z = -rb * sinq;
    } else if (wvno == xkb) {
      cosq = 1.0; y = dpth; z = 0.0;
    } else {
      sex = q;
      double fac = (q < 16.0) ? exp(-2.0 * q) : 0.0;
      cosq = (1.0 + fac) * 0.5;
      double sinq = (1.0 - fac) * 0.5;
      y = sinq / rb;
      z = rb * sinq;
    }
    double exa = pex + sex;
    double a0 = (exa < 60.0) ? exp(-exa) : 0.0;
    double cpcq = cosp * cosq, cpy = cosp * y, cpz = cosp * z, cqw = cosq * w, cqx = cosq * x;
    double xy = x * y, xz = x * z, wy = w * y, wz = w * z;
    // --- dnka (:1032-1067) ---
    double gamm1 = gam - 1.0;
    double twgm1 = gam + gamm1;
    double gmgmk = gam * gammk;
    double gmgm1 = gam * gamm1;
    double gm1sq = gamm1 * gamm1;
    double rho2 = rho * rho;
    double a0pq = a0 - cpcq;
    double c11 = cpcq - 2.0 * gmgm1 * a0pq - gmgmk * xz - wvno2 * gm1sq * wy;
    double c12 = (wvno2 * cpy - cqx) / rho;
    double c13 = -(twgm1 * a0pq + gammk * xz + wvno2 * gamm1 * wy) / rho;
    double c14 = (cpz - wvno2 * cqw) / rho;
    double c15 = -(2.0 * wvno2 * a0pq + xz + wvno2 * wvno2 * wy) / rho2;
    double c21 = (gmgmk * cpz - gm1sq * cqw) * rho;
    double c22 = cpcq;
    double c23 = gammk * cpz - gamm1 * cqw;
    double c24 = -wz;
    double c41 = (gm1sq * cpy - gmgmk * cqx) * rho;
    double c42 = -xy;
    double c43 = gamm1 * cpy - gammk * cqx;
    double c51 = -(2.0 * gmgmk * gm1sq * a0pq + gmgmk * gmgmk * xz + gm1sq * gm1sq * wy) * rho2;
    double c53 = -(gammk * gamm1 * twgm1 * a0pq + gam * gammk * gammk * xz + gamm1 * gm1sq * wy) * rho;
    double tt = -2.0 * wvno2;
    double c31 = tt * c53, c32 = tt * c43, c33 = a0 + 2.0 * (cpcq - c11), c34 = tt * c23, c35 = tt * c13;
    // remaining entries by symmetry (:1048-1060): c25=c14, c44=c22, c45=c12,
    // c52=c41, c54=c21, c55=c11
    // --- ee = e * ca (:838-844) ---
    double n0 = e0 * c11 + e1 * c21 + e2 * c31 + e3 * c41 + e4 * c51;
    double n1 = e0 * c12 + e1 * c22 + e2 * c32 + e3 * c42 + e4 * c41;
    double n2 = e0 * c13 + e1 * c23 + e2 * c33 + e3 * c43 + e4 * c53;
    double n3 = e0 * c14 + e1 * c24 + e2 * c34 + e3 * c22 + e4 * c21;
    double n4 = e0 * c15 + e1 * c14 + e2 * c35 + e3 * c12 + e4 * c11;
    // --- normc (:1004-1014); the stored log is never read by the caller ---
    double t1 = fmax(fmax(fmax(fabs(n0), fabs(n1)), fmax(fabs(n2), fabs(n3))), fabs(n4));
    if (t1 < 1.0e-40) t1 = 1.0;
    // e0 is what the caller finally reads: keep the reference's true division so
    // that a saturated component is exactly +-1.0 (see secular_love).
    e0 = n0 / t1; e1 = n1 / t1; e2 = n2 / t1; e3 = n3 / t1; e4 = n4 / t1;
  }
  if (ltop != 0) {
    // water layer on top (:850-867): only the P terms w, cosp of `var` are read
    LayerRow r = rows[0];
    double dpth = (double)r.x, rho1 = (double)r.w;
    double xka = omega / (double)r.y;
    double ra = sqrt((wvno + xka) * fabs(wvno - xka));
    double p = ra * dpth;
    double w, cosp;
    if (wvno < xka) {
      double sinp;
      sincos_d(p, &sinp, &cosp);
      w = sinp / ra;
    } else if (wvno == xka) {
      cosp = 1.0; w = dpth;
    } else {
      double fac = (p < 16.0) ? exp(-2.0 * p) : 0.0;
      cosp = (1.0 + fac) * 0.5;
      w = ((1.0 - fac) * 0.5) / ra;
    }
    return cosp * e0 + (-rho1 * w) * e1;
  }
  return e0;
}

// ---------------------------------------------------------------------------
// Device formulation of the same two functions: identical algebra, but
//  * per-layer constants that do not depend on the trial phase velocity
//    (1/vp, 1/vs, 1/rho, 2 vs^2, rho vs^2 ...) are derived ONCE per search into
//    fp64 "records" (swd_make_rec) that the kernel keeps in shared memory,
//    field-major so that 32 lanes reading the same field of 32 different
//    models hit 32 different banks,
//  * branch-free per layer (the oscillatory / evanescent cases are selects; a
//    warp mixing them paid for both sides anyway) so that the P and S halves
//    and the matrix algebra of a layer form one long basic block with
//    instruction-level parallelism for the fp64 pipe; the grazing case
//    k == omega/v is folded into the oscillatory one by flooring the radicand,
//  * exp(-2p) = exp(-p)^2 and exp(-(p+q)) = exp(-p) exp(-q): two exponentials
//    per layer instead of three,
//  * reciprocals and rsqrt seeds instead of IEEE divisions / square roots
//    (bh_math.cuh),
//  * the per-layer renormalisation of the propagated vector multiplies by an
//    exact power of two (integer compare of the exponent fields) instead of
//    dividing by the largest component: the secular value is invariant under
//    that scale, and the serial vector chain loses its longest-latency link.
//    Only the LAST normalisation is the reference's: its IEEE division keeps a
//    saturated secular value at exactly +-1.0 (nevill's sign/ratio tests sit
//    on that tie, surfdisp96.f:619,628).
// ---------------------------------------------------------------------------
// 1: layers in groups (4 Love layers / 2 Rayleigh layers side by side, more ILP, more code);
// 0: one layer at a time (smallest instruction-cache footprint)
#ifndef BH_SWD_WIDE
#define BH_SWD_WIDE 0
#endif
// 1: the P half of a Rayleigh layer takes its sincos only when some lane of the warp is oscillatory there
#ifndef BH_P_SINCOS_COND
#define BH_P_SINCOS_COND 1
#endif
// Love layers whose (vector-independent) terms are formed side by side: 1, 2 or 4
#ifndef BH_LOVE_GROUP
#define BH_LOVE_GROUP (BH_SWD_WIDE ? 4 : 1)
#endif
// 1: the compares of non-negative doubles against constants with a zero low word (floor of the
// radicand, p < 16, p + q < 60) and the sign test k <= xk are integer compares on the high
// words -- they leave the fp64 pipe, which bounds the layer loops (2 issue cycles per warp
// instruction); the square root takes ONE coupled Newton step after the MUFU seed (2^-21)
// and finishes sqrt and 1/sqrt with one residual correction each (10 fp64 instructions, was 13).
#ifndef BH_SWD_LEAN
#define BH_SWD_LEAN 1
#endif
// 1: e * C(layer) through the rank-one structure of Dunkin's compound matrix (dunkin_apply_factored:
// ~70 fp64 instructions per layer instead of ~100 for forming the 19 distinct entries and the 5x5 product)
#ifndef BH_DUNKIN_FACTORED
#define BH_DUNKIN_FACTORED 1
#endif
constexpr int SWD_REC_FIELDS = 6;
// Rayleigh record fields
enum { RR_D = 0, RR_IA = 1, RR_IB = 2, RR_RHO = 3, RR_IRHO = 4, RR_TB2 = 5 };
// Love record fields (half-space row: LR_D holds rho)
enum { LR_D = 0, LR_IB = 1, LR_MU = 2, LR_IMU = 3 };

// out[f * fs], f < SWD_REC_FIELDS
BH_HD void swd_make_rec(int wave, const LayerRow& r, bool halfspace, double* out, int fs) {
  double d = (double)r.x, a = (double)r.y, b = (double)r.z, rho = (double)r.w;
  if (wave == 1) {
    double mu = rho * b * b;
    out[LR_D * fs] = halfspace ? rho : d;
    out[LR_IB * fs] = fm::rcp(b);
    out[LR_MU * fs] = mu;
    out[LR_IMU * fs] = fm::rcp(mu);
    out[4 * fs] = 0.0;
    out[5 * fs] = 0.0;
  } else {
    out[RR_D * fs] = d;
    out[RR_IA * fs] = fm::rcp(a);
    out[RR_IB * fs] = fm::rcp(b);
    out[RR_RHO * fs] = rho;
    out[RR_IRHO * fs] = fm::rcp(rho);
    out[RR_TB2 * fs] = 2.0 * b * b;
  }
}

struct HalfTerms { double cs, sn_over_r, r_sn, ex, em, r; };   // cos-like, sin/r, +-r*sin, exponent, exp(-ex), r = sqrt of the (floored) radicand

#define BH_N(...) _Pragma("unroll") for (int i = 0; i < N; ++i) { __VA_ARGS__; }

// `var` for N (wave type, layer) pairs at once (surfdisp96.f:929-968): k = wvno,
// xk[i] = omega/v_i, s = (k+xk)|k-xk|, d[i] = thickness.  Per pair: cosp,
// w = sinp/ra, x = -+ra*sinp, the evanescent exponent pex (0 unless k > xk) and
// exp(-pex).
//
// The device version is written "N wide": every step of the rsqrt / exp / sincos
// sequences is applied to all N pairs before the next step, so that N (and, where
// exp and sincos run side by side, 2N-3N) independent DFMA chains are adjacent in
// program order.  One warp then keeps the fp64 pipe busy on its own (measured:
// DFMA latency 8 cycles, issue 2 cycles -> 4 independent chains saturate it),
// which matters because the searches leave only 1-3 warps per SM sub-partition.
// kCond: bit i set = item i takes its sincos only when some lane of the warp is oscillatory there
// (the P half of a Rayleigh layer: c > vp of that layer is rare) -- a warp-uniform branch.
template <int N, unsigned kCond = (BH_P_SINCOS_COND && N == 2) ? 1u : 0u>
BH_HD void half_terms_nk(const double* kk, const double* xk, const double* d, HalfTerms* h);
template <int N, unsigned kCond = (BH_P_SINCOS_COND && N == 2) ? 1u : 0u>
BH_HD void half_terms_n(double k, const double* xk, const double* d, HalfTerms* h) {
  double kk[N];
#pragma unroll
  for (int i = 0; i < N; ++i) kk[i] = k;
  half_terms_nk<N, kCond>(kk, xk, d, h);
}
// the same with a wavenumber per item (two trial velocities side by side)
template <int N, unsigned kCond>
BH_HD void half_terms_nk(const double* kk, const double* xk, const double* d, HalfTerms* h) {
#if defined(__CUDA_ARCH__)
  double s[N], y[N], g[N], hh[N], r[N], p[N], pm[N];
  bool osc[N];
  // grazing (k == xk, reference: cosp = 1, w = d, x = 0): a floored radicand on
  // the oscillatory side gives cos(1e-100 d) = 1, sin(p)/r = d, r sin(p) = 1e-200 d
#if BH_SWD_LEAN
  // s >= 0: doubles order like their high words; k - xk is exactly +0 or at least one ulp of k
  BH_N(double dk = kk[i] - xk[i]; s[i] = (kk[i] + xk[i]) * fabs(dk); osc[i] = __double2hiint(dk) <= 0)
  BH_N(s[i] = fm::hi_lo(max(__double2hiint(s[i]), 0x16687e92), __double2loint(s[i])))      // hi word of 1e-200
  BH_N(asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y[i]) : "d"(s[i])))
  // halving / doubling of normal numbers: exponent field arithmetic on the high word
  BH_N(g[i] = s[i] * y[i]; hh[i] = fm::hi_lo(__double2hiint(y[i]) - 0x00100000, __double2loint(y[i])))
  BH_N(r[i] = fma(-g[i], hh[i], 0.5))
  BH_N(g[i] = fma(g[i], r[i], g[i]); hh[i] = fma(hh[i], r[i], hh[i]))       // ~2^-40
  BH_N(r[i] = fma(-g[i], g[i], s[i]))
  BH_N(g[i] = fma(r[i], hh[i], g[i]))                                      // g = sqrt(s)
  BH_N(hh[i] = fm::hi_lo(__double2hiint(hh[i]) + 0x00100000, __double2loint(hh[i])))
  BH_N(r[i] = fma(-g[i], hh[i], 1.0))                  // g is final: the residual of g * hh is the error of hh alone
  BH_N(hh[i] = fma(hh[i], r[i], hh[i]))                                    // hh = 1/sqrt(s)
#else
  BH_N(s[i] = (kk[i] + xk[i]) * fabs(kk[i] - xk[i]))
  BH_N(s[i] = (s[i] < 1.0e-200) ? 1.0e-200 : s[i]; osc[i] = kk[i] <= xk[i])
  BH_N(asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y[i]) : "d"(s[i])))
  BH_N(g[i] = s[i] * y[i]; hh[i] = 0.5 * y[i])
  BH_N(r[i] = fma(-g[i], hh[i], 0.5))
  BH_N(g[i] = fma(g[i], r[i], g[i]); hh[i] = fma(hh[i], r[i], hh[i]))
  BH_N(r[i] = fma(-g[i], hh[i], 0.5))
  BH_N(g[i] = fma(g[i], r[i], g[i]); hh[i] = fma(hh[i], r[i], hh[i]))
  BH_N(r[i] = fma(-g[i], g[i], s[i]))
  BH_N(g[i] = fma(r[i], hh[i], g[i]); hh[i] = hh[i] + hh[i])      // g = sqrt(s), hh = 1/sqrt(s)
#endif
  BH_N(p[i] = g[i] * d[i]; pm[i] = osc[i] ? 0.0 : p[i])
  // exp(-pm) and sincos(p), step by step side by side.  With BH_P_SINCOS_COND the P half of a
  // Rayleigh layer (i = 0 of N = 2) takes its sincos only when some lane of the warp has an
  // oscillatory P term (c > vp of that layer: rare) -- a warp-uniform branch.
  double te[N], ts[N], fe[N], ft[N], re[N], rt[N], z[N], pe[N], ps[N], pc[N];
  int ne[N], q[N];
  double sn[N], cn[N];
#define BH_SC(...) if (!((kCond >> i) & 1u)) { __VA_ARGS__; }
  BH_N(te[i] = fma(-pm[i], BH_K(K_LOG2E), BH_K(K_MAGIC)); BH_SC(ts[i] = fma(p[i], BH_K(K_TWO_OVER_PI), BH_K(K_MAGIC))))
  BH_N(ne[i] = __double2loint(te[i]); fe[i] = te[i] - BH_K(K_MAGIC);
       BH_SC(q[i] = __double2loint(ts[i]); ft[i] = ts[i] - BH_K(K_MAGIC)))
  BH_N(re[i] = fma(-fe[i], BH_K(K_LN2_HI), -pm[i]); BH_SC(rt[i] = fma(-ft[i], BH_K(K_PIO2_1), p[i])))
  BH_N(re[i] = fma(-fe[i], BH_K(K_LN2_LO), re[i]); BH_SC(rt[i] = fma(-ft[i], BH_K(K_PIO2_2), rt[i])))
  // exp(r) = 1 + r + r^2 Q(r), Q of degree 9 (bh_math.cuh: exp_small)
  BH_N(pe[i] = fma(BH_K_Q9, re[i], BH_K_Q8); BH_SC(z[i] = rt[i] * rt[i]))
  BH_N(pe[i] = fma(pe[i], re[i], BH_K(K_Q7)); BH_SC(ps[i] = fma(BH_K_S6, z[i], BH_K(K_S5)); pc[i] = fma(BH_K_C6, z[i], BH_K(K_C5))))
  BH_N(pe[i] = fma(pe[i], re[i], BH_K(K_Q6)); BH_SC(ps[i] = fma(ps[i], z[i], BH_K(K_S4)); pc[i] = fma(pc[i], z[i], BH_K(K_C4))))
  BH_N(pe[i] = fma(pe[i], re[i], BH_K(K_Q5)); BH_SC(ps[i] = fma(ps[i], z[i], BH_K(K_S3)); pc[i] = fma(pc[i], z[i], BH_K(K_C3))))
  BH_N(pe[i] = fma(pe[i], re[i], BH_K(K_Q4)); BH_SC(ps[i] = fma(ps[i], z[i], BH_K(K_S2)); pc[i] = fma(pc[i], z[i], BH_K(K_C2))))
  BH_N(pe[i] = fma(pe[i], re[i], BH_K(K_Q3)); BH_SC(ps[i] = fma(ps[i], z[i], BH_K(K_S1)); pc[i] = fma(pc[i], z[i], BH_K(K_C1))))
  BH_N(pe[i] = fma(pe[i], re[i], BH_K(K_Q2)); BH_SC(sn[i] = fma(rt[i] * z[i], ps[i], rt[i]); pc[i] = fma(z[i], pc[i], -0.5)))
  BH_N(pe[i] = fma(pe[i], re[i], BH_K(K_Q1)); BH_SC(cn[i] = fma(z[i], pc[i], 1.0)))
  BH_N(pe[i] = fma(pe[i], re[i], 0.5))
  BH_N(pe[i] = fma(pe[i], re[i], 1.0))
  BH_N(pe[i] = fma(pe[i], re[i], 1.0))
#undef BH_SC
  if (kCond != 0u) {
    bool any = false;
    BH_N(if ((kCond >> i) & 1u) { q[i] = 0; sn[i] = 0.0; cn[i] = 1.0; any = any || osc[i]; })
    if (__any_sync(__activemask(), any)) {
      BH_N(if ((kCond >> i) & 1u) {
        double a, c;
        fm::sincos_cw(p[i], &a, &c);      // the same sequence as above (bh_math.cuh), quadrant already applied
        sn[i] = a; cn[i] = c; })
    }
  }
  double em[N], fac[N], ch[N], sh[N];
  BH_N(em[i] = fm::hi_lo(__double2hiint(pe[i]) + (int)((unsigned)ne[i] << 20), __double2loint(pe[i])))
  // quadrant fix-up of sin/cos: swap on bit 0, sign flips as XORs on the high words
  BH_N(double a = (q[i] & 1) ? cn[i] : sn[i]; double b = (q[i] & 1) ? sn[i] : cn[i];
       sn[i] = fm::hi_lo(__double2hiint(a) ^ ((q[i] & 2) << 30), __double2loint(a));
       cn[i] = fm::hi_lo(__double2hiint(b) ^ (((q[i] + 1) & 2) << 30), __double2loint(b)))
#if BH_SWD_LEAN
  BH_N(fac[i] = (__double2hiint(pm[i]) < 0x40300000) ? em[i] * em[i] : 0.0)      // pm >= 0: pm < 16
#else
  BH_N(fac[i] = (pm[i] < 16.0) ? em[i] * em[i] : 0.0)
#endif
  BH_N(ch[i] = fma(fac[i], 0.5, 0.5); sh[i] = fma(fac[i], -0.5, 0.5))
  BH_N(h[i].cs = osc[i] ? cn[i] : ch[i]; sh[i] = osc[i] ? sn[i] : sh[i])
#if BH_SWD_LEAN
  BH_N(h[i].sn_over_r = sh[i] * hh[i]; double rs = g[i] * sh[i];
       h[i].r_sn = fm::hi_lo(__double2hiint(rs) ^ (osc[i] ? (int)0x80000000 : 0), __double2loint(rs));
       h[i].ex = pm[i]; h[i].em = em[i]; h[i].r = g[i])
#else
  BH_N(h[i].sn_over_r = sh[i] * hh[i]; double rs = g[i] * sh[i]; h[i].r_sn = osc[i] ? -rs : rs;
       h[i].ex = pm[i]; h[i].em = em[i]; h[i].r = g[i])
#endif
#else
  for (int i = 0; i < N; ++i) {
    const double k = kk[i];
    double s = (k + xk[i]) * fabs(k - xk[i]);
    s = (s < 1.0e-200) ? 1.0e-200 : s;
    const bool osc = k <= xk[i];
    double r = sqrt(s), ir = 1.0 / r;
    double p = r * d[i];
    double pm = osc ? 0.0 : p;
    double em = exp(-pm);
    double sn = sin(p), cs = cos(p);
    double fac = (pm < 16.0) ? em * em : 0.0;
    double ch = fac * 0.5 + 0.5, sh = fac * -0.5 + 0.5;
    h[i].cs = osc ? cs : ch;
    double sx = osc ? sn : sh;
    h[i].sn_over_r = sx * ir;
    double rs = r * sx;
    h[i].r_sn = osc ? -rs : rs;
    h[i].ex = pm;
    h[i].em = em;
    h[i].r = r;
  }
#endif
}

// sqrt of N radicands (half-space terms), same rsqrt sequence, floored like above
template <int N>
BH_HD void sqrt_n(const double* sIn, double* out) {
#if defined(__CUDA_ARCH__)
  double s[N], y[N], g[N], hh[N], r[N];
  BH_N(s[i] = (sIn[i] < 1.0e-200) ? 1.0e-200 : sIn[i])
  BH_N(asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y[i]) : "d"(s[i])))
  BH_N(g[i] = s[i] * y[i]; hh[i] = 0.5 * y[i])
  BH_N(r[i] = fma(-g[i], hh[i], 0.5))
  BH_N(g[i] = fma(g[i], r[i], g[i]); hh[i] = fma(hh[i], r[i], hh[i]))
#if !BH_SWD_LEAN
  BH_N(r[i] = fma(-g[i], hh[i], 0.5))
  BH_N(g[i] = fma(g[i], r[i], g[i]); hh[i] = fma(hh[i], r[i], hh[i]))
#endif
  BH_N(r[i] = fma(-g[i], g[i], s[i]))
  BH_N(out[i] = (sIn[i] < 1.0e-200) ? 0.0 : fma(r[i], hh[i], g[i]))
#else
  for (int i = 0; i < N; ++i) out[i] = sqrt(sIn[i]);
#endif
}

// Love: per-layer terms that do not depend on the propagated vector
struct LoveLayer { double cs, y_over_mu, mu_z; };

// N consecutive layers l0, l0-1, ..., l0-N+1 at once (rec points at field 0, layer 0)
template <int N>
BH_HD void love_layers_n(const double* rec, int fs, int ls, int l0, double wvno, double omega, LoveLayer* m) {
  double xk[N], d[N];
  HalfTerms q[N];
  BH_N(const double* r = rec + (l0 - i) * ls; xk[i] = omega * r[LR_IB * fs]; d[i] = r[LR_D * fs])
  half_terms_n<N>(wvno, xk, d, q);
  BH_N(const double* r = rec + (l0 - i) * ls; m[i].cs = q[i].cs; m[i].y_over_mu = q[i].sn_over_r * r[LR_IMU * fs];
       m[i].mu_z = r[LR_MU * fs] * q[i].r_sn)
}

BH_HD void love_apply(const LoveLayer& m, double& e1, double& e2) {
  double e10 = e1 * m.cs + e2 * m.mu_z;
  double e20 = e1 * m.y_over_mu + e2 * m.cs;
  double sc = fm::pow2_rescale2(e10, e20);
  e1 = e10 * sc;
  e2 = e20 * sc;
}

// rec: field f of layer l at rec[f * fs + l * ls]; l = L-1 is the half-space.
// Layers are taken four at a time (their terms are independent of the propagated
// vector), then applied in turn; every layer ends with the exact power-of-two
// rescale, and the reference's normalisation e1 / max(|e1|, |e2|) of the top layer
// is taken once at the end -- the quotient is bit-identical because both operands
// carry the same power of two.  Uniform counted loops, each code block once
// (the kernel is sensitive to its instruction-cache footprint).
BH_HD double secular_love_rec(const double* rec, int fs, int ls, int L, double wvno, double omega) {
  const double* hs = rec + (L - 1) * ls;
  double ib = hs[LR_IB * fs];
  double xkb = omega * ib;
  double srb = (wvno + xkb) * fabs(wvno - xkb), rb;
  sqrt_n<1>(&srb, &rb);
  double e1 = hs[LR_D * fs] * rb;     // rho * rb
  double e2 = ib * ib;
  if (L < 2) return e1;
  int l = L - 2;
#pragma unroll 1
  for (int r = (BH_LOVE_GROUP > 1) ? ((L - 1) % BH_LOVE_GROUP) : (L - 1); r > 0; --r, --l) {
    LoveLayer m;
    love_layers_n<1>(rec, fs, ls, l, wvno, omega, &m);
    love_apply(m, e1, e2);
  }
#if BH_LOVE_GROUP > 1
#pragma unroll 1
  for (int g = (L - 1) / BH_LOVE_GROUP; g > 0; --g, l -= BH_LOVE_GROUP) {     // layers l .. l-G+1
    LoveLayer m[BH_LOVE_GROUP];
    love_layers_n<BH_LOVE_GROUP>(rec, fs, ls, l, wvno, omega, m);
#pragma unroll
    for (int j = 0; j < BH_LOVE_GROUP; ++j) love_apply(m[j], e1, e2);
  }
#endif
  double xnor = fm::absmax(e1, e2);
  if (xnor < 1.0e-40) xnor = 1.0;
  return e1 / xnor;                          // IEEE division: exact +-1.0 when saturated
}

// Rayleigh: the distinct entries of Dunkin's compound matrix of one layer
struct DunkinLayer {
  double c11, c12, c13, c14, c15, c21, c22, c23, c24, c31, c32, c33, c34, c35, c41, c42, c43, c51, c53;
};

// var's products + dnka (:969-983, :1032-1067) from the P and S terms of a layer
BH_HD DunkinLayer dunkin_from_terms(const double* rec, int fs, const HalfTerms& P, const HalfTerms& S,
                                    double wvno2, double iomega2) {
  double rho = rec[RR_RHO * fs], rinv = rec[RR_IRHO * fs];
  double gammk = rec[RR_TB2 * fs] * iomega2;          // 2 (b/omega)^2
  double gam = gammk * wvno2;
  double cosp = P.cs, w = P.sn_over_r, x = P.r_sn;
  double cosq = S.cs, y = S.sn_over_r, z = S.r_sn;
  double exa = P.ex + S.ex;
#if BH_SWD_LEAN && defined(__CUDA_ARCH__)
  double a0 = (__double2hiint(exa) < 0x404e0000) ? P.em * S.em : 0.0;     // exa >= 0: exa < 60
#else
  double a0 = (exa < 60.0) ? P.em * S.em : 0.0;
#endif
  double cpcq = cosp * cosq, cpy = cosp * y, cpz = cosp * z, cqw = cosq * w, cqx = cosq * x;
  double xy = x * y, xz = x * z, wy = w * y, wz = w * z;
  double gamm1 = gam - 1.0;
  double twgm1 = gam + gamm1;
  double gmgmk = gam * gammk;
  double gmgm1 = gam * gamm1;
  double gm1sq = gamm1 * gamm1;
  double rho2 = rho * rho;
  double a0pq = a0 - cpcq;
  DunkinLayer m;
  m.c11 = cpcq - 2.0 * gmgm1 * a0pq - gmgmk * xz - wvno2 * gm1sq * wy;
  m.c12 = (wvno2 * cpy - cqx) * rinv;
  m.c13 = -(twgm1 * a0pq + gammk * xz + wvno2 * gamm1 * wy) * rinv;
  m.c14 = (cpz - wvno2 * cqw) * rinv;
  m.c15 = -(2.0 * wvno2 * a0pq + xz + wvno2 * wvno2 * wy) * (rinv * rinv);
  m.c21 = (gmgmk * cpz - gm1sq * cqw) * rho;
  m.c22 = cpcq;
  m.c23 = gammk * cpz - gamm1 * cqw;
  m.c24 = -wz;
  m.c41 = (gm1sq * cpy - gmgmk * cqx) * rho;
  m.c42 = -xy;
  m.c43 = gamm1 * cpy - gammk * cqx;
  m.c51 = -(2.0 * gmgmk * gm1sq * a0pq + gmgmk * gmgmk * xz + gm1sq * gm1sq * wy) * rho2;
  m.c53 = -(gammk * gamm1 * twgm1 * a0pq + gam * gammk * gammk * xz + gamm1 * gm1sq * wy) * rho;
  double tt = -2.0 * wvno2;
  m.c31 = tt * m.c53; m.c32 = tt * m.c43; m.c33 = a0 + 2.0 * (cpcq - m.c11);
  m.c34 = tt * m.c23; m.c35 = tt * m.c13;
  return m;
}

// e <- e * C(layer) without forming C.  With A = a0 - cosp cosq, X = x z, W = w y the block of rows /
// columns (1, 3, 5) of Dunkin's matrix (:1032-1067) is
//     cosp cosq * I + A uA vA^T + X uX vX^T + W uW vW^T
//     uA = (1, -(2 gam - 1) rho, gammk gamm1 rho^2)     vA = -(2 gam gamm1, (2 gam - 1) / rho, 2 k^2 / rho^2)
//     uX = (1, -2 gam rho, gam gammk rho^2)             vX = -(gam gammk, gammk / rho, 1 / rho^2)
//     uW = (k^2, -2 k^2 gamm1 rho, gamm1^2 rho^2)       vW = -(gamm1^2, gamm1 / rho, k^2 / rho^2)
// (gam = gammk k^2), and the entries that couple them with components 2 and 4 are built from the same vectors:
//     rows 2, 4 -> columns (1, 3, 5):  -cosp z rho vX + cosq w rho vW,   -cosp y rho vW + cosq x rho vX
//     rows (1, 3, 5) -> columns 2, 4:  (cosp y uW - cosq x uX) / rho,    (cosp z uX - cosq w uW) / rho
// so the product needs three dot products e.u, three scaled sums and two short rows.  Same algebra as
// dunkin_from_terms + BH_DUNKIN_APPLY, re-associated (values agree to a few ulp of the vector norm).
// a0 = exp(-(pex + sex)) of var (:969-971) from the two half terms
BH_HD double dunkin_a0(const HalfTerms& P, const HalfTerms& S) {
  const double exa = P.ex + S.ex;
#if BH_SWD_LEAN && defined(__CUDA_ARCH__)
  return (__double2hiint(exa) < 0x404e0000) ? P.em * S.em : 0.0;     // exa >= 0: exa < 60
#else
  return (exa < 60.0) ? P.em * S.em : 0.0;
#endif
}
// The vector-independent terms of a layer as they cross from the half-term pass to the propagation pass
struct DunkinTerms { double cosp, w, x, cosq, y, z, a0; };
BH_HD void dunkin_apply_terms(double rho, double ri, double gk, const DunkinTerms& t, double k2,
                              double& e0, double& e1, double& e2, double& e3, double& e4) {
  const double gam = gk * k2;
  const double gm1 = gam - 1.0;
  const double tw = gam + gm1;
  const double cosp = t.cosp, w = t.w, x = t.x;
  const double cosq = t.cosq, y = t.y, z = t.z;
  const double a0 = t.a0;
  const double cpcq = cosp * cosq, cpy = cosp * y, cpz = cosp * z, cqw = cosq * w, cqx = cosq * x;
  const double A = a0 - cpcq, X = x * z, W = w * y;
  const double f3 = e2 * rho, f5 = e4 * (rho * rho);
  const double ggm1 = gk * gm1, gmgk = gam * gk, gm1sq = gm1 * gm1, k2g1 = k2 * gm1;
  const double dA = fma(ggm1, f5, fma(-tw, f3, e0));
  const double dX = fma(gmgk, f5, fma(-2.0 * gam, f3, e0));
  const double dW = fma(gm1sq, f5, fma(-2.0 * k2g1, f3, k2 * e0));
  const double sA = A * dA;
  const double sX = fma(-rho, fma(cpz, e1, -(cqx * e3)), X * dX);
  const double sW = fma(rho, fma(cqw, e1, -(cpy * e3)), W * dW);
  const double t1 = fma(gm1sq, sW, fma(gmgk, sX, (2.0 * gam * gm1) * sA));
  const double t3 = fma(gm1, sW, fma(gk, sX, tw * sA));
  const double t5 = fma(k2, sW, fma(2.0 * k2, sA, sX));
  const double n0 = fma(cpcq, e0, -t1);
  const double n2 = fma(-ri, t3, cpcq * e2);
  const double n4 = fma(-(ri * ri), t5, cpcq * e4);
  const double n1 = fma(ri, fma(cpy, dW, -(cqx * dX)), fma(cpcq, e1, -((x * y) * e3)));
  const double n3 = fma(ri, fma(cpz, dX, -(cqw * dW)), fma(cpcq, e3, -((w * z) * e1)));
  const double sc = fm::pow2_rescale5(n0, n1, n2, n3, n4);
  e0 = n0 * sc; e1 = n1 * sc; e2 = n2 * sc; e3 = n3 * sc; e4 = n4 * sc;
}
BH_HD void dunkin_apply_factored(const double* rec, int fs, const HalfTerms& P, const HalfTerms& S, double k2,
                                 double iomega2, double& e0, double& e1, double& e2, double& e3, double& e4) {
  DunkinTerms t;
  t.cosp = P.cs; t.w = P.sn_over_r; t.x = P.r_sn;
  t.cosq = S.cs; t.y = S.sn_over_r; t.z = S.r_sn;
  t.a0 = dunkin_a0(P, S);
  dunkin_apply_terms(rec[RR_RHO * fs], rec[RR_IRHO * fs], rec[RR_TB2 * fs] * iomega2, t, k2, e0, e1, e2, e3, e4);
}

// NL consecutive layers l0, l0-1, ... at once: 2*NL half-term chains side by side
template <int NL>
BH_HD void dunkin_layers_n(const double* rec, int fs, int ls, int l0, double wvno, double wvno2, double omega,
                           double iomega2, DunkinLayer* m) {
  constexpr int N = 2 * NL;
  double xk[N], d[N];
  HalfTerms h[N];
#pragma unroll
  for (int j = 0; j < NL; ++j) {
    const double* r = rec + (l0 - j) * ls;
    xk[2 * j] = omega * r[RR_IA * fs]; xk[2 * j + 1] = omega * r[RR_IB * fs];
    d[2 * j] = d[2 * j + 1] = r[RR_D * fs];
  }
  half_terms_n<N>(wvno, xk, d, h);
#pragma unroll
  for (int j = 0; j < NL; ++j)
    m[j] = dunkin_from_terms(rec + (l0 - j) * ls, fs, h[2 * j], h[2 * j + 1], wvno2, iomega2);
}

#define BH_DUNKIN_APPLY(m)                                                          \
  double n0 = e0 * m.c11 + e1 * m.c21 + e2 * m.c31 + e3 * m.c41 + e4 * m.c51;      \
  double n1 = e0 * m.c12 + e1 * m.c22 + e2 * m.c32 + e3 * m.c42 + e4 * m.c41;      \
  double n2 = e0 * m.c13 + e1 * m.c23 + e2 * m.c33 + e3 * m.c43 + e4 * m.c53;      \
  double n3 = e0 * m.c14 + e1 * m.c24 + e2 * m.c34 + e3 * m.c22 + e4 * m.c21;      \
  double n4 = e0 * m.c15 + e1 * m.c14 + e2 * m.c35 + e3 * m.c12 + e4 * m.c11;
#define BH_DUNKIN_STEP(m)                                                           \
  {                                                                                 \
    BH_DUNKIN_APPLY(m)                                                              \
    double sc = fm::pow2_rescale5(n0, n1, n2, n3, n4);                              \
    e0 = n0 * sc; e1 = n1 * sc; e2 = n2 * sc; e3 = n3 * sc; e4 = n4 * sc;           \
  }
BH_HD double secular_rayleigh_rec(const double* rec, int fs, int ls, int L, double wvno, double omga) {
  double omega = (omga < 1.0e-4) ? 1.0e-4 : omga;
  double iomega = fm::rcp(omega);
  double iomega2 = iomega * iomega;
  double wvno2 = wvno * wvno;
  double e0, e1, e2, e3, e4;
  {
    const double* hs = rec + (L - 1) * ls;
    double rho1 = hs[RR_RHO * fs];
    double xka = omega * hs[RR_IA * fs], xkb = omega * hs[RR_IB * fs];
    double sr[2] = {(wvno + xka) * fabs(wvno - xka), (wvno + xkb) * fabs(wvno - xkb)}, rr[2];
    sqrt_n<2>(sr, rr);
    double ra = rr[0], rb = rr[1];
    double gammk = hs[RR_TB2 * fs] * iomega2;
    double gam = gammk * wvno2;
    double gamm1 = gam - 1.0;
    e0 = rho1 * rho1 * (gamm1 * gamm1 - gam * gammk * ra * rb);
    e1 = -rho1 * ra;
    e2 = rho1 * (gamm1 - gammk * ra * rb);
    e3 = rho1 * rb;
    e4 = wvno2 - ra * rb;
  }
  if (L < 2) return e0;
  // Two layers per iteration: their matrices are independent of each other and of
  // the propagated vector, so their four half-term chains run side by side, then
  // the two matrices are applied in turn.
  // n = L-1 finite layers: an odd n starts with the bottom layer alone, then pairs.
  // Every layer ends with the exact power-of-two rescale; the reference's
  // normalisation e0 / max|e| of the top layer is taken once at the end (bit-identical
  // quotient, see secular_love_rec).  Uniform counted loops, each code block once.
  int l = L - 2;
#pragma unroll 1
  for (int r = BH_SWD_WIDE ? ((L - 1) & 1) : (L - 1); r > 0; --r, --l) {
#if BH_DUNKIN_FACTORED
    const double* rl = rec + l * ls;
    double xk[2] = {omega * rl[RR_IA * fs], omega * rl[RR_IB * fs]}, dd[2] = {rl[RR_D * fs], rl[RR_D * fs]};
    HalfTerms h[2];
    half_terms_n<2>(wvno, xk, dd, h);
    dunkin_apply_factored(rl, fs, h[0], h[1], wvno2, iomega2, e0, e1, e2, e3, e4);
#else
    DunkinLayer m;
    dunkin_layers_n<1>(rec, fs, ls, l, wvno, wvno2, omega, iomega2, &m);
    BH_DUNKIN_STEP(m)
#endif
  }
#if BH_SWD_WIDE
#pragma unroll 1
  for (int g = (L - 1) >> 1; g > 0; --g, l -= 2) {     // layers l, l-1
    DunkinLayer m[2];
    dunkin_layers_n<2>(rec, fs, ls, l, wvno, wvno2, omega, iomega2, m);
    BH_DUNKIN_STEP(m[0])
    BH_DUNKIN_STEP(m[1])
  }
#endif
  double t1 = fm::absmax(fm::absmax(fm::absmax(e0, e1), fm::absmax(e2, e3)), e4);
  if (t1 < 1.0e-40) t1 = 1.0;
  return e0 / t1;                            // IEEE division: exact +-1.0 when saturated
}

BH_HD double secular_rec(int wave, const double* rec, int fs, int ls, int L, double wvno, double omega) {
  return wave == 1 ? secular_love_rec(rec, fs, ls, L, wvno, omega)
                   : secular_rayleigh_rec(rec, fs, ls, L, wvno, omega);
}

// Row-based entry (host simulation and tests): BH_SECULAR_REFERENCE_ORDER selects
// the formulation that follows the Fortran operation by operation, for
// bit-equality with the oracle; otherwise records are derived on the fly and the
// device formulation above is evaluated.
BH_HD double secular(int wave /*1 Love, 2 Rayleigh*/, const LayerRow* rows, int stride, int L,
                     double wvno, double omega) {
#if defined(BH_SECULAR_REFERENCE_ORDER)
  return wave == 1 ? secular_love_reforder(rows, stride, L, wvno, omega)
                   : secular_rayleigh_reforder(rows, stride, L, wvno, omega);
#else
  double rec[SWD_REC_FIELDS * SWD_MAX_LAYERS];
  for (int l = 0; l < L; ++l) swd_make_rec(wave, rows[l * stride], l == L - 1, rec + l, SWD_MAX_LAYERS);
  return secular_rec(wave, rec, SWD_MAX_LAYERS, 1, L, wvno, omega);
#endif
}

// ---------------------------------------------------------------------------
// gtsolh (:367-388): five REAL*4 Newton steps on the half-space Rayleigh
// equation; every operation individually rounded (no FMA).
// ---------------------------------------------------------------------------
BH_HD float halfspace_start(float a, float b) {
  float c = fmul(0.95f, b);
  for (int i = 0; i < 5; ++i) {
    float gamma = fdiv(b, a);
    float kappa = fdiv(c, b);
    float k2 = fmul(kappa, kappa);
    float gk = fmul(gamma, kappa);
    float gk2 = fmul(gk, gk);
    float fac1 = fsqrt(fsub(1.0f, gk2));
    float fac2 = fsqrt(fsub(1.0f, k2));
    float tk = fsub(2.0f, k2);
    float fr = fsub(fmul(tk, tk), fmul(fmul(4.0f, fac1), fac2));
    float t1 = fmul(fmul(-4.0f, tk), kappa);
    float t2 = fdiv(fmul(fmul(fmul(fmul(4.0f, fac2), gamma), gamma), kappa), fac1);
    float t3 = fdiv(fmul(fmul(4.0f, fac1), kappa), fac2);
    float frp = fadd(fadd(t1, t2), t3);
    frp = fdiv(frp, b);
    c = fsub(c, fdiv(fr, frp));
  }
  return c;
}

// ---------------------------------------------------------------------------
// Search state machine
//
// A phase-velocity curve is ONE serial chain of root searches: period k starts
// from c(k-1) - 1.5 dc (surfdisp96.f:268-271).  A group-velocity curve needs two
// roots per period, at t1a = T/(1+h) and t1b = T/(1-h) (:231-239, :282-294); the
// reference finds them one after the other, but the data flow is
//     first root c(k)   <- c(k-1)               (same chain as a phase curve at t1a)
//     second root cb(k) <- c(k) only            (start c(k) - 1.5 dc, clow = 0.01 dc)
// so the second roots hang off the first-root chain and never feed back into it
// (cb(k) is read only for higher modes).  The engine therefore runs a group
// curve as TWO chains on two lanes: role A searches the first roots and
// publishes c(k); role B searches the second roots as they become available.
// Each chain consumes exactly the reference's candidate sequence.
// ---------------------------------------------------------------------------
enum SearchStage : int {
  ST_BR_FIRST = 0,   // waiting for del1 = secular(c1)             (getsol :429)
  ST_BR_STEP = 1,    // waiting for del2 at c2 = c1 +- dc           (getsol :448-470)
  ST_RF_TOP = 2,     // waiting for del3; resume at nevill label 100 (:587)
  ST_RF_POST = 3,    // waiting for del3 of the range-fix half (:596); resume at :599
  ST_WAIT = 4,       // role B: first root of period k not published yet
  ST_DONE = 5,       // all periods found
  ST_FAILED = 6      // err = 1 (no root for a period of the fundamental mode)
};

struct Search {
  // per-search constants
  double cc, dc, betmx;      // start value, dble(0.005f), dble(REAL*4 max vs)
  int kmax, role;            // role 0: first roots (A), 1: second roots of a group curve (B)
  // period bookkeeping
  int k;                     // 0-based period index
  int stage;
  int ifirst, idir;
  double omega;              // twopi / t1 of the root being searched
  double c1, c2, del1, del2, clow, del1st;
  double cprev;              // A: c(k-1);  B: c(k) of the period being searched
  // nevill
  double c3, del3;
  int nev, nctrl, m;
  // Neville tableau x(1..11), y(1..11) (:571): x(j) = tab[j * ts], y(j) = tab[(11 + j) * ts].
  // On the device it lives in shared memory, one column per lane (ts = 32), because a
  // dynamically indexed member would put the whole struct into local memory.
  double* tab;
  int ts;
};
constexpr int SWD_TAB_ROWS = 22;

// What role A publishes for role B (one per group search; shared memory on the device)
struct SearchLink {
  int na;          // first roots found so far: B may search periods k < na
  int a_failed;    // A ended in ST_FAILED
  double del1st;   // getsol's SAVEd del1st (:415,430), set at the first period
};

// Per-curve context of one search: period tables and root storage
struct SearchCtx {
  const double* omA;   // [kmax] twopi / t1       (phase: t1 = T;  group: t1 = dble(t1a))
  const double* omB;   // [kmax] twopi / dble(t1b) (group only)
  double* ra;          // [kmax] first roots c(k)
  double* rb;          // [kmax] second roots cb(k) (group only)
  SearchLink* link;    // group only
};

constexpr double SWD_TWOPI = 2.0 * 3.141592653589793;

// t1a = t1/(1.+h), t1b = t1/(1.-h): REAL*4 sums, REAL*8 quotient, REAL*4 store (:233-235)
BH_HD void swd_group_periods(double period, float* t1a, float* t1b) {
  *t1a = (float)(period / (double)fadd(1.0f, 0.005f));
  *t1b = (float)(period / (double)fsub(1.0f, 0.005f));
}
// angular frequencies of the searches of one period (getsol :426: omega = twopi/t1)
BH_HD void swd_period_omegas(int igr, double period, double* omA, double* omB) {
  if (igr > 0) {
    float t1a, t1b;
    swd_group_periods(period, &t1a, &t1b);
    *omA = SWD_TWOPI / (double)t1a;
    *omB = SWD_TWOPI / (double)t1b;
  } else {
    *omA = SWD_TWOPI / period;
    *omB = 0.0;
  }
}
// cg(k) from the stored roots (:298-310): phase = sngl(c(k)); group velocity formula
// evaluated entirely in REAL*4, each operation rounded
BH_HD double swd_curve_value(int igr, double period, double ra, double rb) {
  if (igr <= 0) return (double)(float)ra;
  float t1a, t1b;
  swd_group_periods(period, &t1a, &t1b);
  float cc0 = (float)ra, cc1 = (float)rb;
  float num = fsub(fdiv(1.0f, t1a), fdiv(1.0f, t1b));
  float den = fsub(fdiv(1.0f, fmul(t1a, cc0)), fdiv(1.0f, fmul(t1b, cc1)));
  return (double)fdiv(num, den);
}

BH_HD bool sign_differs(double a, double b) {   // dsign(1,a) != dsign(1,b), +-0 aware
#if defined(__CUDA_ARCH__)
  return (__double2hiint(a) ^ __double2hiint(b)) < 0;
#else
  return (signbit(a) != 0) != (signbit(b) != 0);
#endif
}

// Extremal velocities + start value (surfdisp96.f:139-156, 197-217).  Water
// layers (vs <= 0.01) are outside this engine's scope: vs > 0 is enforced by
// BayHunter's priors (SingleChain.py:358-363); such a model is reported failed.
BH_HD bool search_setup(Search& s, const LayerRow* rows, int stride, int L, int kmax, int role,
                        double* tab, int ts) {
  s.tab = tab; s.ts = ts;
  float betmx = -1.e20f, betmn = 1.e20f;
  int jmn = 0;
  bool solid = true;
  for (int i = 0; i < L; ++i) {
    LayerRow r = rows[i * stride];
    if (r.z > 0.01f && r.z < betmn) { betmn = r.z; jmn = i; }
    else if (r.z <= 0.01f) solid = false;
    if (r.z > betmx) betmx = r.z;
  }
  s.kmax = kmax; s.role = role;
  s.dc = fabs((double)0.005f);
  s.betmx = (double)betmx;
  s.k = 0;
  s.cprev = 0.0; s.del1st = 0.0; s.omega = 0.0;
  s.c1 = s.c2 = s.del1 = s.del2 = s.clow = 0.0;
  s.c3 = s.del3 = 0.0; s.nev = 0; s.nctrl = 0; s.m = 0;
  s.ifirst = 0; s.idir = 1;
  if (!solid || L < 1 || kmax < 1) { s.cc = 0.0; s.stage = ST_FAILED; return false; }
  LayerRow rm = rows[jmn * stride];
  float cc1 = halfspace_start(rm.y, rm.z);
  cc1 = fmul(.95f, cc1);
  cc1 = fmul(.90f, cc1);
  s.cc = (double)cc1;
  s.stage = role ? ST_WAIT : ST_BR_FIRST;
  return true;
}

// Role A: start the search of period s.k (first root)
BH_HD void search_begin_a(Search& s, const SearchCtx& ctx) {
  if (s.k == 0) {            // :253-256
    s.c1 = s.cc; s.clow = s.cc; s.ifirst = 1;
  } else {                   // :268-271
    s.ifirst = 0;
    s.c1 = dadd(s.cprev, -dmul(1.5, s.dc));
    s.clow = s.cc;           // clow = cm = cc
  }
  s.omega = ctx.omA[s.k];
  s.stage = ST_BR_FIRST;
}

// Role B: start the second root of period s.k; c(k) = ctx.ra[s.k] is known (:282-287)
BH_HD void search_begin_b(Search& s, const SearchCtx& ctx, double del1st) {
  s.cprev = ctx.ra[s.k];
  s.del1st = del1st;
  s.ifirst = 0;
  s.clow = dmul(1.0e-2, s.dc);                  // cb(k) + one*dc, cb(k) = 0 in the fundamental mode
  s.c1 = dadd(s.cprev, -dmul(1.5, s.dc));
  s.omega = ctx.omB[s.k];
  s.stage = ST_BR_FIRST;
}
// ... if role A has published c(k)
BH_HD void search_poll_b(Search& s, const SearchCtx& ctx) {
  if (s.stage != ST_WAIT) return;
  if (s.k < ctx.link->na) search_begin_b(s, ctx, ctx.link->del1st);
  else if (ctx.link->a_failed) s.stage = ST_FAILED;
}

// Next bracket candidate of getsol's loop (:448-458), advancing (c1, idir).
BH_HD double bracket_next(double& c1, int& idir, double clow, double dc) {
  double c2 = (idir > 0) ? c1 + dc : c1 - dc;
  if (c2 <= clow) {
    idir = +1;
    c1 = clow;
    c2 = c1 + dc;
  }
  return c2;
}

// How many candidates the search can use this round (>= 1 while running).
// A search that evaluates its first candidate c1 (ST_BR_FIRST) walks upwards from it next unless the sign of
// that first value says otherwise (:432-438; it does in ~6 % of the searches): the candidates c1 + dc, c1 + 2 dc, ...
// ride along on spare lanes and are consumed only if the direction turns out to be +1.
BH_HD int search_nwant(const Search& s, int nmax) {
  if (s.stage >= ST_WAIT) return 0;
  return s.stage <= ST_BR_STEP ? nmax : 1;
}

// i-th pending candidate (i = 0 is the one the reference evaluates next).
// The published (stage, c, idir, clow) tuple is all a worker lane needs.
// A refining chain (stage > ST_BR_STEP) publishes c = c3, clow = c1 and c2: candidate 0 is c3, candidates 1 and 2 are
// the two midpoints the NEXT step asks for if it bisects, 3..6 the four midpoints the step after that can ask for if
// both bisect, 7..14 and 15..30 those of a third and fourth step: a binary tree in heap order (node i has the children 2i + 1, 2i + 2; the
// left child halves the bracket on the c1 side).  search_consume uses a value only if the step asks for exactly that
// velocity.
BH_HD double refine_guess(double c1, double c2, double c3, int i) {
  int dirs = 0, depth = 0;
  for (int j = i; j > 0; j = (j - 1) >> 1) { dirs = (dirs << 1) | ((j & 1) ? 0 : 1); ++depth; }   // root step in bit 0
  double lo = c1, hi = c2, pt = c3;
  for (int d = 0; d < depth; ++d) {
    if (dirs & 1) lo = pt; else hi = pt;
    dirs >>= 1;
    pt = 0.5 * (lo + hi);
  }
  return pt;
}
// Guess lanes per refining chain: the deepest tree that fits `room` lanes for `nrf` refining chains.
constexpr int kRefineGuesses = 2, kRefineGuesses2 = 6, kRefineGuesses3 = 14, kRefineGuesses4 = 30;
BH_HD int refine_guess_lanes(int nrf, int room) {
  if (nrf <= 0) return 0;
  if (kRefineGuesses4 * nrf <= room) return kRefineGuesses4;      // a chain alone in its warp
  return kRefineGuesses3 * nrf <= room ? kRefineGuesses3 : (kRefineGuesses2 * nrf <= room ? kRefineGuesses2 : (kRefineGuesses * nrf <= room ? kRefineGuesses : 0));
}
// fast_up: the shortcut for upward walks (same operations; measured faster in swd_pool_kernel, slower in swd_kernel)
BH_HD double candidate_from(int stage, double c, int idir, double clow, double dc, int i, double c2r = 0.0,
                            const bool fast_up = false) {
  if (stage > ST_BR_STEP) return i == 0 ? c : refine_guess(clow, c2r, c, i);
  if (stage == ST_BR_FIRST) {            // c1 itself, then the upward walk from it
    if (i == 0) return c;
    idir = +1;
    i -= 1;
  }
  double c1 = c, c2 = c;
  if (fast_up && idir > 0) {
    // walking up, only the first step can be lifted to the lower end of the window; from then on c2 = c1 + dc
    c2 = bracket_next(c1, idir, clow, dc);
    for (int j = 1; j <= i; ++j) c2 += dc;
    return c2;
  }
  for (int j = 0; j <= i; ++j) {
    c2 = bracket_next(c1, idir, clow, dc);
    c1 = c2;
  }
  return c2;
}
BH_HD double search_pending_c(const Search& s) {
  return (s.stage == ST_BR_FIRST || s.stage == ST_BR_STEP) ? s.c1 : s.c3;
}

// ---- lane dealing ---------------------------------------------------------
// One round of a warp: `active` = lanes whose chain needs a secular value, `bracket` (a subset) = lanes whose
// chain walks a bracket and can use more than one.  A refining chain gets one lane; the 32 - popc(active)
// spare lanes go evenly to the walking chains (quo or quo + 1 more each, at most max_spec lanes per chain).
// Returns, for `lane`: cnt = lanes of its chain, excl = first lane of its run (the runs are laid out in lane
// order), total = lanes in use.  Closed form from the two masks -- no scan, no integer division:
// (extra + 0.5) / nbr stays >= 1/64 away from every integer for 0 <= extra <= 32, 1 <= nbr <= 32, far beyond
// the error of the approximate fp32 quotient.
// When lanes are left over even after every walking chain has four (small batches, deep models), each refining chain gets
// 2 more for the two midpoints its next step may ask for, or 6 / 14 / 30 for those of its next two / three / four steps
// (refine_guess_lanes, search_consume).
#ifndef BH_GUESS_WALK_EXTRA
#define BH_GUESS_WALK_EXTRA 3
#endif
struct LaneDeal { int cnt, excl, total; };
BH_HD LaneDeal deal_lanes(unsigned active, unsigned bracket, int lane, int max_spec) {
#if defined(__CUDA_ARCH__)
#define BH_POPC(x) __popc(x)
#else
#define BH_POPC(x) __builtin_popcount(x)
#endif
  const int nact = BH_POPC(active), nbr = BH_POPC(bracket);
  const int nrf = nact - nbr;
  const int g = refine_guess_lanes(nrf, 32 - nact - BH_GUESS_WALK_EXTRA * nbr);
  const int extra = 32 - nact - g * nrf;
#if defined(__CUDA_ARCH__)
  const int quo = nbr ? __float2int_rz(__fdividef((float)extra + 0.5f, (float)nbr)) : 0;
#else
  const int quo = nbr ? (int)(((float)extra + 0.5f) / (float)nbr) : 0;
#endif
  const int rem = extra - quo * nbr;
  const int per = quo + 1;                               // lanes of a walking chain of rank >= rem
  const bool capped = per >= max_spec;
  const unsigned below = (1u << lane) - 1u;
  const int rank = BH_POPC(bracket & below);
  const int act_below = BH_POPC(active & below);
  const unsigned me = 1u << lane;
  LaneDeal d;
  d.cnt = 0;
  if (active & me) {
    d.cnt = 1 + g;
    if (bracket & me) { d.cnt = per + (rank < rem ? 1 : 0); if (d.cnt > max_spec) d.cnt = max_spec; }
  }
  // walking chains below this lane hold rank * per + min(rank, rem) lanes (or rank * max_spec when capped),
  // refining chains below it 1 + g each
  d.excl = (act_below - rank) * (1 + g) + (capped ? rank * max_spec : rank * per + (rank < rem ? rank : rem));
  d.total = nrf * (1 + g) + (capped ? nbr * max_spec : nbr * per + (nbr < rem ? nbr : rem));
#undef BH_POPC
  return d;
}

// ---- nevill pieces -------------------------------------------------------
// x(j) = (-y(j) x(j+1) + y(m+1) x(j)) / (y(m+1) - y(j))  (:651-653), one spelling for every kernel
BH_HD double neville_step(double yj, double xnext, double ym, double xj, double denom) {
#if defined(__CUDA_ARCH__)
  return fm::div(fma(-yj, xnext, ym * xj), denom);
#else
  return (-yj * xnext + ym * xj) / denom;      // host simulation: the Fortran's operations, no contraction
#endif
}

BH_HD void nevill_request_half(Search& s, int next_stage) {
  s.c3 = 0.5 * (s.c1 + s.c2);
  s.stage = next_stage;
}

// The search of the current root ended with value `root`; found = false when
// getsol gave up (bracket left the window, or root > betmx :476).
BH_HD void search_root_end(Search& s, const SearchCtx& ctx, double root, bool found) {
  if (s.role == 0) {
    if (!found) {                                  // :277 -> 1700, err = 1
      s.stage = ST_FAILED;
      if (ctx.link) ctx.link->a_failed = 1;
      return;
    }
    ctx.ra[s.k] = root;                            // c(k) = c1
    s.cprev = root;
    s.k += 1;
    if (ctx.link) ctx.link->na = s.k;
    if (s.k >= s.kmax) { s.stage = ST_DONE; return; }
    search_begin_a(s, ctx);
  } else {
    ctx.rb[s.k] = found ? root : s.cprev;          // :291-293: no second root -> c1 = c(k)
    s.k += 1;
    // the next period starts at the caller's next search_poll_b: not here, where the first-root lane of the
    // same warp may be publishing c(k) in this very phase (compute-sanitizer racecheck, profiles/r02_sanitizer.txt)
    s.stage = (s.k >= s.kmax) ? ST_DONE : ST_WAIT;
  }
}

// nevill main loop from label 100 (:587) until the next secular evaluation is
// needed (returns false with stage = ST_RF_TOP / ST_RF_POST and c3 set) or the
// root is accepted (returns true).
BH_HD bool nevill_resume(Search& s, bool at_top) {
  for (;;) {
    if (at_top) {
      s.nctrl += 1;
      if (s.nctrl >= 100) return true;                                   // :589
      if (s.c3 < fmin(s.c1, s.c2) || s.c3 > fmax(s.c1, s.c2)) {          // :594-598
        s.nev = 0;
        nevill_request_half(s, ST_RF_POST);
        return false;
      }
    }
    at_top = true;
    double s13 = s.del1 - s.del3;
    double s32 = s.del3 - s.del2;
    if (sign_differs(s.del3, s.del1)) { s.c2 = s.c3; s.del2 = s.del3; }  // :604-610
    else { s.c1 = s.c3; s.del1 = s.del3; }
    if (fabs(s.c1 - s.c2) <= 1.0e-6 * s.c1) return true;                 // :614
    if (sign_differs(s13, s32)) s.nev = 0;                               // :619
    double ss1 = fabs(s.del1);
    double s1 = (double)0.01f * ss1;                                     // REAL*4 literal (:625)
    double ss2 = fabs(s.del2);
    double s2 = (double)0.01f * ss2;
    if (s1 > ss2 || s2 > ss1 || s.nev == 0) {                            // :628-632
      nevill_request_half(s, ST_RF_TOP);
      s.nev = 1;
      s.m = 1;
      return false;
    }
    double* const x = s.tab;
    double* const y = s.tab + 11 * s.ts;
    const int ts = s.ts;
    if (s.nev == 2) {                                                    // :634-643
      x[s.m * ts] = s.c3;
      y[s.m * ts] = s.del3;
    } else {
      x[0] = s.c1; y[0] = s.del1;
      x[ts] = s.c2; y[ts] = s.del2;
      s.m = 1;
    }
    bool bad = false;
    const double ym = y[s.m * ts];
    double xn = x[s.m * ts];                                             // x(j+1), updated as we go down
    for (int kk = 1; kk <= s.m; ++kk) {                                  // :649-654
      int j = s.m - kk;
      const double yj = y[j * ts];
      double denom = ym - yj;
      if (fabs(denom) < 1.0e-10 * fabs(ym)) { bad = true; break; }
      xn = neville_step(yj, xn, ym, x[j * ts], denom);
      x[j * ts] = xn;
    }
    if (bad) {                                                           // :663-667
      nevill_request_half(s, ST_RF_TOP);
      s.nev = 1;
      s.m = 1;
      return false;
    }
    s.c3 = x[0];                                                         // :655-661
    s.nev = 2;
    s.m = s.m + 1;
    if (s.m > 10) s.m = 10;
    s.stage = ST_RF_TOP;
    return false;
  }
}

// Consume n secular values del[0..n) for the candidates this search published
// this round (n = 1 unless stage == ST_BR_STEP).  Returns how many of the n
// values were consumed (the rest was speculation past a sign change or past
// the search window).
BH_HD int search_consume(Search& s, const double* del, int n, const SearchCtx& ctx, const bool guesses = true,
                         const bool fast_up = false) {
  int first = 0;
  switch (s.stage) {
    case ST_BR_FIRST: {
      s.del1 = del[0];
      if (s.ifirst == 1) {                                               // :430
        s.del1st = s.del1;
        if (ctx.link) ctx.link->del1st = s.del1;
      }
      s.idir = (s.ifirst == 1 || !sign_differs(s.del1st, s.del1)) ? +1 : -1;   // :432-438
      s.stage = ST_BR_STEP;
      if (n == 1 || s.idir < 0) return 1;
      first = 1;                          // the values that rode along belong to the upward walk
    }
    // fall through
    case ST_BR_STEP: {
      if (fast_up && s.idir > 0 && first < n) {
        // the upward walk (94 % of the searches): the same steps with the state in registers and without the
        // per-step test against the lower end of the window, which only the first step can fail
        const double lim = s.betmx + s.dc;
        double c1 = s.c1, d1 = s.del1;
        double c2 = bracket_next(c1, s.idir, s.clow, s.dc), d2 = 0.0;
        int i = first;
        for (;;) {
          d2 = del[i];
          if (sign_differs(d1, d2)) {                                    // :462 -> nevill
            s.c1 = c1; s.del1 = d1; s.c2 = c2; s.del2 = d2;
            nevill_request_half(s, ST_RF_TOP);                           // :583
            s.nev = 1;
            s.nctrl = 1;
            return i + 1;
          }
          c1 = c2; d1 = d2;
          if (c1 < s.cc || c1 >= lim) {                                  // :468-469 (cm = cc)
            s.c1 = c1; s.del1 = d1; s.c2 = c2; s.del2 = d2;
            search_root_end(s, ctx, 0.0, false);
            return i + 1;
          }
          if (++i >= n) break;
          c2 = c1 + s.dc;
        }
        s.c1 = c1; s.del1 = d1; s.c2 = c2; s.del2 = d2;
        return n;
      }
      for (int i = first; i < n; ++i) {
        s.c2 = bracket_next(s.c1, s.idir, s.clow, s.dc);
        s.del2 = del[i];
        if (sign_differs(s.del1, s.del2)) {                              // :462 -> nevill
          nevill_request_half(s, ST_RF_TOP);                             // :583
          s.nev = 1;
          s.nctrl = 1;
          return i + 1;
        }
        s.c1 = s.c2;
        s.del1 = s.del2;
        if (s.c1 < s.cc || s.c1 >= (s.betmx + s.dc)) {                   // :468-469 (cm = cc)
          search_root_end(s, ctx, 0.0, false);
          return i + 1;
        }
      }
      return n;
    }
    case ST_RF_TOP:
    case ST_RF_POST: {
      // values 1 and 2 (if dealt) belong to the midpoints of the two half brackets this step can leave behind
      double lo = s.c1, hi = s.c2, pt = s.c3;
      s.del3 = del[0];
      int used = 1, node = 0;
      bool done = nevill_resume(s, s.stage == ST_RF_TOP);
      // the step bisected into one of the two half brackets: its value is here already, and so on down the tree
      while (guesses && !done && 2 * node + 2 < n) {
        const double gl = 0.5 * (lo + pt), gr = 0.5 * (pt + hi);
        if (s.c3 == gl) { hi = pt; node = 2 * node + 1; }
        else if (s.c3 == gr) { lo = pt; node = 2 * node + 2; }
        else break;
        pt = s.c3;
        s.del3 = del[node];
        used += 1;
        done = nevill_resume(s, s.stage == ST_RF_TOP);
      }
      if (done) search_root_end(s, ctx, s.c3, !(s.c3 > s.betmx));        // :475-476
      return used;
    }
    default:
      return 0;
  }
}

}  // namespace bh
