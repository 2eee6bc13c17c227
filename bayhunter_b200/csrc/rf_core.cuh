// rf_core.cuh -- receiver-function core: earth flattening, interface
// reflection/transmission matrices, per-frequency reflectivity recursion,
// P/SV decomposition, spectral division and Gauss filter.
//
// Behavioural reference (BayHunter, src/extensions/rfmini/):
//   synrf.cpp:16-55          model build (thickness from z differences)
//   model.cpp:209-251        FlatLayer::flatten, isLowerHalfspace (R = 6371 km)
//   greens.cpp:19-112        coeffm / coeffs
//   greens.cpp:307-322       displacement_matrix
//   greens.cpp:528-549       complex-Q vertical slownesses and phase terms
//   greens.cpp:196-224       top_down (Mueller 1985 / Kennett recursion)
//   greens.cpp:324-398       decomp, compute_rf
//   wrap.cpp:13,55,73-76     vptop from Poisson ratio, s/deg -> s/km
//
// Layout decisions are this engine's own: per model a small table of flattened
// layer constants plus one 32-double block (rd, td, ru, tu) per interface is
// produced once by a preparation kernel; the spectrum kernel is parallel over
// (model, frequency) items and only reads those tables.
#pragma once
#include "bh_common.cuh"
#include "bh_math.cuh"

namespace bh {

constexpr double RF_EARTH_RADIUS = 6371.0;     // model.cpp:220
constexpr double RF_DEG_PER_KM = 0.00899;      // wrap.cpp:55
constexpr double RF_PI = 3.14159265358979323846;

// flattened layer constants, 8 doubles
struct RfLayer {
  double h;      // flattened thickness (half-space: -1, unused)
  double vp, vs; // flattened velocities
  double rho;    // flattened density
  double cqp;    // 1 / (pi * qp)
  double bqp;    // 1 / (2 * qp)
  double cqs, bqs;
};

// model.cpp:223-251
BH_HD void rf_flatten(double z, double h, double vp, double vs, double rho,
                      double* h_out, double* vp_out, double* vs_out, double* rho_out) {
  double zb = z + h;
  double r = RF_EARTH_RADIUS - z;
  double q = RF_EARTH_RADIUS / r;
  double zf = RF_EARTH_RADIUS * log(q);
  double vpf = vp * q, vsf = vs * q, rhof = rho / q;
  bool lower_halfspace = !(h > 0.0) && !(vpf < 1.0 && rhof < 0.1);
  double hf = h;
  if (!lower_halfspace) {
    r = RF_EARTH_RADIUS - zb;
    q = RF_EARTH_RADIUS / r;
    zb = RF_EARTH_RADIUS * log(q);
    hf = zb - zf;
  }
  *h_out = hf; *vp_out = vpf; *vs_out = vsf; *rho_out = rhof;
}

BH_HD cd rf_vslow(double v, double u2, bool conj) {
  cd s = csqrt_p(mk(1.0 / (v * v) - u2, 0.0));
  return conj ? cconj(s) : s;
}

// greens.cpp:19-77 -- P-SV coefficients of a welded interface (1 above, 2 below).
// Cmat2(pp, sp, ps, ss): a11 = pp, a12 = sp, a21 = ps, a22 = ss.
BH_HD void rf_coeff_interface(double u, double vp1, double vs1, double rho1, double vp2, double vs2,
                              double rho2, cm2* rd, cm2* td, cm2* ru, cm2* tu) {
  double mue1 = rho1 * vs1 * vs1, mue2 = rho2 * vs2 * vs2;
  double c = 2.0 * (mue1 - mue2), u2 = u * u, cu2 = c * u2;
  cd a1 = rf_vslow(vp1, u2, true), a2 = rf_vslow(vp2, u2, true);
  cd b1 = rf_vslow(vs1, u2, true), b2 = rf_vslow(vs2, u2, true);
  double t1 = cu2 - rho1 + rho2, t2 = cu2 - rho1, t3 = cu2 + rho2;
  cd t4 = t3 * a1 - t2 * a2;
  double r12 = rho1 * rho2;
  cd a2b2 = a2 * b2, a2b1 = a2 * b1, a1b1 = a1 * b1, a1b2 = a1 * b2;
  cd quad = (c * c * u2) * (a1 * a2 * b1 * b2);
  // downgoing incidence (table 1)
  cd d1 = mk(t1 * t1 * u2, 0.0) + (t2 * t2) * a2b2 + r12 * a2b1;
  cd d2 = quad + (t3 * t3) * a1b1 + r12 * a1b2;
  cd t5 = crecip(d1 + d2);
  cd t7 = (2.0 * rho1) * t5;
  cd k1 = mk(t1 * t3, 0.0) + (c * t2) * a2b2;
  rd->a11 = (d2 - d1) * t5;                                   // rpp
  rd->a21 = (-2.0 * u) * a1 * t5 * k1;                        // rps
  td->a11 = a1 * t7 * (t3 * b1 - t2 * b2);                    // tpp
  td->a21 = -(a1 * t7 * u * (mk(t1, 0.0) + c * a2b1));        // tps
  rd->a22 = (d2 - d1 - (2.0 * r12) * (a1b2 - a2b1)) * t5;     // rss
  rd->a12 = (2.0 * u) * b1 * t5 * k1;                         // rsp
  td->a22 = b1 * t7 * t4;                                     // tss
  td->a12 = b1 * t7 * u * (mk(t1, 0.0) + c * a1b2);           // tsp
  // upgoing incidence (table 2)
  d1 = mk(t1 * t1 * u2, 0.0) + (t3 * t3) * a1b1 + r12 * a1b2;
  d2 = quad + (t2 * t2) * a2b2 + r12 * a2b1;
  t5 = crecip(d1 + d2);
  t7 = (2.0 * rho2) * t5;
  cd k2 = mk(t1 * t2, 0.0) + (c * t3) * a1b1;
  ru->a11 = (d2 - d1) * t5;                                   // rpp
  ru->a21 = (2.0 * u) * a2 * t5 * k2;                         // rps
  tu->a11 = a2 * t7 * (t3 * b1 - t2 * b2);                    // tpp
  tu->a21 = -(a2 * t7 * u * (mk(t1, 0.0) + c * a1b2));        // tps
  ru->a22 = (d2 - d1 - (2.0 * r12) * (a2b1 - a1b2)) * t5;     // rss
  ru->a12 = (-2.0 * u) * b2 * t5 * k2;                        // rsp
  tu->a22 = b2 * t7 * t4;                                     // tss
  tu->a12 = b2 * t7 * u * (mk(t1, 0.0) + c * a2b1);           // tsp
}

// greens.cpp:87-108 -- free-surface reflection (slownesses NOT conjugated)
BH_HD void rf_coeff_surface(double u, double vp, double vs, cm2* ru) {
  double u2 = u * u;
  cd a = rf_vslow(vp, u2, false), b = rf_vslow(vs, u2, false);
  double t1 = 2.0 * vs * vs;
  double t2 = t1 * u2 - 1.0;
  double d1 = t2 * t2;
  cd d2 = (t1 * t1 * u2) * (a * b);
  cd d = mk(d1, 0.0) + d2;
  cd dinv = crecip(d);
  cd t3 = (2.0 * t1 * u * t2) * dinv;
  cd rpp = (d2 - mk(d1, 0.0)) * dinv;
  ru->a11 = rpp;
  ru->a12 = -(b * t3);   // rsp
  ru->a21 = a * t3;      // rps
  ru->a22 = rpp;
}

// greens.cpp:307-322 -- free-surface displacement matrix, already doubled
// (the reference forms t = 2*h*g, greens.cpp:572)
BH_HD void rf_displacement2(double p, double vp, double vs, cm2* m) {
  double vs2 = vs * vs, p2 = p * p, x = 1.0 - 2.0 * vs2 * p2;
  cd a1 = rf_vslow(vp, p2, true), b1 = rf_vslow(vs, p2, true);
  cd q = crecip(mk(x * x, 0.0) + (4.0 * vs2 * vs2 * p2) * (a1 * b1));
  cd q2 = 2.0 * q;
  m->a11 = q2 * a1 * b1 * (2.0 * vs2 * p);
  m->a12 = q2 * b1 * x;
  m->a21 = q2 * a1 * x;
  m->a22 = -(q2 * a1 * b1 * (2.0 * vs2 * p));
}

// Constants of compute_rf's P/SV decomposition (greens.cpp:324-341,365-367);
// m[4] = (m11, m12, m21, m22); enabled = vs_top > 0.01 && |u| > 1e-4.
BH_HD bool rf_decomp_consts(double u, double nsv, double sigma, double* m) {
  double vptop = nsv * sqrt((1.0 - sigma) / (0.5 - sigma));   // wrap.cpp:13,73
  double vstop = nsv;
  bool on = (vstop > 0.01) && (fabs(u) > 0.0001);
  double a = sqrt(1.0 / (vptop * vptop) - u * u), b = sqrt(1.0 / (vstop * vstop) - u * u);
  m[0] = -(2.0 * vstop * vstop * u * u - 1.0) / (vptop * a);
  m[1] = 2.0 * u * vstop * vstop / vptop;
  m[2] = -2.0 * u * vstop;
  m[3] = (1.0 - 2.0 * vstop * vstop * u * u) / (vstop * b);
  return on;
}

// Complex helpers of the per-(layer, frequency) inner loop.  On the device they use the straight-line
// reciprocal / square root / exp / sincos of bh_math.cuh (<= 2 ulp, no slow-path branches; arguments here
// are finite, normal and of modest size: |Im| of a phase <= a few hundred, Re <= 0 up to attenuation);
// the host build (tests/host_sim) keeps libm.  BH_RF_FAST=0 restores the CUDA library versions.
#ifndef BH_RF_FAST
#define BH_RF_FAST 1
#endif
BH_HD cd rf_crecip(cd a) {
#if BH_RF_FAST && defined(__CUDA_ARCH__)
  double d = fm::rcp(cnorm(a));
  return mk(a.re * d, -a.im * d);
#else
  return crecip(a);
#endif
}
BH_HD cd rf_csqrt(cd z) {
#if BH_RF_FAST && defined(__CUDA_ARCH__)
  double x = z.re, y = z.im;
  if (x == 0.0 && y == 0.0) return mk(0.0, y);
  double r, rs, t, ts;
  fm::sqrt_rsqrt(fma(x, x, y * y), &r, &rs);
  fm::sqrt_rsqrt(0.5 * (r + fabs(x)), &t, &ts);
  double h = 0.5 * ts;                                   // 1 / (2 t)
  if (x >= 0.0) return mk(t, y * h);
  return mk(fabs(y) * h, copysign(t, y));
#else
  return csqrt_p(z);
#endif
}
BH_HD cd rf_cexp(cd z) {
#if BH_RF_FAST && defined(__CUDA_ARCH__)
  double s, c;
  fm::sincos_cw(z.im, &s, &c);
  double e = fm::exp_small(z.re < -700.0 ? -700.0 : z.re);     // below: 1e-304 stands for 0
  return mk(e * c, e * s);
#else
  return cexp_d(z);
#endif
}

// Phase terms of one layer at angular frequency w (greens.cpp:533-548):
// complex-Q velocities v*(1 + lgw/(pi*Q) + i/(2Q)), principal-branch vertical
// slownesses, e = exp(-i*w*d*slowness).
BH_HD void rf_phase(const RfLayer& L, double u2, double w, double lgw, cd* ep, cd* es) {
  cd vpc = L.vp * mk(1.0 + lgw * L.cqp, L.bqp);
  cd vsc = L.vs * mk(1.0 + lgw * L.cqs, L.bqs);
  cd plc = rf_csqrt(rf_crecip(vpc * vpc) - mk(u2, 0.0));
  cd slc = rf_csqrt(rf_crecip(vsc * vsc) - mk(u2, 0.0));
  double wd = -w * L.h;
  // (0, wd) * plc = (-wd*plc.im, wd*plc.re)
  *ep = rf_cexp(mk(-wd * plc.im, wd * plc.re));
  *es = rf_cexp(mk(-wd * slc.im, wd * slc.re));
}

// One (model, frequency) item: reflectivity recursion over the nlay-1 finite
// layers, returns t = 2*H*g[nlay-1].  coef holds, for interface i = 1..nlay
// (0-based block i-1), the 4 matrices rd, td, ru, tu (block 0: only ru is
// meaningful = free surface).
BH_HD cm2 rf_transfer(const RfLayer* lay, const cm2* coef, const cm2& h2, int nlay, double u2,
                      double w, double lgw) {
  cm2 nb, q, g;
  nb.a11 = nb.a12 = nb.a21 = nb.a22 = mk(0.0, 0.0);
  q = nb; g = nb;
  for (int i = 1; i < nlay; ++i) {
    cd ep, es;
    rf_phase(lay[i - 1], u2, w, lgw, &ep, &es);
    const cm2* ci = coef + 4 * (i - 1);       // interface i: rd, td, ru, tu
    const cm2* cn = coef + 4 * i;             // interface i+1
    cm2 nt;
    if (i == 1) nt = ci[2];
    else {
      cm2 tq = mmul(mmul(ci[1], nb), q);
      nt.a11 = ci[2].a11 + tq.a11; nt.a12 = ci[2].a12 + tq.a12;
      nt.a21 = ci[2].a21 + tq.a21; nt.a22 = ci[2].a22 + tq.a22;
    }
    cd epp = ep * ep, eps = ep * es, ess = es * es;     // exe(), greens.cpp:829-845
    nb.a11 = nt.a11 * epp; nb.a12 = nt.a12 * eps;
    nb.a21 = nt.a21 * eps; nb.a22 = nt.a22 * ess;
    cm2 r = mmul(cn[0], nb);                            // rd[i+1]*nb[i]
    cd m11 = mk(1.0, 0.0) - r.a11, m12 = -r.a12, m21 = -r.a21, m22 = mk(1.0, 0.0) - r.a22;
    cd dinv = rf_crecip(m11 * m22 - m12 * m21);
    cm2 inv;
    inv.a11 = dinv * m22; inv.a12 = -(dinv * m12);
    inv.a21 = -(dinv * m21); inv.a22 = dinv * m11;
    q = mmul(inv, cn[3]);                               // * tu[i+1]
    if (i == 1) {
      g.a11 = ep * q.a11; g.a12 = ep * q.a12;           // e[1]*q (e diagonal)
      g.a21 = es * q.a21; g.a22 = es * q.a22;
    } else {
      cm2 ge;                                           // g[i-1]*e[i]
      ge.a11 = g.a11 * ep; ge.a12 = g.a12 * es;
      ge.a21 = g.a21 * ep; ge.a22 = g.a22 * es;
      g = mmul(ge, q);
    }
  }
  return mmul(h2, g);
}

struct RfSpecConsts {
  double dw;        // 2*pi*fsamp/nsamp
  double wref;      // 2*pi*fref, fref = 1 Hz (synrf.cpp:25)
  double a;         // Gauss parameter
  double tshift;
  double qn;        // sqrt(pi)*fsamp/a
  double u;         // slowness s/km
  int waveno;       // 0 P, 1 SV
  int nsamp;
};

// Spectral receiver function value at frequency index j (compute_rf,
// greens.cpp:365-395).  dm = decomposition constants, dec_on its enable flag.
BH_HD cd rf_spectral_value(const cm2& t, const RfSpecConsts& k, const double* dm, bool dec_on, int j) {
  cd cr, cz;
  if (k.waveno == 1) { cr = t.a12; cz = t.a22; }    // greens.cpp:579-581
  else               { cr = t.a11; cz = t.a21; }    // greens.cpp:576-578
  if (dec_on) {
    cd x = dm[0] * cz + dm[1] * cr;
    cd y = dm[2] * cz + dm[3] * cr;
    cz = x; cr = y;
  }
  if (k.waveno == 1) { cd tmp = cz; cz = cr; cr = tmp; }   // :369-373
  double w = k.dw * j;
  double denom = cnorm(cz);
  cd num = cr * cconj(cz);
  double dinv = 1.0 / denom;
  cd crf = mk(num.re * dinv, num.im * dinv);
  double wa = w / k.a;
  wa = (wa > 50.0) ? 50.0 : wa;
  cd cq = k.qn * cexp_d(mk(-0.25 * (wa * wa), -w * k.tshift));
  return crf * cq;
}

}  // namespace bh
