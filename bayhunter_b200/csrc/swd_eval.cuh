// swd_eval.cuh -- the secular functions of swd_core.cuh in two passes per evaluation.
//
// A secular value is a product over layers, e <- e * C(layer), carried from the half-space to
// the surface (surfdisp96.f:813-849 Rayleigh, :735-765 Love): a serial chain.  But most of a
// layer's arithmetic -- var's square roots, exponentials and sines (:929-968), ~100 of the ~175
// fp64 instructions of a Rayleigh layer, ~50 of the 59 of a Love layer -- depends only on the
// trial wavenumber and the layer, not on the propagated vector.  The single-pass formulation
// (secular_*_rec) leaves those hidden inside the serial loop, two streams at a time, and a warp
// on its own keeps the fp64 pipe (one warp instruction per 2 cycles, 8 cycles latency) about a
// quarter busy.  The batch has too few chains to cover that with more warps (~83 chains per SM
// sub-partition at 8192 models), so the parallelism has to come from inside an evaluation:
//
//   pass 1  all vector-independent terms of ALL layers, NL layers (2 NL square-root / exp /
//           sincos streams) side by side in one basic block; results go to a lane-private
//           column of shared memory (7 doubles per Rayleigh layer, 3 per Love layer);
//   pass 2  the serial chain, reading those terms back: ~75 (Rayleigh) / 9 (Love) fp64
//           instructions per layer, of which only ~45 hang on the propagated vector.
//
// Same operations as secular_*_rec in the same order per value: results are bit-identical.
#pragma once
#include "swd_core.cuh"

namespace bh {

constexpr int SWD_HT_SLOTS = 7;     // doubles per layer and lane in the staging buffer (Love uses 3)

#if defined(__CUDACC__)
// ht: this lane's column of the staging buffer; value v of layer l at ht[(l * SWD_HT_SLOTS + v) * hs]
// (hs = 32 on the device: 32 lanes hit 32 banks).  Layers [0, L): L-1 is the half-space, whose
// slot holds its two square roots.
template <int NL>
__device__ __forceinline__ double secular_rayleigh_2pass(const double* __restrict__ rec, int fs, int ls, int L,
                                                         double wvno, double omga, double* __restrict__ ht,
                                                         int hs) {
  constexpr int N = 2 * NL;
  // P halves (even items) take their sincos only when some lane is oscillatory there
  constexpr unsigned kCond = BH_P_SINCOS_COND ? (0x55555555u & ((1u << N) - 1u)) : 0u;
  const double omega = (omga < 1.0e-4) ? 1.0e-4 : omga;
  const double iomega = fm::rcp(omega);
  const double iomega2 = iomega * iomega;
  const double wvno2 = wvno * wvno;
  // ---- pass 1 ----
#pragma unroll 1
  for (int l0 = L - 1; l0 >= 0; l0 -= NL) {
    double xk[N], d[N];
    HalfTerms h[N];
#pragma unroll
    for (int j = 0; j < NL; ++j) {
      const int l = max(l0 - j, 0);                     // a short last group recomputes layer 0
      const double* r = rec + l * ls;
      xk[2 * j] = omega * r[RR_IA * fs]; xk[2 * j + 1] = omega * r[RR_IB * fs];
      d[2 * j] = d[2 * j + 1] = r[RR_D * fs];
    }
    half_terms_n<N, kCond>(wvno, xk, d, h);
#pragma unroll
    for (int j = 0; j < NL; ++j) {
      const int l = max(l0 - j, 0);
      double* o = ht + l * SWD_HT_SLOTS * hs;
      if (j == 0 && l0 == L - 1) {
        // half-space: sqrt of the radicands, 0 when k == omega/v exactly (sqrt_n)
        o[0] = (wvno == xk[0]) ? 0.0 : h[0].r;
        o[hs] = (wvno == xk[1]) ? 0.0 : h[1].r;
      } else {
        const HalfTerms& P = h[2 * j];
        const HalfTerms& S = h[2 * j + 1];
        o[0] = P.cs; o[hs] = P.sn_over_r; o[2 * hs] = P.r_sn;
        o[3 * hs] = S.cs; o[4 * hs] = S.sn_over_r; o[5 * hs] = S.r_sn;
        o[6 * hs] = dunkin_a0(P, S);
      }
    }
  }
  // ---- pass 2 ----
  double e0, e1, e2, e3, e4;
  {
    const double* r = rec + (L - 1) * ls;
    const double* o = ht + (L - 1) * SWD_HT_SLOTS * hs;
    const double rho1 = r[RR_RHO * fs];
    const double ra = o[0], rb = o[hs];
    const double gammk = r[RR_TB2 * fs] * iomega2;
    const double gam = gammk * wvno2;
    const double gamm1 = gam - 1.0;
    e0 = rho1 * rho1 * (gamm1 * gamm1 - gam * gammk * ra * rb);
    e1 = -rho1 * ra;
    e2 = rho1 * (gamm1 - gammk * ra * rb);
    e3 = rho1 * rb;
    e4 = wvno2 - ra * rb;
  }
  if (L < 2) return e0;
#pragma unroll 1
  for (int l = L - 2; l >= 0; --l) {
    const double* r = rec + l * ls;
    const double* o = ht + l * SWD_HT_SLOTS * hs;
    DunkinTerms t;
    t.cosp = o[0]; t.w = o[hs]; t.x = o[2 * hs]; t.cosq = o[3 * hs]; t.y = o[4 * hs]; t.z = o[5 * hs];
    t.a0 = o[6 * hs];
    dunkin_apply_terms(r[RR_RHO * fs], r[RR_IRHO * fs], r[RR_TB2 * fs] * iomega2, t, wvno2, e0, e1, e2, e3, e4);
  }
  double t1 = fm::absmax(fm::absmax(fm::absmax(e0, e1), fm::absmax(e2, e3)), e4);
  if (t1 < 1.0e-40) t1 = 1.0;
  return e0 / t1;                            // IEEE division: exact +-1.0 when saturated
}

template <int N>
__device__ __forceinline__ double secular_love_2pass(const double* __restrict__ rec, int fs, int ls, int L,
                                                     double wvno, double omega, double* __restrict__ ht, int hs) {
  // ---- pass 1: N layers at a time, the half-space first ----
#pragma unroll 1
  for (int l0 = L - 1; l0 >= 0; l0 -= N) {
    double xk[N], d[N];
    HalfTerms h[N];
#pragma unroll
    for (int j = 0; j < N; ++j) {
      const double* r = rec + max(l0 - j, 0) * ls;
      xk[j] = omega * r[LR_IB * fs];
      d[j] = r[LR_D * fs];                               // half-space row: holds rho; its exp / sincos are not used
    }
    half_terms_n<N, 0u>(wvno, xk, d, h);
#pragma unroll
    for (int j = 0; j < N; ++j) {
      const int l = max(l0 - j, 0);
      const double* r = rec + l * ls;
      double* o = ht + l * SWD_HT_SLOTS * hs;
      if (j == 0 && l0 == L - 1) {
        o[0] = (wvno == xk[0]) ? 0.0 : h[0].r;
      } else {
        o[0] = h[j].cs;
        o[hs] = h[j].sn_over_r * r[LR_IMU * fs];
        o[2 * hs] = r[LR_MU * fs] * h[j].r_sn;
      }
    }
  }
  // ---- pass 2 ----
  const double* rh = rec + (L - 1) * ls;
  const double ib = rh[LR_IB * fs];
  double e1 = rh[LR_D * fs] * ht[(L - 1) * SWD_HT_SLOTS * hs];     // rho * rb
  double e2 = ib * ib;
  if (L < 2) return e1;
#pragma unroll 1
  for (int l = L - 2; l >= 0; --l) {
    const double* o = ht + l * SWD_HT_SLOTS * hs;
    LoveLayer m;
    m.cs = o[0]; m.y_over_mu = o[hs]; m.mu_z = o[2 * hs];
    love_apply(m, e1, e2);
  }
  double xnor = fm::absmax(e1, e2);
  if (xnor < 1.0e-40) xnor = 1.0;
  return e1 / xnor;                          // IEEE division: exact +-1.0 when saturated
}

// ---- rotated single-pass forms: the half terms of the NEXT layer are formed in the same basic block as the
// propagation through the CURRENT one (they are independent), so the serial chain of an evaluation is
//   max(half terms, propagation) per layer instead of their sum.
__device__ __forceinline__ void rayleigh_terms_of(const double* __restrict__ r, int fs, double wvno, double omega,
                                                  DunkinTerms& t) {
  double xk[2] = {omega * r[RR_IA * fs], omega * r[RR_IB * fs]}, dd[2] = {r[RR_D * fs], r[RR_D * fs]};
  HalfTerms h[2];
  half_terms_n<2>(wvno, xk, dd, h);
  t.cosp = h[0].cs; t.w = h[0].sn_over_r; t.x = h[0].r_sn;
  t.cosq = h[1].cs; t.y = h[1].sn_over_r; t.z = h[1].r_sn;
  t.a0 = dunkin_a0(h[0], h[1]);
}

__device__ __forceinline__ double secular_rayleigh_rot(const double* __restrict__ rec, int fs, int ls, int L,
                                                       double wvno, double omga) {
  const double omega = (omga < 1.0e-4) ? 1.0e-4 : omga;
  const double iomega = fm::rcp(omega);
  const double iomega2 = iomega * iomega;
  const double wvno2 = wvno * wvno;
  double e0, e1, e2, e3, e4;
  DunkinTerms cur;
  {
    const double* hs = rec + (L - 1) * ls;
    if (L >= 2) rayleigh_terms_of(rec + (L - 2) * ls, fs, wvno, omega, cur);
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
#pragma unroll 1
  for (int l = L - 2; l > 0; --l) {
    const double* r = rec + l * ls;
    DunkinTerms nxt;
    rayleigh_terms_of(r - ls, fs, wvno, omega, nxt);
    dunkin_apply_terms(r[RR_RHO * fs], r[RR_IRHO * fs], r[RR_TB2 * fs] * iomega2, cur, wvno2, e0, e1, e2, e3, e4);
    cur = nxt;
  }
  dunkin_apply_terms(rec[RR_RHO * fs], rec[RR_IRHO * fs], rec[RR_TB2 * fs] * iomega2, cur, wvno2, e0, e1, e2, e3, e4);
  double t1 = fm::absmax(fm::absmax(fm::absmax(e0, e1), fm::absmax(e2, e3)), e4);
  if (t1 < 1.0e-40) t1 = 1.0;
  return e0 / t1;
}

// fully unrolled single pass for a compile-time layer count (what ptxas makes of the whole evaluation as one block)
template <int LL>
__device__ __forceinline__ double secular_rayleigh_unrolled(const double* __restrict__ rec, int fs, int ls,
                                                            double wvno, double omga) {
  const double omega = (omga < 1.0e-4) ? 1.0e-4 : omga;
  const double iomega = fm::rcp(omega);
  const double iomega2 = iomega * iomega;
  const double wvno2 = wvno * wvno;
  double e0, e1, e2, e3, e4;
  {
    const double* hs = rec + (LL - 1) * ls;
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
#pragma unroll
  for (int l = LL - 2; l >= 0; --l) {
    const double* r = rec + l * ls;
    DunkinTerms t;
    rayleigh_terms_of(r, fs, wvno, omega, t);
    dunkin_apply_terms(r[RR_RHO * fs], r[RR_IRHO * fs], r[RR_TB2 * fs] * iomega2, t, wvno2, e0, e1, e2, e3, e4);
  }
  double t1 = fm::absmax(fm::absmax(fm::absmax(e0, e1), fm::absmax(e2, e3)), e4);
  if (t1 < 1.0e-40) t1 = 1.0;
  return e0 / t1;
}

// Love, G layers per iteration (the half-space rides along as the first item): G half-term streams side by side
template <int G>
__device__ __forceinline__ double secular_love_grp(const double* __restrict__ rec, int fs, int ls, int L,
                                                   double wvno, double omega) {
  double e1 = 0.0, e2 = 0.0;
#pragma unroll 1
  for (int l0 = L - 1; l0 >= 0; l0 -= G) {
    double xk[G], d[G];
    HalfTerms h[G];
#pragma unroll
    for (int j = 0; j < G; ++j) {
      const double* r = rec + max(l0 - j, 0) * ls;
      xk[j] = omega * r[LR_IB * fs];
      d[j] = r[LR_D * fs];
    }
    half_terms_n<G, 0u>(wvno, xk, d, h);
#pragma unroll
    for (int j = 0; j < G; ++j) {
      const int l = l0 - j;
      const double* r = rec + max(l, 0) * ls;
      if (j == 0 && l0 == L - 1) {
        const double ib = r[LR_IB * fs];
        const double rb = (wvno == xk[0]) ? 0.0 : h[0].r;
        e1 = r[LR_D * fs] * rb;
        e2 = ib * ib;
      } else if (l >= 0) {
        LoveLayer m;
        m.cs = h[j].cs; m.y_over_mu = h[j].sn_over_r * r[LR_IMU * fs]; m.mu_z = r[LR_MU * fs] * h[j].r_sn;
        love_apply(m, e1, e2);
      }
    }
  }
  if (L < 2) return e1;
  double xnor = fm::absmax(e1, e2);
  if (xnor < 1.0e-40) xnor = 1.0;
  return e1 / xnor;
}
// ---- two trial velocities of ONE chain side by side (same model column, same omega): the record loads and the
// omega-only terms are shared, and the two evaluations are independent instruction streams in one basic block.
__device__ __forceinline__ void secular_rayleigh_rec2(const double* __restrict__ rec, int fs, int ls, int L,
                                                      const double wv[2], double omga, double out[2]) {
  const double omega = (omga < 1.0e-4) ? 1.0e-4 : omga;
  const double iomega = fm::rcp(omega);
  const double iomega2 = iomega * iomega;
  const double k2[2] = {wv[0] * wv[0], wv[1] * wv[1]};
  double e[2][5];
  {
    const double* hs = rec + (L - 1) * ls;
    const double rho1 = hs[RR_RHO * fs];
    const double xka = omega * hs[RR_IA * fs], xkb = omega * hs[RR_IB * fs];
    double sr[4] = {(wv[0] + xka) * fabs(wv[0] - xka), (wv[0] + xkb) * fabs(wv[0] - xkb),
                    (wv[1] + xka) * fabs(wv[1] - xka), (wv[1] + xkb) * fabs(wv[1] - xkb)}, rr[4];
    sqrt_n<4>(sr, rr);
    const double gammk = hs[RR_TB2 * fs] * iomega2;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const double ra = rr[2 * c], rb = rr[2 * c + 1];
      const double gam = gammk * k2[c];
      const double gamm1 = gam - 1.0;
      e[c][0] = rho1 * rho1 * (gamm1 * gamm1 - gam * gammk * ra * rb);
      e[c][1] = -rho1 * ra;
      e[c][2] = rho1 * (gamm1 - gammk * ra * rb);
      e[c][3] = rho1 * rb;
      e[c][4] = k2[c] - ra * rb;
    }
  }
  if (L >= 2) {
#pragma unroll 1
    for (int l = L - 2; l >= 0; --l) {
      const double* r = rec + l * ls;
      const double xa = omega * r[RR_IA * fs], xb = omega * r[RR_IB * fs], dl = r[RR_D * fs];
      const double kk[4] = {wv[0], wv[0], wv[1], wv[1]};
      const double xk[4] = {xa, xb, xa, xb}, dd[4] = {dl, dl, dl, dl};
      HalfTerms h[4];
      half_terms_nk<4, BH_P_SINCOS_COND ? 5u : 0u>(kk, xk, dd, h);
      const double rho = r[RR_RHO * fs], ri = r[RR_IRHO * fs], gk = r[RR_TB2 * fs] * iomega2;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        DunkinTerms t;
        t.cosp = h[2 * c].cs; t.w = h[2 * c].sn_over_r; t.x = h[2 * c].r_sn;
        t.cosq = h[2 * c + 1].cs; t.y = h[2 * c + 1].sn_over_r; t.z = h[2 * c + 1].r_sn;
        t.a0 = dunkin_a0(h[2 * c], h[2 * c + 1]);
        dunkin_apply_terms(rho, ri, gk, t, k2[c], e[c][0], e[c][1], e[c][2], e[c][3], e[c][4]);
      }
    }
  }
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    if (L < 2) { out[c] = e[c][0]; continue; }
    double t1 = fm::absmax(fm::absmax(fm::absmax(e[c][0], e[c][1]), fm::absmax(e[c][2], e[c][3])), e[c][4]);
    if (t1 < 1.0e-40) t1 = 1.0;
    out[c] = e[c][0] / t1;
  }
}

__device__ __forceinline__ void secular_love_rec2(const double* __restrict__ rec, int fs, int ls, int L,
                                                  const double wv[2], double omega, double out[2]) {
  const double* hs = rec + (L - 1) * ls;
  const double ib = hs[LR_IB * fs];
  const double xkb = omega * ib;
  double srb[2] = {(wv[0] + xkb) * fabs(wv[0] - xkb), (wv[1] + xkb) * fabs(wv[1] - xkb)}, rb[2];
  sqrt_n<2>(srb, rb);
  double e1[2] = {hs[LR_D * fs] * rb[0], hs[LR_D * fs] * rb[1]};
  double e2[2] = {ib * ib, ib * ib};
  if (L >= 2) {
#pragma unroll 1
    for (int l = L - 2; l >= 0; --l) {
      const double* r = rec + l * ls;
      const double xb = omega * r[LR_IB * fs], dl = r[LR_D * fs];
      const double xk[2] = {xb, xb}, dd[2] = {dl, dl};
      HalfTerms q[2];
      half_terms_nk<2, 0u>(wv, xk, dd, q);
      const double imu = r[LR_IMU * fs], mu = r[LR_MU * fs];
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        LoveLayer m;
        m.cs = q[c].cs; m.y_over_mu = q[c].sn_over_r * imu; m.mu_z = mu * q[c].r_sn;
        love_apply(m, e1[c], e2[c]);
      }
    }
  }
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    if (L < 2) { out[c] = e1[c]; continue; }
    double xnor = fm::absmax(e1[c], e2[c]);
    if (xnor < 1.0e-40) xnor = 1.0;
    out[c] = e1[c] / xnor;
  }
}
#endif  // __CUDACC__


}  // namespace bh
