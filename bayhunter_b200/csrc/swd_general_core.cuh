// swd_general_core.cuh -- per-(model, curve) walk of SURF96's full loop nest (higher
// modes, earth flattening, water layer); see swd_general.cu.  BH_HD so that the
// CPU test-suite can run the very same code against the oracle (tests/host_sim).
#pragma once
#include "swd_core.cuh"

namespace bh {

// sphere(ifunc, 0) followed by sphere(ifunc, 1) (:486-553) on the rows of one model:
// thickness, vp, vs through REAL*8 intermediates stored REAL*4; density by the
// REAL*4 power btp**(-5) (Love) or btp**(-2.275) (Rayleigh).
BH_HD void sphere_rows(LayerRow* r, int n, int wave) {
  const double ar = 6370.0;
  double dr = 0.0, r0 = ar;
  r[n - 1].x = 1.0f;
  for (int i = 0; i < n; ++i) {
    dr = dr + (double)r[i].x;
    const double r1 = ar - dr;
    const double z0 = ar * log(ar / r0);
    const double z1 = ar * log(ar / r1);
    r[i].x = (float)(z1 - z0);
    const double tmp = (ar + ar) / (r0 + r1);
    r[i].y = (float)((double)r[i].y * tmp);
    r[i].z = (float)((double)r[i].z * tmp);
    const float btp = (float)tmp;
    if (wave == 1) {
      // integer power, expanded like the compiler's powi chain: x^5 = (x^2 x) x^2, then 1/x^5
      const float x2 = fmul(btp, btp), x3 = fmul(x2, btp), x5 = fmul(x3, x2);
      r[i].w = fmul(r[i].w, fdiv(1.0f, x5));
    } else {
      // REAL*4 pow: the double result rounded once equals a correctly rounded powf
      r[i].w = fmul(r[i].w, (float)pow((double)btp, (double)-2.275f));
    }
    r0 = r1;
  }
  r[n - 1].x = 0.0f;
}

BH_HD double secular_general(int wave, const LayerRow* rows, int L, int ltop,
                                                  double wvno, double omega) {
  return wave == 1 ? secular_love_reforder(rows, 1, L, wvno, omega, ltop)
                   : secular_rayleigh_reforder(rows, 1, L, wvno, omega, ltop);
}

// One dispersion curve of one model.  rows_in: REAL*4 rows (d, vp, vs, rho), flat-earth
// values.  Writes cg[0..kmax); returns err like surfdisp96 (1: no fundamental root).
// *nsec_out += secular evaluations.
BH_HD int swd_general_curve(const LayerRow* rows_in, int row_step, int L, int wave, int igr, int kmax,
                            int mode, int flsph, const double* periods, double* cg,
                            unsigned long long* nsec_out) {
  LayerRow rows[SWD_MAX_LAYERS];
  for (int i = 0; i < L; ++i) rows[i] = rows_in[i * row_step];
  const int ltop = (rows[0].z <= 0.0f) ? 1 : 0;          // llw - 1 (:134-135)
  if (flsph == 1) sphere_rows(rows, L, wave);

  // extremal velocities and the start value (:139-156, :197-217)
  float betmx = -1.e20f, betmn = 1.e20f;
  int jmn = 0, jsol = 1;
  for (int i = 0; i < L; ++i) {
    if (rows[i].z > 0.01f && rows[i].z < betmn) { betmn = rows[i].z; jmn = i; jsol = 1; }
    else if (rows[i].z <= 0.01f && rows[i].y < betmn) { betmn = rows[i].y; jmn = i; jsol = 0; }
    if (rows[i].z > betmx) betmx = rows[i].z;
  }
  float cc1 = (jsol == 0) ? betmn : halfspace_start(rows[jmn].y, rows[jmn].z);
  cc1 = fmul(.95f, cc1);
  cc1 = fmul(.90f, cc1);

  double tab[SWD_TAB_ROWS];
  double c[SWD_MAX_PERIODS], cb[SWD_MAX_PERIODS];
  for (int k = 0; k < kmax; ++k) { c[k] = 0.0; cb[k] = 0.0; }
  Search s;
  s.tab = tab; s.ts = 1;
  s.cc = (double)cc1;
  s.dc = fabs((double)0.005f);
  s.betmx = (double)betmx;
  s.role = 0; s.k = 0; s.kmax = 1;
  s.cprev = 0.0; s.del1st = 0.0; s.omega = 0.0;
  s.c1 = s.c2 = s.del1 = s.del2 = s.clow = 0.0;
  s.c3 = s.del3 = 0.0; s.nev = 0; s.nctrl = 0; s.m = 0;
  s.ifirst = 0; s.idir = 1;
  s.stage = ST_DONE;
  double root = 0.0, omA = 0.0, omB = 0.0;
  SearchCtx ctx;
  ctx.omA = &omA; ctx.omB = &omB; ctx.ra = &root; ctx.rb = nullptr; ctx.link = nullptr;
  const double one_dc = dmul(1.0e-2, s.dc);              // one*dc (:136)
  const double onea_dc = dmul(1.5, s.dc);                // onea*dc (:198,:204)

  int iq = 1, k = 0, second = 0, ift = 999, err = 0;
  bool finished = false;
  unsigned long long nsec = 0;

  // Advance the loop nest until a search is pending (s.stage = ST_BR_FIRST) or all
  // modes are done.  `from_fail`: enter at label 1700.
  auto advance = [&](bool from_fail) {
    for (;;) {
      if (!from_fail) {
        if (k >= kmax) {                                 // 1600 loop ran out -> next mode
          iq += 1; k = 0;
          if (iq > mode) { finished = true; return; }
        }
        if (k + 1 < ift) {                               // (:230) k >= ift -> 1700
          swd_period_omegas(igr, periods[k], &omA, &omB);
          double c1, clow;
          int ifirst;
          if (k == 0 && iq == 1) { c1 = s.cc; clow = s.cc; ifirst = 1; }                  // :253-256
          else if (k == 0) { c1 = dadd(c[0], one_dc); clow = c1; ifirst = 1; }            // :257-260
          else if (iq > 1) {                                                              // :261-267
            ifirst = 0;
            clow = dadd(c[k], one_dc);
            c1 = c[k - 1];
            if (c1 < clow) c1 = clow;
          } else { ifirst = 0; c1 = dadd(c[k - 1], -onea_dc); clow = s.cc; }              // :268-271
          s.k = 0; s.kmax = 1;
          s.c1 = c1; s.clow = clow; s.ifirst = ifirst; s.omega = omA; s.stage = ST_BR_FIRST;
          second = 0;
          return;
        }
      }
      // label 1700 / 1750 (:313-355)
      from_fail = false;
      if (iq <= 1) err = 1;
      ift = k + 1;
      for (int i = k; i < kmax; ++i) cg[i] = 0.0;
      k = kmax;                                          // falls into "next mode" above
    }
  };
  advance(false);

  while (!finished) {
    const double cand = candidate_from(s.stage, search_pending_c(s), s.idir, s.clow, s.dc, 0);
    const double v = secular_general(wave, rows, L, ltop, s.omega / cand, s.omega);
    nsec += 1;
    search_consume(s, &v, 1, ctx);
    if (s.stage == ST_BR_FIRST || s.stage == ST_BR_STEP || s.stage == ST_RF_TOP || s.stage == ST_RF_POST)
      continue;
    const bool found = s.stage == ST_DONE;
    if (!second) {
      if (!found) { advance(true); continue; }           // :277 iret = -1 -> 1700
      c[k] = root;
      if (igr > 0) {                                     // :282-287
        second = 1;
        s.k = 0; s.kmax = 1;
        s.ifirst = 0;
        s.clow = dadd(cb[k], one_dc);
        s.c1 = dadd(root, -onea_dc);
        s.omega = omB;
        s.stage = ST_BR_FIRST;
        continue;
      }
      cg[k] = swd_curve_value(0, periods[k], c[k], 0.0);
    } else {
      const double c1 = found ? root : c[k];             // :291-294
      cb[k] = c1;
      cg[k] = swd_curve_value(1, periods[k], c[k], c1);
    }
    k += 1;
    advance(false);
  }

  if (nsec_out) *nsec_out += nsec;
  return err;
}

}  // namespace bh
