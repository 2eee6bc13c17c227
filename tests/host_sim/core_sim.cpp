// core_sim.cpp -- TEST INFRASTRUCTURE.  Host-side lock-step simulation of the
// device cores (bayhunter_b200/csrc/{swd,rf}_core.cuh compiled as plain C++).
// It lets the CPU test-suite check the search state machine (with arbitrary,
// even random, speculation widths) and the receiver-function algebra against
// the oracle without a GPU.  Nothing in the product path links this file.
#include <complex>
#include <cstdint>
#include <cstdlib>
#include <vector>

#include "../../include/bayhunter_b200.h"
#include "../../bayhunter_b200/csrc/rf_core.cuh"
#include "../../bayhunter_b200/csrc/sampler_core.cuh"
#include "../../bayhunter_b200/csrc/swd_core.cuh"
#include "../../bayhunter_b200/csrc/swd_general_core.cuh"

using namespace bh;

namespace {
uint32_t lcg(uint32_t& s) { s = s * 1664525u + 1013904223u; return s >> 8; }
}  // namespace

extern "C" {

// The kernel's lane dealing for one round (swd_core.cuh: deal_lanes), all 32 lanes.
void swd_sim_deal(unsigned active, unsigned bracket, int max_spec, int* cnt, int* excl, int* total) {
  for (int lane = 0; lane < 32; ++lane) {
    LaneDeal d = deal_lanes(active, bracket, lane, max_spec);
    cnt[lane] = d.cnt; excl[lane] = d.excl; total[lane] = d.total;
  }
}

// Secular values of n trial phase velocities for one model at one period, in the formulation this
// library was built with (device formulation, or the Fortran's operation order under
// BH_SECULAR_REFERENCE_ORDER).
void swd_sim_secular(const float* rows4, int nlayer, int wave, double omega, int n, const double* c, double* out) {
  std::vector<LayerRow> rows(nlayer);
  for (int i = 0; i < nlayer; ++i) {
    rows[i].x = rows4[4 * i]; rows[i].y = rows4[4 * i + 1];
    rows[i].z = rows4[4 * i + 2]; rows[i].w = rows4[4 * i + 3];
  }
  for (int j = 0; j < n; ++j) out[j] = secular(wave, rows.data(), 1, nlayer, omega / c[j], omega);
}

// One dispersion curve for one model.  rows: (d, vp, vs, rho) REAL*4 x nlayer.
// spec_mode: 0 -> one candidate per round (reference order), k > 0 -> fixed k
// speculative bracket candidates per round, < 0 -> random 1..32 per round.
// counts[0] consumed, counts[1] evaluated secular values.
// A group curve runs as two chains (first roots / second roots) in lock step,
// exactly like two lanes of the kernel's warp.
int swd_sim_curve(const float* rows4, int nlayer, int wave, int igr, int kmax, const double* periods,
                  int spec_mode, unsigned seed, double* cg, long long* counts) {
  std::vector<LayerRow> rows(nlayer);
  for (int i = 0; i < nlayer; ++i) {
    rows[i].x = rows4[4 * i]; rows[i].y = rows4[4 * i + 1];
    rows[i].z = rows4[4 * i + 2]; rows[i].w = rows4[4 * i + 3];
  }
  for (int k = 0; k < kmax; ++k) cg[k] = 0.0;
  double omA[SWD_MAX_PERIODS], omB[SWD_MAX_PERIODS], ra[SWD_MAX_PERIODS], rb[SWD_MAX_PERIODS];
  for (int k = 0; k < kmax; ++k) { swd_period_omegas(igr, periods[k], &omA[k], &omB[k]); ra[k] = rb[k] = 0.0; }
  SearchLink link; link.na = 0; link.a_failed = 0; link.del1st = 0.0;
  SearchCtx ctx; ctx.omA = omA; ctx.omB = omB; ctx.ra = ra; ctx.rb = rb; ctx.link = igr > 0 ? &link : nullptr;
  Search m[2];
  long long consumed = 0, evaluated = 0;
  double tab[2][SWD_TAB_ROWS];
  if (search_setup(m[0], rows.data(), 1, nlayer, kmax, 0, tab[0], 1)) search_begin_a(m[0], ctx);
  else link.a_failed = 1;
  if (igr > 0) search_setup(m[1], rows.data(), 1, nlayer, kmax, 1, tab[1], 1);
  else m[1].stage = ST_DONE;
  uint32_t rng = seed;
  double del[2][32];
  int n[2];
  for (;;) {
    if (igr > 0) search_poll_b(m[1], ctx);
    int any = 0;
    for (int r = 0; r < 2; ++r) {
      Search& s = m[r];
      int nmax = spec_mode == 0 ? 1 : (spec_mode > 0 ? spec_mode : 1 + (int)(lcg(rng) % 32));
      n[r] = search_nwant(s, nmax);
      // a refining chain with lanes to spare also gets the two midpoints its next step may ask for (deal_lanes)
      const bool refining = s.stage == ST_RF_TOP || s.stage == ST_RF_POST;
      if (refining && spec_mode != 0 && (spec_mode > 0 ? spec_mode >= 3 : (lcg(rng) & 1)))
        n[r] = 1 + (spec_mode > 0 ? (spec_mode >= 31 ? kRefineGuesses4 : spec_mode >= 15 ? kRefineGuesses3 : spec_mode >= 7 ? kRefineGuesses2 : kRefineGuesses)
                                  : (lcg(rng) % 4 == 0 ? kRefineGuesses4 : lcg(rng) % 3 == 0 ? kRefineGuesses3 : (lcg(rng) & 1) ? kRefineGuesses2 : kRefineGuesses));
      any += n[r];
      double cpub = search_pending_c(s);
      for (int i = 0; i < n[r]; ++i) {
        double c = candidate_from(s.stage, cpub, s.idir, refining ? s.c1 : s.clow, s.dc, i, s.c2);
        del[r][i] = secular(wave, rows.data(), 1, nlayer, s.omega / c, s.omega);
        ++evaluated;
      }
    }
    if (!any) break;
    for (int r = 0; r < 2; ++r)
      if (n[r] > 0) consumed += search_consume(m[r], del[r], n[r], ctx);
  }
  if (counts) { counts[0] = consumed; counts[1] = evaluated; }
  const bool ok = m[0].stage == ST_DONE && (igr <= 0 || m[1].stage == ST_DONE);
  if (ok)
    for (int k = 0; k < kmax; ++k) cg[k] = swd_curve_value(igr, periods[k], ra[k], rb[k]);
  return ok ? 0 : 1;   // err like surfdisp96
}

// The general dispersion path (higher modes, earth flattening, water layer) exactly as
// one thread of swd_general_kernel runs it.  Returns err like surfdisp96.
int swd_sim_general(const float* rows4, int nlayer, int wave, int igr, int kmax, int mode, int flsph,
                    const double* periods, double* cg, long long* nsec) {
  std::vector<LayerRow> rows(nlayer);
  for (int i = 0; i < nlayer; ++i) {
    rows[i].x = rows4[4 * i]; rows[i].y = rows4[4 * i + 1];
    rows[i].z = rows4[4 * i + 2]; rows[i].w = rows4[4 * i + 3];
  }
  for (int k = 0; k < kmax; ++k) cg[k] = 0.0;
  unsigned long long n = 0;
  int err = swd_general_curve(rows.data(), 1, nlayer, wave, igr, kmax, mode, flsph, periods, cg, &n);
  if (nsec) *nsec = (long long)n;
  return err;
}

// ---- sampler core (sampler_core.cuh): what one thread of the propose / accept kernels runs ----
void sampler_sim_philox(uint32_t* ctr, const uint32_t* key) { philox4x32_10(ctr, key[0], key[1]); }

void sampler_sim_draw(unsigned long long seed, unsigned long long chain, long long iter, double* out4) {
  Draw d = sampler_draw(seed, chain, iter);
  out4[0] = d.u_mod; out4[1] = d.u_idx; out4[2] = d.gauss; out4[3] = d.u_acc;
}

// model: [2*maxl], vs of the k nuclei in the first half, depths in the second; in/out.
// rows: [maxl*4] packed engine rows of the proposal (valid proposals only).  Returns valid.
int sampler_sim_propose(const bh_sampler_config* cfg, int ntargets, long long iiter, const double* propdist,
                        const double* draws4, double* model, int* k, double* vpvs, double* noise,
                        int* modify, double* dvs2, double* rows) {
  SamplerCfg c = sampler_cfg_from_public(*cfg, ntargets);
  const int maxl = c.maxlayers;
  double vs[SMP_MAX_ROWS], z[SMP_MAX_ROWS], h[SMP_MAX_ROWS];
  for (int i = 0; i < *k; ++i) { vs[i] = model[i]; z[i] = model[maxl + i]; }
  Draw d; d.u_mod = draws4[0]; d.u_idx = draws4[1]; d.gauss = draws4[2]; d.u_acc = draws4[3];
  int valid = sampler_propose(c, iiter, propdist, d, vs, z, k, vpvs, noise, modify, dvs2, h);
  if (*k > maxl) valid = 0;
  const int kk = *k < maxl ? *k : maxl;
  for (int i = 0; i < kk; ++i) { model[i] = vs[i]; model[maxl + i] = z[i]; }
  if (valid) sampler_pack_rows(c, vs, h, *k, *vpvs, rows);
  return valid;
}

double sampler_sim_alpha(const bh_sampler_config* cfg, int ntargets, int modify, const double* propdist,
                         double dvs2, double like_prop, double like_cur) {
  SamplerCfg c = sampler_cfg_from_public(*cfg, ntargets);
  return sampler_alpha(c, modify, propdist, dvs2, like_prop, like_cur);
}

void sampler_sim_adjust(const bh_sampler_config* cfg, int ntargets, double* propdist, const long long* accepted,
                        const long long* proposed) {
  SamplerCfg c = sampler_cfg_from_public(*cfg, ntargets);
  sampler_adjust_propdist(c, propdist, accepted, proposed);
}

// Receiver function through rf_core.cuh; same arguments as synrf_cwrap minus fz/fr.
int rf_sim(int nsamp, double fsamp, double tshift, double p, double a, double nsv, double sigma,
           int waveno, int nlay, const double* z, const double* vp, const double* vs,
           const double* rh, const double* qp, const double* qs, double* rf) {
  if (nlay < 2) return 0;
  const double u = p * RF_DEG_PER_KM;
  std::vector<RfLayer> lay(nlay);
  std::vector<cm2> coef(4 * nlay);
  std::vector<double> fvp(nlay), fvs(nlay), frho(nlay);
  const cd zero = mk(0.0, 0.0);
  for (int i = 0; i < nlay; ++i) {
    double h = (i < nlay - 1) ? z[i + 1] - z[i] : -1.0;
    double hf;
    rf_flatten(z[i], h, vp[i], vs[i], rh[i], &hf, &fvp[i], &fvs[i], &frho[i]);
    lay[i].h = hf; lay[i].vp = fvp[i]; lay[i].vs = fvs[i]; lay[i].rho = frho[i];
    lay[i].cqp = 1.0 / (RF_PI * qp[i]); lay[i].bqp = 1.0 / (2.0 * qp[i]);
    lay[i].cqs = 1.0 / (RF_PI * qs[i]); lay[i].bqs = 1.0 / (2.0 * qs[i]);
    cm2 rd, td, ru, tu;
    rd.a11 = rd.a12 = rd.a21 = rd.a22 = zero; td = rd; ru = rd; tu = rd;
    if (i == 0) rf_coeff_surface(u, fvp[0], fvs[0], &ru);
    else rf_coeff_interface(u, fvp[i - 1], fvs[i - 1], frho[i - 1], fvp[i], fvs[i], frho[i], &rd, &td, &ru, &tu);
    coef[4 * i] = rd; coef[4 * i + 1] = td; coef[4 * i + 2] = ru; coef[4 * i + 3] = tu;
  }
  cm2 h2;
  rf_displacement2(u, fvp[0], fvs[0], &h2);
  double dm[4];
  bool dec_on = rf_decomp_consts(u, nsv, sigma, dm);
  RfSpecConsts k;
  k.dw = 2.0 * RF_PI * fsamp / nsamp; k.wref = 2.0 * RF_PI; k.a = a; k.tshift = tshift;
  k.qn = sqrt(RF_PI) * fsamp / a; k.u = u; k.waveno = waveno; k.nsamp = nsamp;
  const int nfreq = nsamp / 2 + 1;
  std::vector<std::complex<double>> x(nsamp);
  for (int j = 0; j < nfreq; ++j) {
    double w = k.dw * j;
    double lgw = j ? log(w / k.wref) : 0.0;
    cm2 t = rf_transfer(lay.data(), coef.data(), h2, nlay, u * u, w, lgw);
    cd v = rf_spectral_value(t, k, dm, dec_on, j);
    x[j] = std::complex<double>(v.re, v.im);
  }
  for (int i = nfreq; i < nsamp; ++i) x[i] = std::conj(x[nsamp - i]);
  // plain O(N log N) radix-2 inverse transform (sign +1), scale 1/N
  int logn = 0;
  while ((1 << logn) < nsamp) ++logn;
  std::vector<std::complex<double>> y(nsamp);
  for (int i = 0; i < nsamp; ++i) {
    int r = 0;
    for (int bit = 0; bit < logn; ++bit) if (i & (1 << bit)) r |= 1 << (logn - 1 - bit);
    y[r] = x[i];
  }
  for (int l = 1; l < nsamp; l <<= 1)
    for (int m = 0; m < l; ++m) {
      std::complex<double> wv = std::polar(1.0, RF_PI * m / l);
      for (int i = m; i < nsamp; i += 2 * l) {
        std::complex<double> bb = wv * y[i + l];
        y[i + l] = y[i] - bb;
        y[i] += bb;
      }
    }
  for (int i = 0; i < nsamp; ++i) rf[i] = y[i].real() / nsamp;
  return 1;
}

}  // extern "C"
