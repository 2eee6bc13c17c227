// sampler_core.cuh -- one Metropolis-Hastings step of a BayHunter chain, written for a
// lock-step ensemble of chains on the GPU.
//
// Behavioural reference: src/SingleChain.py of BayHunter
//   proposal kinds           _model_vschange :288-293, _model_zvnoi_move :295-300,
//                            _model_layerbirth :246-267, _model_layerdeath :269-286,
//                            _get_hyperparameter_proposal :391-398, _get_vpvs_proposal :407-411
//   ordering                 _sort_modelproposal :317-330
//   prior checks             _validmodel :332-389, _validnoise :400-405, _validvpvs :413-418
//   acceptance               get_acceptance_probability :453-489, iterate :511-563
//   proposal-width control   adjust_propdist :423-451 (every 1000th iteration, :585-587)
//   model adapter            Model.get_vp_vs_h src/Models.py:40-52 (+ get_vp :26-37)
//
// What is NOT the reference's: the random stream.  The reference draws from one
// numpy RandomState (MT19937) per chain; here every (chain, iteration) owns a
// counter-based Philox4x32-10 block, so a step needs no RNG state, is reproducible
// for any batch split / GPU count, and can be replayed.  The four variates of a step
// (`Draw`) are consumed exactly where the reference consumes its draws, which is what
// lets the tests replay the SAME variates through the reference's own SingleChain
// (tests/golden/make_sampler_fixtures.py) and compare step by step.
//
// Arithmetic that feeds comparisons against the priors is written with explicit
// single-rounding helpers (dmul/dadd): numpy never fuses a multiply into an add.
#pragma once
#include "bh_common.cuh"

namespace bh {

constexpr int SMP_MAX_ROWS = 101;      // nuclei per model: priors['layers'][1] + 1 <= 100 (+1 during a birth)
constexpr int SMP_MAX_TARGETS = 8;
constexpr int SMP_NPAR = 5;            // propdist / accepted / proposed entries (PAR_MAP, SingleChain.py:21-22)

enum SamplerModify : int { MOD_VS = 0, MOD_ZV = 1, MOD_BIRTH = 2, MOD_DEATH = 3, MOD_NOISE = 4, MOD_VPVS = 5 };
BH_HD int sampler_paridx(int modify) {         // PAR_MAP
  return modify == MOD_VS ? 0 : modify == MOD_ZV ? 1 : (modify == MOD_BIRTH || modify == MOD_DEATH) ? 2
       : modify == MOD_NOISE ? 3 : 4;
}

struct SamplerCfg {
  int maxlayers;                       // rows of a stored model = priors['layers'][1] + 1
  int ntargets;
  int layers_min, layers_max;          // priors['layers']
  double vs_min, vs_max, z_min, z_max; // priors['vs'], priors['z']
  int vpvs_fixed;                      // priors['vpvs'] is a float
  double vpvs_min, vpvs_max;
  int has_mantle;                      // priors['mantle'] is not None
  double mantle_vs, mantle_vpvs;
  int noise_fixed[2 * SMP_MAX_TARGETS];           // corr_0, sigma_0, corr_1, ...
  double noise_min[2 * SMP_MAX_TARGETS], noise_max[2 * SMP_MAX_TARGETS];
  double thickmin;
  int has_lvz, has_hvz;
  double lvz, hvz;
  double acc_lo, acc_hi;               // initparams['acceptance'] in percent
  int iter_burnin, iter_main;
  unsigned long long seed;
};

// ---------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. 2011), counter = (iteration, chain), key = seed
// ---------------------------------------------------------------------------
BH_HD void philox_round(uint32_t* c, uint32_t k0, uint32_t k1) {
  const uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
  const uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
  const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
  const uint32_t n1 = (uint32_t)p1;
  const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
  const uint32_t n3 = (uint32_t)p0;
  c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}
BH_HD void philox4x32_10(uint32_t* c, uint32_t k0, uint32_t k1) {
  for (int r = 0; r < 10; ++r) {
    philox_round(c, k0, k1);
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}
BH_HD double u01_53(uint32_t hi, uint32_t lo) {          // [0, 1), 53 bits
  const uint64_t v = (((uint64_t)hi << 32) | lo) >> 11;
  return (double)v * (1.0 / 9007199254740992.0);
}

struct Draw {
  double u_mod;   // which modification            (rstate.choice, iterate :511-517)
  double u_idx;   // which nucleus / noise index, or the birth depth in [0, 1)
  double gauss;   // N(0, 1) of the perturbation    (rstate.normal(0, propdist[i]) = propdist[i] * gauss)
  double u_acc;   // acceptance draw in [0, 1)     (:556)
};

BH_HD Draw sampler_draw(unsigned long long seed, unsigned long long chain, long long iter) {
  const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  const unsigned long long it = (unsigned long long)iter;
  uint32_t a[4] = {(uint32_t)it, (uint32_t)(it >> 32), (uint32_t)chain, (uint32_t)(chain >> 32) | 0x00000000u};
  uint32_t b[4] = {a[0], a[1], a[2], a[3] | 0x40000000u};
  uint32_t c[4] = {a[0], a[1], a[2], a[3] | 0x80000000u};
  philox4x32_10(a, k0, k1);
  philox4x32_10(b, k0, k1);
  philox4x32_10(c, k0, k1);
  Draw d;
  d.u_mod = u01_53(a[0], a[1]);
  d.u_idx = u01_53(a[2], a[3]);
  const double u1 = 1.0 - u01_53(b[0], b[1]);            // (0, 1]
  const double u2 = u01_53(b[2], b[3]);
  double sn, cs;
  sincos_d(6.283185307179586 * u2, &sn, &cs);
  d.gauss = sqrt(-2.0 * log(u1)) * cs;                   // Box-Muller
  d.u_acc = u01_53(c[0], c[1]);
  (void)sn;
  return d;
}

// ---------------------------------------------------------------------------
// Model.get_vp_vs_h + the prior checks of _validmodel on k nuclei (vs[], z[] ordered)
// h_out[k] (last = 0).  Returns 1 when the model passes.
// ---------------------------------------------------------------------------
BH_HD void sampler_thickness(const double* z, int k, double* h) {
  double prev = 0.0;
  for (int i = 0; i + 1 < k; ++i) {
    const double zd = (z[i] + z[i + 1]) / 2.0;           // z_disc (Models.py:44)
    h[i] = zd - prev;                                    // h_lay  (:45)
    prev = zd;
  }
  h[k - 1] = 0.0;
}

BH_HD int sampler_validmodel(const SamplerCfg& c, const double* vs, const double* z, int k, double* h) {
  if (k < 1) return 0;
  sampler_thickness(z, k, h);
  const int layers = k - 1;                                              // :344-349
  if (!(layers >= c.layers_min && layers <= c.layers_max)) return 0;
  for (int i = 0; i + 1 < k; ++i) if (h[i] < c.thickmin) return 0;       // :352
  for (int i = 0; i < k; ++i) if (vs[i] < c.vs_min || vs[i] > c.vs_max) return 0;   // :358-363
  double zc = 0.0;                                                       // z = np.cumsum(h) (:368)
  for (int i = 0; i < k; ++i) {
    zc = zc + h[i];
    if (zc < c.z_min || zc > c.z_max) return 0;
  }
  if (c.has_lvz)                                                         // :374-381
    for (int i = 0; i + 1 < k; ++i)
      if (!(dadd(vs[i + 1], -dmul(vs[i], 1.0 - c.lvz)) > 0.0)) return 0;
  if (c.has_hvz)                                                         // :383-389
    for (int i = 0; i + 1 < k; ++i)
      if (!(dadd(dmul(vs[i], 1.0 + c.hvz), -vs[i + 1]) > 0.0)) return 0;
  return 1;
}

// Packed engine rows (vs, vp/vs, z_top, h) of a model: vp = vs * vpvs, or vs * mantle_vpvs from
// the first row with vs >= mantle_vs downwards (Models.py:26-37); z_top = cumsum(h) shifted
// (rfmini_modrf.py:122-123).
BH_HD void sampler_pack_rows(const SamplerCfg& c, const double* vs, const double* h, int k, double vpvs,
                             double* rows /*[k][4]*/) {
  bool mantle = false;
  double ztop = 0.0;
  for (int i = 0; i < k; ++i) {
    if (c.has_mantle && vs[i] >= c.mantle_vs) mantle = true;
    rows[4 * i + 0] = vs[i];
    rows[4 * i + 1] = mantle ? c.mantle_vpvs : vpvs;
    rows[4 * i + 2] = ztop;
    rows[4 * i + 3] = h[i];
    ztop = ztop + h[i];
  }
}

// Stable insertion sort of the nuclei by depth (_sort_modelproposal).
BH_HD void sampler_sort(double* vs, double* z, int k) {
  for (int i = 1; i < k; ++i) {
    const double zi = z[i], vi = vs[i];
    int j = i - 1;
    while (j >= 0 && z[j] > zi) { z[j + 1] = z[j]; vs[j + 1] = vs[j]; --j; }
    z[j + 1] = zi; vs[j + 1] = vi;
  }
}

BH_HD int sampler_argmin_abs(const double* z, int k, double z0) {       // np.argmin(abs(z - z0)): first minimum
  int ind = 0;
  double best = fabs(z[0] - z0);
  for (int i = 1; i < k; ++i) {
    const double d = fabs(z[i] - z0);
    if (d < best) { best = d; ind = i; }
  }
  return ind;
}

// ---------------------------------------------------------------------------
// Proposal of one chain (iterate :511-547).  vs/z: in = current nuclei (k of them, ordered),
// out = proposed nuclei (*pk of them).  noise/vpvs likewise.  Returns 1 if the proposal passes
// its prior check; h (thicknesses of the proposal) is valid then.
// ---------------------------------------------------------------------------
BH_HD int sampler_propose(const SamplerCfg& c, long long iiter, const double* propdist, const Draw& d,
                          double* vs, double* z, int* pk, double* vpvs, double* noise,
                          int* modify_out, double* dvs2_out, double* h) {
  int k = *pk;
  // choice of the modification (:512-517): the first 1 % of all iterations only vs / z moves
  int nfree = 0;
  for (int i = 0; i < 2 * c.ntargets; ++i) nfree += c.noise_fixed[i] ? 0 : 1;
  const int iterations = c.iter_burnin + c.iter_main;
  const bool early = (double)iiter < ((double)(-c.iter_burnin) + (double)iterations * 0.01);
  int mods[6], nm = 0;
  mods[nm++] = MOD_VS; mods[nm++] = MOD_ZV;
  if (!early) { mods[nm++] = MOD_BIRTH; mods[nm++] = MOD_DEATH; }
  if (nfree > 0) mods[nm++] = MOD_NOISE;
  if (!c.vpvs_fixed) mods[nm++] = MOD_VPVS;
  int mi = (int)(d.u_mod * nm);
  if (mi >= nm) mi = nm - 1;
  const int modify = mods[mi];
  *modify_out = modify;
  *dvs2_out = 0.0;
  int valid = 1;
  if (modify == MOD_VS) {
    int ind = (int)(d.u_idx * k); if (ind >= k) ind = k - 1;             // randint(0, k)
    vs[ind] = dadd(vs[ind], dmul(propdist[0], d.gauss));
  } else if (modify == MOD_ZV) {
    int ind = (int)(d.u_idx * k); if (ind >= k) ind = k - 1;             // randint(k, 2k)
    z[ind] = dadd(z[ind], dmul(propdist[1], d.gauss));
    sampler_sort(vs, z, k);
  } else if (modify == MOD_BIRTH) {
    const double z_birth = dadd(c.z_min, dmul(c.z_max - c.z_min, d.u_idx));   // uniform(zmin, zmax)
    const int ind = sampler_argmin_abs(z, k, z_birth);
    const double vs_before = vs[ind];
    const double vs_birth = dadd(vs_before, dmul(propdist[2], d.gauss));
    z[k] = z_birth; vs[k] = vs_birth;
    k += 1;
    const double dv = vs_birth - vs_before;
    *dvs2_out = dmul(dv, dv);
    sampler_sort(vs, z, k);
  } else if (modify == MOD_DEATH) {
    int ind = (int)(d.u_idx * k); if (ind >= k) ind = k - 1;             // randint(0, k)
    const double z_before = z[ind], vs_before = vs[ind];
    for (int i = ind; i + 1 < k; ++i) { z[i] = z[i + 1]; vs[i] = vs[i + 1]; }
    k -= 1;
    if (k < 1) { valid = 0; }                                            // the reference would raise (argmin of [])
    else {
      const int ia = sampler_argmin_abs(z, k, z_before);
      const double dv = vs[ia] - vs_before;
      *dvs2_out = dmul(dv, dv);
    }
  } else if (modify == MOD_NOISE) {
    int pick = (int)(d.u_idx * nfree); if (pick >= nfree) pick = nfree - 1;   // rstate.choice(noiseinds)
    int ind = 0;
    for (int i = 0; i < 2 * c.ntargets; ++i)
      if (!c.noise_fixed[i]) { if (pick == 0) { ind = i; break; } --pick; }
    noise[ind] = dadd(noise[ind], dmul(propdist[3], d.gauss));
    for (int i = 0; i < 2 * c.ntargets; ++i)                                  // _validnoise
      if (!c.noise_fixed[i] && (noise[i] < c.noise_min[i] || noise[i] > c.noise_max[i])) valid = 0;
  } else {
    *vpvs = dadd(*vpvs, dmul(propdist[4], d.gauss));
    if (*vpvs < c.vpvs_min || *vpvs > c.vpvs_max) valid = 0;                  // _validvpvs
  }
  *pk = k;
  if (modify <= MOD_DEATH) {
    if (valid) valid = sampler_validmodel(c, vs, z, k, h);
  } else {
    sampler_thickness(z, k, h);      // model unchanged: rows are re-packed with the new vpvs / as they were
  }
  return valid;
}

// log acceptance probability (get_acceptance_probability :453-489); dv = vs prior width
BH_HD double sampler_alpha(const SamplerCfg& c, int modify, const double* propdist, double dvs2,
                           double like_prop, double like_cur) {
  const double C = like_prop - like_cur;
  if (modify != MOD_BIRTH && modify != MOD_DEATH) return C;
  const double theta = propdist[2];
  const double dv = c.vs_max - c.vs_min;
  const double s2pi = sqrt(2.0 * 3.141592653589793);
  const double B = dvs2 / dmul(2.0, dmul(theta, theta));
  if (modify == MOD_BIRTH) {
    const double A = dmul(theta, s2pi) / dv;
    return dadd(dadd(log(A), B), C);
  }
  const double A = dv / dmul(theta, s2pi);
  return dadd(dadd(log(A), -B), C);
}

// adjust_propdist (:423-451) with cumulative counters
BH_HD void sampler_adjust_propdist(const SamplerCfg& c, double* propdist, const long long* accepted,
                                   const long long* proposed) {
  for (int i = 0; i < SMP_NPAR; ++i) {
    if (proposed[i] == 0) continue;                      // rate is NaN: not inverted for
    const double rate = (double)accepted[i] / (double)proposed[i] * 100.0;
    if (rate < c.acc_lo) {
      double nw = propdist[i] * 0.95;
      if (nw < 0.001) nw = 0.001;
      propdist[i] = nw;
    } else if (rate > c.acc_hi) {
      propdist[i] = propdist[i] * 1.05;
    }
  }
}

#ifdef BAYHUNTER_B200_H
// public bh_sampler_config (include/bayhunter_b200.h) -> kernel-side configuration
inline SamplerCfg sampler_cfg_from_public(const bh_sampler_config& c, int ntargets) {
  SamplerCfg k;
  k.maxlayers = c.layers_max + 1; k.ntargets = ntargets;
  k.layers_min = c.layers_min; k.layers_max = c.layers_max;
  k.vs_min = c.vs_min; k.vs_max = c.vs_max; k.z_min = c.z_min; k.z_max = c.z_max;
  k.vpvs_fixed = c.vpvs_fixed; k.vpvs_min = c.vpvs_min; k.vpvs_max = c.vpvs_max;
  k.has_mantle = c.has_mantle; k.mantle_vs = c.mantle_vs; k.mantle_vpvs = c.mantle_vpvs;
  for (int i = 0; i < 2 * SMP_MAX_TARGETS; ++i) {
    k.noise_fixed[i] = i < 2 * ntargets ? c.noise_fixed[i] : 1;
    k.noise_min[i] = c.noise_min[i]; k.noise_max[i] = c.noise_max[i];
  }
  k.thickmin = c.thickmin; k.has_lvz = c.has_lvz; k.has_hvz = c.has_hvz; k.lvz = c.lvz; k.hvz = c.hvz;
  k.acc_lo = c.acceptance[0]; k.acc_hi = c.acceptance[1];
  k.iter_burnin = c.iter_burnin; k.iter_main = c.iter_main; k.seed = c.seed;
  return k;
}
#endif

}  // namespace bh
