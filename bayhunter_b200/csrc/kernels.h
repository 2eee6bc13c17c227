// kernels.h -- internal launcher interface between engine.cu and the kernel
// translation units.  Not part of the public ABI (see include/bayhunter_b200.h).
#pragma once
#include <cuda_runtime.h>
#include <stdlib.h>
#include "bh_common.cuh"
#include "rf_core.cuh"
#include "swd_core.cuh"

namespace bh {

constexpr int kMaxTargets = 8;

// Every kernel of the engine asks for the SAME shared-memory carve-out: an SM
// only hosts CTAs of kernels that agree on its L1/shared split, and the engine
// relies on the dispersion (Rayleigh, Love) and receiver-function kernels of one
// evaluation sharing the SMs from three streams.
inline int bh_carveout_pct() {
  static int pct = -2;
  if (pct == -2) {
    const char* e = getenv("BH_CARVEOUT");     // developer override; -1 = driver default
    pct = e ? atoi(e) : 100;                   // all of it shared: the kernels keep no local memory
  }
  return pct;
}
template <class K>
inline void bh_set_carveout(K kernel) {
  cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, bh_carveout_pct());
}

// Function attributes are per device: a process that drives several GPUs (bh_set_device) must set them on each.
// One KernelAttrs per kernel instantiation remembers the device it was last configured on.
struct KernelAttrs { int dev = -1; size_t smem = 0; };
template <class K>
inline void bh_configure_kernel(K kernel, size_t dyn_smem, KernelAttrs& a) {
  int dev = 0;
  cudaGetDevice(&dev);
  if (a.dev != dev) { bh_set_carveout(kernel); a.dev = dev; a.smem = 0; }
  if (dyn_smem > 48 * 1024 && dyn_smem > a.smem) {
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn_smem);
    a.smem = dyn_smem;
  }
}

// ---- per-target device-side description --------------------------------
struct TargetDev {
  int ref;          // BH_REF_*
  int n;            // observed samples
  int cov;          // BH_COV_*
  int synth_off;    // offset of this target inside a synth row
  const double* x;  // [n] device: periods / time axis
  const double* y;  // [n] device: observed
  const double* serr;       // [n] device: yerr / min(yerr) (WHITE_SCALED) or null
  const double* corr_inv;   // [n*n] device (GAUSS): (R^-1 + R^-T) / 2, the part of R^-1 a quadratic form sees; else null
  double logcorr_det;       // slogdet(R) (GAUSS)
  double log_serr_prod;     // log(prod(serr)) (WHITE_SCALED)
  // SWD
  int wave, igr, kmax;      // kmax = min(n, 60) periods actually searched
  int mode, flsph;          // SurfDisp.set_modelparams keys (surf96_modsw.py:28-31)
  const double* periods;    // [kmax] device: x, or the 60-point resampling (n > 60)
  // RF
  int nsamp, waveno;
  double fsamp, tshift, gauss, p, nsv, qp, qs;
};

struct TargetSet {
  int ntargets;
  int synth_stride;
  TargetDev t[kMaxTargets];
};

// ---- scratch produced by the preparation kernel --------------------------
struct PrepOut {
  LayerRow* swd_rows;   // [B][swd_stride] REAL*4 rows (d, vp, vs, rho)
  int swd_stride;       // odd number of rows >= lmax (bank-conflict-free staging)
  RfLayer* rf_lay;      // [B][lmax]
  cm2* rf_coef;         // [B][lmax][4]  (rd, td, ru, tu) per interface
  double* rf_mc;        // [B][16]: h2 (8), decomposition m (4), dec_on, pad
};

// packed model rows (vs, vp/vs, z_top, h) -> SWD rows + RF tables
void launch_prepare(const double* model, const int* nlay, const double* rho, int B, int lmax,
                    bool want_swd, bool want_rf, double rf_p, double rf_nsv, double rf_qp,
                    double rf_qs, PrepOut out, cudaStream_t st);
// perm[B]: model indices ordered by decreasing layer count (counting sort; order within a layer
// count is arbitrary).  *maxn (device, may be null) receives the largest layer count.
void launch_layer_order(const int* nlay, int B, int* perm, int* maxn, cudaStream_t st);
// explicit single-model arrays for the synrf shim (z, vp, vs, rho, qp, qs as given)
void launch_prepare_rf_explicit(const double* z, const double* vp, const double* vs,
                                const double* rho, const double* qp, const double* qs, int nlay,
                                double p, double nsv, double sigma, PrepOut out, cudaStream_t st);

// ---- surface-wave dispersion ----------------------------------------------
struct SwdLaunch {
  const LayerRow* rows;         // [B][row_stride] REAL*4 rows (d, vp, vs, rho)
  int row_stride;
  int lcap;                     // layer capacity of the shared-memory records (>= nlay_hi)
  int nlay_lo, nlay_hi;         // models with nlay in (nlay_lo, nlay_hi] belong to this launch
  const int* nlay;
  const int* perm;              // [B] order in which models are dealt to warps (null: 0, 1, 2, ...)
  int B;
  int ncurves;
  int target_id[kMaxTargets];   // index into TargetSet for each curve
  int wave[kMaxTargets], igr[kMaxTargets], kmax[kMaxTargets], synth_off[kMaxTargets];
  const double* periods[kMaxTargets];
  double* curves;               // [B][curve_stride] searched velocities (kmax per curve)
  double* roots;                // [B][2*curve_stride] scratch: first / second roots of every period
  int curve_stride;
  int curve_off[kMaxTargets];
  int* tstatus;                 // [B][kMaxTargets] 1 ok / 0 failed
  unsigned long long* counters; // [2 + 2*kMaxTargets] consumed / evaluated secular values; per curve sum / max of warp rounds
  int spw[kMaxTargets];         // searches per warp of each curve, 1..32
  int warp_begin[kMaxTargets + 1];  // filled by launch_swd: first warp (= CTA) of each curve
  int max_spec;                 // speculative bracket candidates per search per round
  int counter_base;             // index of this launch's first curve in the per-curve counters
  // SM partition of a mixed (Rayleigh + Love) launch: CTAs on SMs [0, sm_split) prefer the
  // Rayleigh work items, the others the Love items, so that an SM's instruction cache holds
  // one of the two code paths.  queue: [0..1] next item per type, [2..2+nsm) CTAs started per SM.
  int* queue;                   // zeroed before the launch; null = static blockIdx mapping
  int sm_split;
  int type_quota[2];            // CTAs per SM that take the SM's own type before preferring the other
  int type_begin[3];            // filled by launch_swd: first warp of Rayleigh items, Love items, end
  int direct;                   // 1: no speculation (swd_kernel<true>), needs spw = 32 / 16
  int lockstep;                 // 1: swd_lockstep_kernel (group curves: at most 16 models per warp)
  int* done;                    // += 1 per retired warp (null: not counted)
};
void launch_swd(SwdLaunch& p, cudaStream_t st);
// the same search with every lane owning a chain (swd_lockstep.cu): full batches
void launch_swd_lockstep(SwdLaunch& p, cudaStream_t st);
// the same search with the 128 lanes of a CTA as one pool over the chains of M models (swd_pool.cu)
bool swd_pool_fits(const SwdLaunch& p);
int swd_pool_models(const SwdLaunch& p, int lcap, int want);
size_t swd_pool_smem_bytes(int lcap, int M);
int swd_pool_warp_count(const SwdLaunch& p, int M);
void launch_swd_pool(const SwdLaunch& p, int M, cudaStream_t st);
void launch_swd_gate(const int* done, int threshold, cudaStream_t st);
int swd_warp_count(const SwdLaunch& p);
size_t swd_smem_bytes(int lcap, int S);

// Non-default SURF96 branches (higher modes, earth flattening, water layer): one thread
// per (model, curve), see swd_general.cu.
struct SwdGeneralLaunch {
  const LayerRow* rows;         // [B][row_stride] REAL*4 rows (d, vp, vs, rho), flat-earth values
  int row_stride;
  int lcap;
  const int* nlay;
  int B;
  int ncurves;
  int target_id[kMaxTargets];
  int wave[kMaxTargets], igr[kMaxTargets], kmax[kMaxTargets], mode[kMaxTargets], flsph[kMaxTargets];
  const double* periods[kMaxTargets];
  double* curves;               // [B][curve_stride]
  int curve_stride;
  int curve_off[kMaxTargets];
  int* tstatus;                 // [B][kMaxTargets]
  unsigned long long* counters; // [0], [1] += secular evaluations (may be null)
};
void launch_swd_general(const SwdGeneralLaunch& p, cudaStream_t st);

// ---- receiver function ----------------------------------------------------
struct RfLaunch {
  const RfLayer* lay;
  const cm2* coef;
  const double* mc;
  const int* nlay;
  int B, lmax;
  RfSpecConsts k;
  cd* spec;          // [B][nfreq]
  int nact;          // frequencies 0 .. nact-1 are computed; the Gauss filter makes the rest exactly
                     // negligible (engine.cu: rf_active_frequencies), they enter the transform as 0
  double* out;       // [B][out_stride] time series written at out_off, first ndata samples
  int out_stride, out_off, ndata;
  int* tstatus;      // [B][kMaxTargets]
  int target_id;
};
// Number of leading frequency bins whose Gauss weight exp(-(w/2a)^2) is >= floor (0 < floor < 1;
// floor <= 0: all nsamp/2+1 bins).
int rf_active_frequencies(const RfSpecConsts& k, double wfloor);
void launch_rf_spectrum(const RfLaunch& p, cudaStream_t st);
void launch_rf_synth(const RfLaunch& p, cudaStream_t st);

// ---- likelihood -------------------------------------------------------------
struct LoglikLaunch {
  TargetSet ts;
  const double* curves;  // SWD searched curves [B][curve_stride]
  int curve_stride;
  int curve_off[kMaxTargets];
  const double* rfsynth; // [B][synth_stride] RF traces at synth_off
  const double* given;   // [B][synth_stride] modelled data of EVERY target supplied by the caller, or null
  const int* tstatus;    // [B][kMaxTargets]
  const double* noise;   // [B][2T]
  int B;
  double* logL;          // [B]
  double* misfits;       // [B][T+1]
  int* status;           // [B]
  double* synth;         // [B][synth_stride] or null
  double* gauss_phi;     // [B][kMaxTargets] scratch: d^T R^-1 d of the Gauss-law targets
  double* gauss_res;     // [B][32 * tile rows] scratch: residuals of the Gauss-law target being contracted
  double* gauss_part;    // [B][tile rows] scratch: partial sums per tile row
};
int gauss_tile_rows(int n);   // 32 x 32 tiles per row / column of R^-1
void launch_gauss_quadform(const LoglikLaunch& p, int t, cudaStream_t st);   // before launch_loglik, per Gauss-law target
void launch_loglik(const LoglikLaunch& p, cudaStream_t st);

}  // namespace bh
