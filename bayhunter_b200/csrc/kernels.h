// kernels.h -- internal launcher interface between engine.cu and the kernel
// translation units.  Not part of the public ABI (see include/bayhunter_b200.h).
#pragma once
#include <cuda_runtime.h>
#include "bh_common.cuh"
#include "rf_core.cuh"
#include "swd_core.cuh"

namespace bh {

constexpr int kMaxTargets = 8;

// ---- per-target device-side description --------------------------------
struct TargetDev {
  int ref;          // BH_REF_*
  int n;            // observed samples
  int cov;          // BH_COV_*
  int synth_off;    // offset of this target inside a synth row
  const double* x;  // [n] device: periods / time axis
  const double* y;  // [n] device: observed
  const double* serr;       // [n] device: yerr / min(yerr) (WHITE_SCALED) or null
  const double* corr_inv;   // [n*n] device (GAUSS) or null
  double logcorr_det;       // slogdet(R) (GAUSS)
  double log_serr_prod;     // log(prod(serr)) (WHITE_SCALED)
  // SWD
  int wave, igr, kmax;      // kmax = min(n, 60) periods actually searched
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
// explicit single-model arrays for the synrf shim (z, vp, vs, rho, qp, qs as given)
void launch_prepare_rf_explicit(const double* z, const double* vp, const double* vs,
                                const double* rho, const double* qp, const double* qs, int nlay,
                                double p, double nsv, double sigma, PrepOut out, cudaStream_t st);

// ---- surface-wave dispersion ----------------------------------------------
struct SwdLaunch {
  const LayerRow* rows;         // [B][row_stride] REAL*4 rows (d, vp, vs, rho)
  int row_stride;
  int lcap;                     // layer capacity of the shared-memory records (>= max nlay)
  const int* nlay;
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
  unsigned long long* counters; // [2] consumed / evaluated secular values
  int spw[kMaxTargets];         // searches per warp of each curve, 1..32
  int warp_begin[kMaxTargets + 1];  // filled by launch_swd: first warp (= CTA) of each curve
  int max_spec;                 // speculative bracket candidates per search per round
};
void launch_swd(SwdLaunch& p, cudaStream_t st);
size_t swd_smem_bytes(int lcap, int S);

// ---- receiver function ----------------------------------------------------
struct RfLaunch {
  const RfLayer* lay;
  const cm2* coef;
  const double* mc;
  const int* nlay;
  int B, lmax;
  RfSpecConsts k;
  cd* spec;          // [B][nfreq]
  double* out;       // [B][out_stride] time series written at out_off, first ndata samples
  int out_stride, out_off, ndata;
  int* tstatus;      // [B][kMaxTargets]
  int target_id;
};
void launch_rf_spectrum(const RfLaunch& p, cudaStream_t st);
void launch_rf_synth(const RfLaunch& p, cudaStream_t st);

// ---- likelihood -------------------------------------------------------------
struct LoglikLaunch {
  TargetSet ts;
  const double* curves;  // SWD searched curves [B][curve_stride]
  int curve_stride;
  int curve_off[kMaxTargets];
  const double* rfsynth; // [B][synth_stride] RF traces at synth_off
  const int* tstatus;    // [B][kMaxTargets]
  const double* noise;   // [B][2T]
  int B;
  double* logL;          // [B]
  double* misfits;       // [B][T+1]
  int* status;           // [B]
  double* synth;         // [B][synth_stride] or null
};
void launch_loglik(const LoglikLaunch& p, cudaStream_t st);

}  // namespace bh
