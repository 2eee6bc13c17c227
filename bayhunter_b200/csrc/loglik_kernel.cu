// loglik_kernel.cu -- Gaussian log-likelihood + RMS misfit of all targets of a
// joint evaluation, one warp per model, warp-shuffle reductions.
//
// Behavioural reference: BayHunter src/Targets.py
//   JointTarget.evaluate :314-347 (sentinels -1e15 / 1e15 on an invalid target)
//   Valuation.get_rms :99-103, get_covariance_nocorr :105-115,
//   get_covariance_nocorr_scalederr :117-129, get_corr_inv/get_covariance_exp
//   :131-148, get_covariance_gauss :162-173
//   SurfDisp > 60 periods: 60-point resampling + np.interp
//   (src/surf96_modsw.py:35-43,106-122)
// The reference builds dense n x n inverse covariance matrices per call; here
// the quadratic forms are evaluated in closed form (tridiagonal / diagonal) and
// only the Gauss law reads the dense R^-1 uploaded at engine creation.
#include "kernels.h"
#include "../../include/bayhunter_b200.h"

namespace bh {

namespace {

constexpr int kWarps = 4;
constexpr double kLog2Pi = 1.8378770664093454835606594728112;

// numpy.interp semantics on an increasing grid xp[0..m)
__device__ __forceinline__ double interp_np(double x, const double* __restrict__ xp,
                                            const double* __restrict__ fp, int m) {
  if (x <= xp[0]) return fp[0];
  if (x >= xp[m - 1]) return fp[m - 1];
  int lo = 0, hi = m - 1;
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (xp[mid] <= x) lo = mid; else hi = mid;
  }
  double slope = (fp[lo + 1] - fp[lo]) / (xp[lo + 1] - xp[lo]);
  return slope * (x - xp[lo]) + fp[lo];
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  return v;
}

__device__ __forceinline__ double modelled(const TargetDev& T, const double* __restrict__ curve,
                                           const double* __restrict__ trace, int i) {
  if (T.ref <= BH_REF_LDISPGR) {
    if (T.n == T.kmax) return curve[i];
    return interp_np(T.x[i], T.periods, curve, T.kmax);
  }
  return trace[i];
}

__global__ void __launch_bounds__(kWarps * 32)
loglik_kernel(LoglikLaunch p) {
  extern __shared__ __align__(16) double dsh[];   // kWarps * maxn_gauss residuals
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const int b = blockIdx.x * kWarps + wib;
  if (b >= p.B) return;
  const int T = p.ts.ntargets;
  int maxn = 0;
  for (int t = 0; t < T; ++t)
    if (p.ts.t[t].cov == BH_COV_GAUSS && p.ts.t[t].n > maxn) maxn = p.ts.t[t].n;
  double logL = 0.0, joint = 0.0;
  bool valid = true;
  for (int t = 0; t < T; ++t)
    if (p.tstatus[(size_t)b * kMaxTargets + t] == 0) valid = false;

  for (int t = 0; t < T; ++t) {
    const TargetDev& Tg = p.ts.t[t];
    const int n = Tg.n;
    const bool tvalid = p.tstatus[(size_t)b * kMaxTargets + t] != 0;
    const double* __restrict__ curve = p.curves + (size_t)b * p.curve_stride + p.curve_off[t];
    const double* __restrict__ trace = p.rfsynth + (size_t)b * p.ts.synth_stride + Tg.synth_off;
    double* __restrict__ sy = p.synth ? p.synth + (size_t)b * p.ts.synth_stride + Tg.synth_off : nullptr;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, se = 0.0;
    double* dres = dsh + (size_t)wib * maxn;
    if (tvalid) {
      for (int i = lane; i < n; i += 32) {
        double ym = modelled(Tg, curve, trace, i);
        double d = ym - Tg.y[i];
        if (sy) sy[i] = ym;
        double dd = d * d;
        s0 += dd;
        if (Tg.cov == BH_COV_EXP) {
          if (i > 0 && i < n - 1) s1 += dd;
          if (i < n - 1) {
            double dn = modelled(Tg, curve, trace, i + 1) - Tg.y[i + 1];
            s2 += d * dn;
          }
        } else if (Tg.cov == BH_COV_WHITE_SCALED) {
          se += dd / Tg.serr[i];
        } else if (Tg.cov == BH_COV_GAUSS) {
          dres[i] = d;
        }
      }
    } else if (sy) {
      for (int i = lane; i < n; i += 32) sy[i] = NAN;
    }
    if (!valid) continue;   // sentinels; still fill synth of the remaining targets
    if (Tg.cov == BH_COV_GAUSS) {
      __syncwarp();
      const double* __restrict__ R = Tg.corr_inv;
      double acc = 0.0;
      for (int i = 0; i < n; ++i) {
        const double di = dres[i];
        const double* __restrict__ Ri = R + (size_t)i * n;
        double rowacc = 0.0;
        for (int j = lane; j < n; j += 32) rowacc += Ri[j] * dres[j];
        acc += di * rowacc;
      }
      se = acc;
      __syncwarp();
    }
    s0 = warp_sum(s0); s1 = warp_sum(s1); s2 = warp_sum(s2); se = warp_sum(se);
    const double corr = p.noise[(size_t)b * 2 * T + 2 * t];
    const double sigma = p.noise[(size_t)b * 2 * T + 2 * t + 1];
    const double sig2 = sigma * sigma;
    double phi, logdet = 2.0 * n * log(sigma);
    if (Tg.cov == BH_COV_EXP) {
      const double om = 1.0 - corr * corr;
      phi = (s0 + corr * corr * s1 - 2.0 * corr * s2) / (sig2 * om);
      logdet += (n - 1) * log(om);
    } else if (Tg.cov == BH_COV_WHITE) {
      phi = s0 / sig2;
    } else if (Tg.cov == BH_COV_WHITE_SCALED) {
      phi = se / sig2;
      logdet += Tg.log_serr_prod;
    } else {
      phi = se / sig2;
      logdet += Tg.logcorr_det;
    }
    logL += -0.5 * (n * kLog2Pi + logdet) - phi / 2.0;
    const double rms = sqrt(s0 / n);
    joint += rms;
    if (lane == 0) p.misfits[(size_t)b * (T + 1) + t] = rms;
  }
  if (lane == 0) {
    if (valid) {
      p.logL[b] = logL;
      p.misfits[(size_t)b * (T + 1) + T] = joint;
      p.status[b] = 1;
    } else {
      p.logL[b] = -1e15;
      for (int t = 0; t <= T; ++t) p.misfits[(size_t)b * (T + 1) + t] = 1e15;
      p.status[b] = 0;
    }
  }
}

}  // namespace

void launch_loglik(const LoglikLaunch& p, cudaStream_t st) {
  if (p.B <= 0) return;
  int maxn = 0;
  for (int t = 0; t < p.ts.ntargets; ++t)
    if (p.ts.t[t].cov == BH_COV_GAUSS && p.ts.t[t].n > maxn) maxn = p.ts.t[t].n;
  const size_t smem = sizeof(double) * (size_t)maxn * kWarps;
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    cudaFuncSetAttribute(loglik_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    configured = smem;
  }
  const int blocks = (p.B + kWarps - 1) / kWarps;
  static bool carved = false;
  if (!carved) { bh_set_carveout(loglik_kernel); carved = true; }
  loglik_kernel<<<blocks, kWarps * 32, smem, st>>>(p);
}

}  // namespace bh
