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

// modelled datum i of a target: a searched dispersion curve (resampled back when the target has more than 60
// periods), a receiver-function trace, or -- direct -- data the caller supplied (bh_engine_loglik_host)
__device__ __forceinline__ double modelled(const TargetDev& T, const double* __restrict__ curve,
                                           const double* __restrict__ trace, int i, bool direct = false) {
  if (!direct && T.ref <= BH_REF_LDISPGR) {
    if (T.n == T.kmax) return curve[i];
    return interp_np(T.x[i], T.periods, curve, T.kmax);
  }
  return trace[i];
}


// ---------------------------------------------------------------------------
// Gauss law (Targets.py:162-173, :339-340): Phi_b = d_b^T R^-1 d_b for every model of the batch --
// the one dense contraction of the path (2 n^2 flop per model; n = 512: 4.3 Gflop per 8192 models).
// Written as a batched product instead of one matrix sweep per model: a CTA takes 64 models and walks the
// upper triangle of 32 x 32 tiles of S = (R^-1 + R^-T) / 2 (the antisymmetric part of a matrix does not
// contribute to a quadratic form, so Phi = d^T S d exactly; off-diagonal tiles count twice).  Per tile pair
//   Y[64 x 32] = D[:, kt] (64 x 32) * S[kt, jt] (32 x 32)      fp64 tensor cores, mma.m8n8k4.f64
//   Phi_b     += w * sum_j Y[b, j] * D[b, jt + j]
// with the residual tiles D formed on the fly (modelled - observed) in shared memory.  Every S tile is read
// once per 64 models instead of once per model.
// ---------------------------------------------------------------------------
constexpr int QF_BM = 64;          // models per CTA
constexpr int QF_T = 32;           // tile edge
constexpr int QF_LD = 36;          // padded leading dimension: conflict-free fragment reads (36 = 4 mod 16)

__device__ __forceinline__ void dmma_8x8x4(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// residuals d = modelled - observed of a Gauss-law target, [B][ldr] with ldr = n rounded up to the tile edge
// (zero padded; zero rows for invalid models)
__global__ void __launch_bounds__(256)
gauss_residual_kernel(LoglikLaunch p, int t, int ldr) {
  const TargetDev& Tg = p.ts.t[t];
  const int b = blockIdx.x;
  const bool ok = p.tstatus[(size_t)b * kMaxTargets + t] != 0;
  const double* curve = p.curves + (size_t)b * p.curve_stride + p.curve_off[t];
  const bool direct = p.given != nullptr;
  const double* trace = (direct ? p.given : p.rfsynth) + (size_t)b * p.ts.synth_stride + Tg.synth_off;
  double* out = p.gauss_res + (size_t)b * ldr;
  for (int i = threadIdx.x; i < ldr; i += blockDim.x)
    out[i] = (ok && i < Tg.n) ? modelled(Tg, curve, trace, i, direct) - Tg.y[i] : 0.0;
}

// grid (model blocks of 64, tile rows kt): the CTA walks the tiles (kt, jt >= kt) of its row and leaves its
// partial sum in part[b][kt]; loglik_kernel adds the partials in a fixed order (no atomics: the result does
// not depend on the order in which CTAs finish).
__global__ void __launch_bounds__(128)
gauss_quadform_kernel(LoglikLaunch p, int t, int ldr, int nt) {
  __shared__ __align__(16) double Dk[QF_BM * QF_LD];     // residuals, columns of tile kt
  __shared__ __align__(16) double Dj[QF_BM * QF_LD];     // residuals, columns of tile jt
  __shared__ __align__(16) double Ss[QF_T * QF_LD];      // S[kt, jt]
  const TargetDev& Tg = p.ts.t[t];
  const int n = Tg.n;
  const int b0 = blockIdx.x * QF_BM;
  const int kt = blockIdx.y;
  const int tidx = threadIdx.x, lane = tidx & 31, warp = tidx >> 5;
  const double* __restrict__ S = Tg.corr_inv;
  const double* __restrict__ R = p.gauss_res;
  for (int e = tidx; e < QF_BM * QF_T; e += 128) {
    const int r = e / QF_T, c = e % QF_T;
    Dk[r * QF_LD + c] = (b0 + r < p.B) ? R[(size_t)(b0 + r) * ldr + kt * QF_T + c] : 0.0;
  }
  double phi[2] = {0.0, 0.0};                            // rows 16 * warp + 8 * mt + lane / 4
  for (int jt = kt; jt < nt; ++jt) {
    __syncthreads();
    for (int e = tidx; e < QF_BM * QF_T; e += 128) {
      const int r = e / QF_T, c = e % QF_T;
      Dj[r * QF_LD + c] = (b0 + r < p.B) ? R[(size_t)(b0 + r) * ldr + jt * QF_T + c] : 0.0;
    }
    for (int e = tidx; e < QF_T * QF_T; e += 128) {
      const int r = e / QF_T, c = e % QF_T;
      const int gi = kt * QF_T + r, gj = jt * QF_T + c;
      Ss[r * QF_LD + c] = (gi < n && gj < n) ? S[(size_t)gi * n + gj] : 0.0;
    }
    __syncthreads();
    double acc[2][4][2];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[mt][q][0] = acc[mt][q][1] = 0.0;
#pragma unroll
    for (int k0 = 0; k0 < QF_T; k0 += 4) {
      double a[2], bb[4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) a[mt] = Dk[(16 * warp + 8 * mt + (lane >> 2)) * QF_LD + k0 + (lane & 3)];
#pragma unroll
      for (int q = 0; q < 4; ++q) bb[q] = Ss[(k0 + (lane & 3)) * QF_LD + 8 * q + (lane >> 2)];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int q = 0; q < 4; ++q) dmma_8x8x4(acc[mt][q][0], acc[mt][q][1], a[mt], bb[q]);
    }
    const double w = (jt == kt) ? 1.0 : 2.0;
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      const double* dj = Dj + (16 * warp + 8 * mt + (lane >> 2)) * QF_LD + 2 * (lane & 3);
      double sacc = 0.0;
#pragma unroll
      for (int q = 0; q < 4; ++q) sacc += acc[mt][q][0] * dj[8 * q] + acc[mt][q][1] * dj[8 * q + 1];
      phi[mt] += w * sacc;
    }
  }
#pragma unroll
  for (int mt = 0; mt < 2; ++mt) {
    double v = phi[mt];
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    const int b = b0 + 16 * warp + 8 * mt + (lane >> 2);
    if ((lane & 3) == 0 && b < p.B) p.gauss_part[(size_t)b * nt + kt] = v;
  }
}

__global__ void __launch_bounds__(kWarps * 32)
loglik_kernel(LoglikLaunch p) {
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const int b = blockIdx.x * kWarps + wib;
  if (b >= p.B) return;
  const int T = p.ts.ntargets;
  double logL = 0.0, joint = 0.0;
  bool valid = true;
  for (int t = 0; t < T; ++t)
    if (p.tstatus[(size_t)b * kMaxTargets + t] == 0) valid = false;

  for (int t = 0; t < T; ++t) {
    const TargetDev& Tg = p.ts.t[t];
    const int n = Tg.n;
    const bool tvalid = p.tstatus[(size_t)b * kMaxTargets + t] != 0;
    const double* __restrict__ curve = p.curves + (size_t)b * p.curve_stride + p.curve_off[t];
    const bool direct = p.given != nullptr;
    const double* __restrict__ trace = (direct ? p.given : p.rfsynth) + (size_t)b * p.ts.synth_stride + Tg.synth_off;
    double* __restrict__ sy = p.synth ? p.synth + (size_t)b * p.ts.synth_stride + Tg.synth_off : nullptr;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, se = 0.0;
    if (tvalid) {
      for (int i = lane; i < n; i += 32) {
        double ym = modelled(Tg, curve, trace, i, direct);
        double d = ym - Tg.y[i];
        if (sy) sy[i] = ym;
        double dd = d * d;
        s0 += dd;
        if (Tg.cov == BH_COV_EXP) {
          if (i > 0 && i < n - 1) s1 += dd;
          if (i < n - 1) {
            double dn = modelled(Tg, curve, trace, i + 1, direct) - Tg.y[i + 1];
            s2 += d * dn;
          }
        } else if (Tg.cov == BH_COV_WHITE_SCALED) {
          se += dd / Tg.serr[i];
        }
      }
    } else if (sy) {
      for (int i = lane; i < n; i += 32) sy[i] = NAN;
    }
    if (!valid) continue;   // sentinels; still fill synth of the remaining targets
    s0 = warp_sum(s0); s1 = warp_sum(s1); s2 = warp_sum(s2); se = warp_sum(se);
    if (Tg.cov == BH_COV_GAUSS) se = p.gauss_phi[(size_t)b * kMaxTargets + t];    // d^T R^-1 d (gauss_quadform_kernel)
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

// sum of a model's tile-row partials in a fixed order
__global__ void gauss_reduce_kernel(LoglikLaunch p, int t, int nt) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= p.B) return;
  double v = 0.0;
  for (int k = 0; k < nt; ++k) v += p.gauss_part[(size_t)b * nt + k];
  p.gauss_phi[(size_t)b * kMaxTargets + t] = v;
}

}  // namespace

int gauss_tile_rows(int n) { return (n + QF_T - 1) / QF_T; }

// d^T R^-1 d of Gauss-law target t for the whole batch -> p.gauss_phi[b][t].  Needs the target's modelled data
// (curves / rfsynth) and tstatus; uses the scratch p.gauss_res / p.gauss_part (one contraction at a time).
void launch_gauss_quadform(const LoglikLaunch& p, int t, cudaStream_t st) {
  if (p.B <= 0 || p.ts.t[t].cov != BH_COV_GAUSS) return;
  static KernelAttrs attrs;
  bh_configure_kernel(gauss_quadform_kernel, 0, attrs);
  const int nt = gauss_tile_rows(p.ts.t[t].n), ldr = nt * QF_T;
  gauss_residual_kernel<<<p.B, 256, 0, st>>>(p, t, ldr);
  gauss_quadform_kernel<<<dim3((p.B + QF_BM - 1) / QF_BM, nt), 128, 0, st>>>(p, t, ldr, nt);
  gauss_reduce_kernel<<<(p.B + 255) / 256, 256, 0, st>>>(p, t, nt);
}

void launch_loglik(const LoglikLaunch& p, cudaStream_t st) {
  if (p.B <= 0) return;
  static KernelAttrs attrs;
  bh_configure_kernel(loglik_kernel, 0, attrs);
  const int blocks = (p.B + kWarps - 1) / kWarps;
  loglik_kernel<<<blocks, kWarps * 32, 0, st>>>(p);
}

}  // namespace bh
