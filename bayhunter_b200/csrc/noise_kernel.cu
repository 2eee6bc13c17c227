// noise_kernel.cu -- correlated-noise realisations for synthetic observations, on the device.
//
// Behavioural reference: BayHunter src/SynthObs.py:136-155
//   compute_expnoise    N(0, sigma^2 R), R_ij = corr^|i-j|        (exponential law)
//   compute_gaussnoise  N(0, sigma^2 R), R_ij = corr^((i-j)^2)    (Gaussian law, RF with a Gauss filter)
// The reference draws ONE realisation with numpy's multivariate_normal (an SVD of the dense covariance per
// call).  Here B realisations come out of one launch, from the sampler's counter-based generator
// (Philox4x32-10 keyed by (seed, realisation, sample): reproducible, independent of the batch split):
//   exponential law: the covariance is that of a stationary AR(1) process -- e_0 = sigma z_0,
//     e_i = corr e_{i-1} + sigma sqrt(1 - corr^2) z_i -- exact, O(n) per realisation, no matrix;
//   Gaussian law: e = sigma z F with a factor F (F^T F = R) computed once on the host the way numpy does
//     (SVD: F = sqrt(s) v), because R is numerically singular for the correlations in use (r >= 0.9).
#include <cuda_runtime.h>

#include "../../include/bayhunter_b200.h"
#include "kernels.h"
#include "sampler_core.cuh"

namespace bh {
int bh_set_error_message(int code, const char* what);

namespace {

// two standard normals from one Philox block (Box-Muller), counter = (realisation, pair index)
__device__ __forceinline__ void normal_pair(unsigned long long seed, unsigned long long real, unsigned pair,
                                            double* z0, double* z1) {
  uint32_t c[4] = {(uint32_t)real, (uint32_t)(real >> 32), pair, 0x6e6f6973u /* "nois" */};
  philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
  const double u1 = 1.0 - u01_53(c[0], c[1]);            // (0, 1]
  const double u2 = u01_53(c[2], c[3]);
  const double r = sqrt(-2.0 * log(u1));
  double s, cs;
  sincospi(2.0 * u2, &s, &cs);
  *z0 = r * cs;
  *z1 = r * s;
}

__global__ void expnoise_kernel(int n, int B, double corr, double sigma, unsigned long long seed, double* out) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const double inn = sigma * sqrt(1.0 - corr * corr);
  double e = 0.0;
  for (int i = 0; i < n; i += 2) {
    double z0, z1;
    normal_pair(seed, (unsigned long long)b, (unsigned)(i >> 1), &z0, &z1);
    e = (i == 0) ? sigma * z0 : fma(corr, e, inn * z0);
    out[(size_t)b * n + i] = e;
    if (i + 1 < n) {
      e = fma(corr, e, inn * z1);
      out[(size_t)b * n + i + 1] = e;
    }
  }
}

// one CTA per realisation: white variates into shared memory, then e_i = sigma * sum_k z_k F[k][i]
__global__ void factornoise_kernel(int n, double sigma, unsigned long long seed, const double* __restrict__ F,
                                   double* out) {
  extern __shared__ double z[];
  const int b = blockIdx.x;
  for (int p = threadIdx.x; 2 * p < n; p += blockDim.x) {
    double z0, z1;
    normal_pair(seed, (unsigned long long)b, (unsigned)p, &z0, &z1);
    z[2 * p] = z0;
    if (2 * p + 1 < n) z[2 * p + 1] = z1;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    double acc = 0.0;
    for (int k = 0; k < n; ++k) acc = fma(z[k], F[(size_t)k * n + i], acc);
    out[(size_t)b * n + i] = sigma * acc;
  }
}

}  // namespace
}  // namespace bh

using namespace bh;

extern "C" int bh_correlated_noise(int law, int n, int B, double corr, double sigma, unsigned long long seed,
                                   const double* factor, double* out) {
  if (!out || n < 1 || B < 1) return bh_set_error_message(BH_ERR_ARG, "n, B >= 1 and out required");
  if (law != BH_COV_EXP && law != BH_COV_GAUSS) return bh_set_error_message(BH_ERR_ARG, "law must be BH_COV_EXP or BH_COV_GAUSS");
  if (law == BH_COV_EXP && !(fabs(corr) < 1.0)) return bh_set_error_message(BH_ERR_ARG, "|corr| < 1 required");
  if (law == BH_COV_GAUSS && !factor) return bh_set_error_message(BH_ERR_ARG, "the Gaussian law needs the host-computed factor");
  if (bh_device_count() < 1) return bh_set_error_message(BH_ERR_NO_DEVICE, "no CUDA device visible; this library has no CPU path");
  double *d_out = nullptr, *d_f = nullptr;
  cudaError_t ce = cudaMalloc((void**)&d_out, sizeof(double) * (size_t)n * B);
  if (ce == cudaSuccess && law == BH_COV_GAUSS) {
    ce = cudaMalloc((void**)&d_f, sizeof(double) * (size_t)n * n);
    if (ce == cudaSuccess) ce = cudaMemcpy(d_f, factor, sizeof(double) * (size_t)n * n, cudaMemcpyHostToDevice);
  }
  if (ce == cudaSuccess) {
    if (law == BH_COV_EXP) expnoise_kernel<<<(B + 127) / 128, 128>>>(n, B, corr, sigma, seed, d_out);
    else factornoise_kernel<<<B, 128, sizeof(double) * (size_t)(n + 1)>>>(n, sigma, seed, d_f, d_out);
    ce = cudaMemcpy(out, d_out, sizeof(double) * (size_t)n * B, cudaMemcpyDeviceToHost);
  }
  cudaFree(d_out);
  cudaFree(d_f);
  if (ce != cudaSuccess) {
    bh_set_error_message(BH_ERR_CUDA, cudaGetErrorString(ce));
    return BH_ERR_CUDA;
  }
  return BH_OK;
}
