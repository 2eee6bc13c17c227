// prep_kernel.cu -- model unpacking for both forward kernels.
//
// One thread per (model, layer row).  From the packed fp64 rows
// (vs, vp/vs, z_top, h) it derives what the reference derives on the host:
//   vp  = vs * vpvs                          (Models.py:49-52)
//   rho = 0.32*vp + 0.77 unless given        (Targets.py:319)
// and writes
//   * the REAL*4 model row SURF96 sees after f2py's cast   (surfdisp96.f:82)
//   * the earth-flattened RF layer constants                (model.cpp:223-251)
//   * the R/T matrices of the interface ABOVE this row      (greens.cpp:462-468)
//   * per model (row 0): 2*H and the P/SV decomposition constants.
#include "kernels.h"

namespace bh {

struct RawLayer { double z, h, vp, vs, rho, qp, qs; };

__device__ __forceinline__ RawLayer load_packed(const double* __restrict__ model,
                                                const double* __restrict__ rho, int b, int lmax,
                                                int li, int nl, double qp, double qs) {
  const double4 r = *reinterpret_cast<const double4*>(model + ((size_t)b * lmax + li) * 4);
  RawLayer L;
  L.vs = r.x;
  L.vp = dmul(r.x, r.y);                               // vs * vpvs, one rounding
  L.z = r.z;
  L.rho = rho ? rho[(size_t)b * lmax + li] : dadd(dmul(L.vp, 0.32), 0.77);
  // RF thickness is the z-difference (synrf.cpp:29), half-space -1 (:31)
  if (li < nl - 1) L.h = model[((size_t)b * lmax + li + 1) * 4 + 2] - r.z;
  else L.h = -1.0;
  L.qp = qp; L.qs = qs;
  return L;
}

__device__ __forceinline__ void write_rf(const RawLayer& up, const RawLayer& lo, bool has_up, int li,
                                         double u, double nsv, double sigma, RfLayer* lay_out,
                                         cm2* coef_out, double* mc_out) {
  double h, vp, vs, rho;
  rf_flatten(lo.z, lo.h, lo.vp, lo.vs, lo.rho, &h, &vp, &vs, &rho);
  RfLayer o;
  o.h = h; o.vp = vp; o.vs = vs; o.rho = rho;
  o.cqp = 1.0 / (RF_PI * lo.qp); o.bqp = 1.0 / (2.0 * lo.qp);
  o.cqs = 1.0 / (RF_PI * lo.qs); o.bqs = 1.0 / (2.0 * lo.qs);
  *lay_out = o;
  cm2 rd, td, ru, tu;
  const cd zero = mk(0.0, 0.0);
  rd.a11 = rd.a12 = rd.a21 = rd.a22 = zero;
  td = rd; tu = rd; ru = rd;
  if (!has_up) {
    rf_coeff_surface(u, vp, vs, &ru);
    cm2 h2;
    rf_displacement2(u, vp, vs, &h2);
    mc_out[0] = h2.a11.re; mc_out[1] = h2.a11.im; mc_out[2] = h2.a12.re; mc_out[3] = h2.a12.im;
    mc_out[4] = h2.a21.re; mc_out[5] = h2.a21.im; mc_out[6] = h2.a22.re; mc_out[7] = h2.a22.im;
    double m[4];
    bool on = rf_decomp_consts(u, nsv, sigma, m);
    mc_out[8] = m[0]; mc_out[9] = m[1]; mc_out[10] = m[2]; mc_out[11] = m[3];
    mc_out[12] = on ? 1.0 : 0.0;
    mc_out[13] = mc_out[14] = mc_out[15] = 0.0;
  } else {
    double h1, vp1, vs1, rho1;
    rf_flatten(up.z, up.h, up.vp, up.vs, up.rho, &h1, &vp1, &vs1, &rho1);
    rf_coeff_interface(u, vp1, vs1, rho1, vp, vs, rho, &rd, &td, &ru, &tu);
  }
  coef_out[0] = rd; coef_out[1] = td; coef_out[2] = ru; coef_out[3] = tu;
  (void)li;
}

#ifndef BH_PREP_THREADS
#define BH_PREP_THREADS 128
#endif
__global__ void __launch_bounds__(BH_PREP_THREADS)
prepare_kernel(const double* __restrict__ model, const int* __restrict__ nlay,
               const double* __restrict__ rho, int B, int lmax, int want_swd, int want_rf,
               double rf_p, double rf_nsv, double rf_qp, double rf_qs, PrepOut out) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * lmax) return;
  int b = idx / lmax, li = idx - b * lmax;
  int nl = nlay[b];
  if (nl > lmax) nl = lmax;
  if (li >= nl) return;
  RawLayer lo = load_packed(model, rho, b, lmax, li, nl, rf_qp, rf_qs);
  if (want_swd) {
    const double hpk = model[((size_t)b * lmax + li) * 4 + 3];
    LayerRow r;
    r.x = (float)hpk; r.y = (float)lo.vp; r.z = (float)lo.vs; r.w = (float)lo.rho;
    out.swd_rows[(size_t)b * out.swd_stride + li] = r;
  }
  if (want_rf) {
    RawLayer up = lo;
    double nsv = rf_nsv, sigma = 0.0;
    if (li > 0) up = load_packed(model, rho, b, lmax, li - 1, nl, rf_qp, rf_qs);
    else {
      // rfmini_modrf.py:125-130: Poisson ratio from the top-layer vp/vs, nsv = vs[0]
      double k = lo.vp / lo.vs;
      sigma = (2.0 - k * k) / (2.0 - 2.0 * k * k);
      if (!(nsv > 0.0)) nsv = lo.vs;
    }
    write_rf(up, lo, li > 0, li, rf_p * RF_DEG_PER_KM, nsv, sigma,
             out.rf_lay + (size_t)b * lmax + li, out.rf_coef + ((size_t)b * lmax + li) * 4,
             out.rf_mc + (size_t)b * 16);
  }
}

// single model given as explicit arrays (bh_synrf shim): thread per layer row
__global__ void prepare_rf_explicit_kernel(const double* __restrict__ z, const double* __restrict__ vp,
                                           const double* __restrict__ vs, const double* __restrict__ rho,
                                           const double* __restrict__ qp, const double* __restrict__ qs,
                                           int nl, double p, double nsv, double sigma, PrepOut out) {
  int li = blockIdx.x * blockDim.x + threadIdx.x;
  if (li >= nl) return;
  RawLayer lo, up;
  lo.z = z[li]; lo.vp = vp[li]; lo.vs = vs[li]; lo.rho = rho[li]; lo.qp = qp[li]; lo.qs = qs[li];
  lo.h = (li < nl - 1) ? z[li + 1] - z[li] : -1.0;
  up = lo;
  if (li > 0) {
    up.z = z[li - 1]; up.vp = vp[li - 1]; up.vs = vs[li - 1]; up.rho = rho[li - 1];
    up.qp = qp[li - 1]; up.qs = qs[li - 1];
    up.h = z[li] - z[li - 1];
  }
  write_rf(up, lo, li > 0, li, p * RF_DEG_PER_KM, nsv, sigma, out.rf_lay + li,
           out.rf_coef + (size_t)li * 4, out.rf_mc);
}

// Counting sort of the models by layer count, longest first: one CTA; the bins live in shared memory.
__global__ void __launch_bounds__(1024)
layer_order_kernel(const int* __restrict__ nlay, int B, int* __restrict__ perm, int* __restrict__ maxn) {
  __shared__ int hist[128], start[128];
  const int t = threadIdx.x;
  if (t < 128) hist[t] = 0;
  __syncthreads();
  for (int b = t; b < B; b += blockDim.x) {
    int n = nlay[b]; n = n < 0 ? 0 : (n > 127 ? 127 : n);
    atomicAdd(&hist[n], 1);
  }
  __syncthreads();
  if (t == 0) {
    int acc = 0, mx = 0;
    for (int n = 127; n >= 0; --n) { start[n] = acc; acc += hist[n]; if (hist[n] && n > mx) mx = n; }
    if (maxn) *maxn = mx;
  }
  __syncthreads();
  for (int b = t; b < B; b += blockDim.x) {
    int n = nlay[b]; n = n < 0 ? 0 : (n > 127 ? 127 : n);
    perm[atomicAdd(&start[n], 1)] = b;
  }
}

void launch_layer_order(const int* nlay, int B, int* perm, int* maxn, cudaStream_t st) {
  if (B <= 0) return;
  layer_order_kernel<<<1, 1024, 0, st>>>(nlay, B, perm, maxn);
}

void launch_prepare(const double* model, const int* nlay, const double* rho, int B, int lmax,
                    bool want_swd, bool want_rf, double rf_p, double rf_nsv, double rf_qp,
                    double rf_qs, PrepOut out, cudaStream_t st) {
  int total = B * lmax;
  if (total <= 0) return;
  int threads = BH_PREP_THREADS;
  int blocks = (total + threads - 1) / threads;
  static KernelAttrs attrs;
  bh_configure_kernel(prepare_kernel, 0, attrs);
  prepare_kernel<<<blocks, threads, 0, st>>>(model, nlay, rho, B, lmax, want_swd ? 1 : 0,
                                             want_rf ? 1 : 0, rf_p, rf_nsv, rf_qp, rf_qs, out);
}

void launch_prepare_rf_explicit(const double* z, const double* vp, const double* vs,
                                const double* rho, const double* qp, const double* qs, int nlay,
                                double p, double nsv, double sigma, PrepOut out, cudaStream_t st) {
  int threads = 128;
  int blocks = (nlay + threads - 1) / threads;
  static KernelAttrs attrs;
  bh_configure_kernel(prepare_rf_explicit_kernel, 0, attrs);
  prepare_rf_explicit_kernel<<<blocks, threads, 0, st>>>(z, vp, vs, rho, qp, qs, nlay, p, nsv,
                                                         sigma, out);
}

}  // namespace bh
