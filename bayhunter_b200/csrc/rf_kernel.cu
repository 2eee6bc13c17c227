// rf_kernel.cu -- receiver-function synthesis on sm_100a.
//
// Two kernels:
//   rf_spectrum_kernel  one thread per (model, frequency) item over a flat index
//                       space (perfect balance, ragged layer counts only cost
//                       intra-warp divergence at model boundaries).  Runs the
//                       reflectivity recursion and the spectral division and
//                       writes the filtered spectrum crf[b][0..N/2].
//   rf_synth_kernel     one CTA per model: real inverse FFT of the Hermitian spectrum as one
//                       half-size complex radix-2 transform in shared memory, first ndata samples out.
// The spectrum (B x (N/2+1) x 16 B) stays L2-resident between the two.
#include "kernels.h"

// threads per CTA of the spectrum kernel
#ifndef BH_RF_THREADS
#define BH_RF_THREADS 128
#endif

namespace bh {

namespace {

__global__ void __launch_bounds__(BH_RF_THREADS)
rf_spectrum_kernel(RfLaunch p) {
  const int nfreq = p.k.nsamp / 2 + 1;
  const int nact = p.nact;                       // bins that matter (rf_active_frequencies)
  const long long total = (long long)p.B * nact;
  const double u2 = p.k.u * p.k.u;
  for (long long it = (long long)blockIdx.x * blockDim.x + threadIdx.x; it < total;
       it += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(it / nact);
    const int j = (int)(it - (long long)b * nact);
    const long long idx = (long long)b * nfreq + j;
    int nl = p.nlay[b];
    if (nl > p.lmax) nl = p.lmax;
    cd out = mk(NAN, NAN);
    if (nl >= 2) {
      const double* __restrict__ mc = p.mc + (size_t)b * 16;
      cm2 h2;
      h2.a11 = mk(mc[0], mc[1]); h2.a12 = mk(mc[2], mc[3]);
      h2.a21 = mk(mc[4], mc[5]); h2.a22 = mk(mc[6], mc[7]);
      const double w = p.k.dw * j;
      const double lgw = j ? log(w / p.k.wref) : 0.0;
      cm2 t = rf_transfer(p.lay + (size_t)b * p.lmax, p.coef + (size_t)b * p.lmax * 4, h2, nl, u2,
                          w, lgw);
      double dm[4] = {mc[8], mc[9], mc[10], mc[11]};
      out = rf_spectral_value(t, p.k, dm, mc[12] != 0.0, j);
    }
    p.spec[idx] = out;
  }
}

// Inverse transform of the Hermitian spectrum (iftr, greens.cpp:136-158: Hermitian extension, radix-2
// transform with sign +1 and total scale 1/N, real part -- fork.cpp:10-60).  The N real samples come out
// of ONE complex transform of half the size: with M = N/2 and T[k] = exp(2 pi i k / N),
//     Z[k] = (X[k] + conj X[M-k]) + i T[k] (X[k] - conj X[M-k]),  k < M
//     z = sum_k Z[k] exp(2 pi i k m / M)  =>  x[2m] = Re z[m],  x[2m+1] = Im z[m]
// (taking the real part of the reference's full transform is the same as using Re X[0], Re X[M]).
// Half the butterflies, one stage fewer, and a quarter of the sincos calls of the full-size transform.
__global__ void rf_synth_kernel(RfLaunch p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cd* x = reinterpret_cast<cd*>(smem_raw);
  const int N = p.k.nsamp;
  const int M = N / 2;
  const int nfreq = M + 1;
  const int b = blockIdx.x;
  const cd* __restrict__ spec = p.spec + (size_t)b * nfreq;
  int logm = 0;
  while ((1 << logm) < M) ++logm;
  // twiddles T[k], k < M, once per CTA: a quarter turn is a swap, T[k + M/2] = i T[k]
  // (the reference recomputes exp() inside its butterfly loops, fork.cpp:40-55; same values)
  cd* tw = x + M;
  for (int k = threadIdx.x; k < (M + 1) / 2; k += blockDim.x) {
    double sw, cw;
    sincospi(2.0 * (double)k / (double)N, &sw, &cw);
    tw[k] = mk(cw, sw);
    if (M >= 2) tw[k + M / 2] = mk(-sw, cw);
  }
  __syncthreads();
  const int nact = p.nact;
  for (int k = threadIdx.x; k < M; k += blockDim.x) {
    cd a = mk(0.0, 0.0), c = mk(0.0, 0.0);
    if (k < nact) a = spec[k];
    if (M - k < nact) c = cconj(spec[M - k]);
    if (k == 0) { a.im = 0.0; c.im = 0.0; }
    const cd e = a + c, o = (a - c) * tw[k];
    const int r = logm ? (int)(__brev((unsigned)k) >> (32 - logm)) : 0;
    x[r] = mk(e.re - o.im, e.im + o.re);
  }
  __syncthreads();
  for (int l = 1, sh = logm - 1; l < M; l <<= 1, --sh) {
    for (int t = threadIdx.x; t < M / 2; t += blockDim.x) {
      int m = t & (l - 1);
      int i = ((t - m) << 1) + m;
      cd wv = tw[(m << sh) << 1];               // exp(i pi m / l) = T[m * M / l]
      cd a = x[i], bb = wv * x[i + l];
      x[i] = a + bb;
      x[i + l] = a - bb;
    }
    __syncthreads();
  }
  const double scale = 1.0 / (double)N;
  double* __restrict__ out = p.out + (size_t)b * p.out_stride + p.out_off;
  for (int i = threadIdx.x; i < p.ndata; i += blockDim.x) {
    const cd v = x[i >> 1];
    out[i] = ((i & 1) ? v.im : v.re) * scale;
  }
  if (threadIdx.x == 0 && p.tstatus) {
    int nl = p.nlay[b];
    p.tstatus[(size_t)b * kMaxTargets + p.target_id] = (nl >= 2) ? 1 : 0;
  }
}

}  // namespace

int rf_active_frequencies(const RfSpecConsts& k, double wfloor) {
  const int nfreq = k.nsamp / 2 + 1;
  if (!(wfloor > 0.0) || !(wfloor < 1.0) || !(k.a > 0.0) || !(k.dw > 0.0)) return nfreq;
  // exp(-0.25 (w/a)^2) >= floor  <=>  w <= 2 a sqrt(-ln floor); the reference clips w/a at 50
  // (greens.cpp:386), far above any floor that is not denormal
  const double wmax = 2.0 * k.a * sqrt(-log(wfloor));
  const double j = floor(wmax / k.dw) + 2.0;     // one bin of slack
  return j < (double)nfreq ? (int)j : nfreq;
}

void launch_rf_spectrum(const RfLaunch& p, cudaStream_t st) {
  const long long total = (long long)p.B * p.nact;
  if (total <= 0) return;
  const int threads = BH_RF_THREADS;
  long long blocks = (total + threads - 1) / threads;
  const long long cap = 148LL * 64 * (128 / BH_RF_THREADS);   // grid-stride beyond ~64 CTAs per SM
  if (blocks > cap) blocks = cap;
  static KernelAttrs attrs;
  bh_configure_kernel(rf_spectrum_kernel, 0, attrs);
  rf_spectrum_kernel<<<(int)blocks, threads, 0, st>>>(p);
}

void launch_rf_synth(const RfLaunch& p, cudaStream_t st) {
  if (p.B <= 0) return;
  const int N = p.k.nsamp;
  int threads = N / 4;                                              // one butterfly of the half-size transform each
  if (threads > 512) threads = 512;
  if (threads < 32) threads = 32;
  const size_t smem = sizeof(cd) * ((size_t)N / 2 + (size_t)N / 2 + 1);   // packed samples + twiddles
  static KernelAttrs attrs;
  bh_configure_kernel(rf_synth_kernel, smem, attrs);
  rf_synth_kernel<<<p.B, threads, smem, st>>>(p);
}

}  // namespace bh
