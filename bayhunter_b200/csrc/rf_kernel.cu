// rf_kernel.cu -- receiver-function synthesis on sm_100a.
//
// Two kernels:
//   rf_spectrum_kernel  one thread per (model, frequency) item over a flat index
//                       space (perfect balance, ragged layer counts only cost
//                       intra-warp divergence at model boundaries).  Runs the
//                       reflectivity recursion and the spectral division and
//                       writes the filtered spectrum crf[b][0..N/2].
//   rf_synth_kernel     one CTA per model: Hermitian extension, radix-2
//                       inverse FFT in shared memory, first ndata samples out.
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

// radix-2 decimation-in-time inverse transform, sign +1, total scale 1/N
// (fork.cpp:10-60 with signi = +1 and iftr's second 1/sqrt(N), greens.cpp:147,157)
__global__ void rf_synth_kernel(RfLaunch p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cd* x = reinterpret_cast<cd*>(smem_raw);
  const int N = p.k.nsamp;
  const int nfreq = N / 2 + 1;
  const int b = blockIdx.x;
  const cd* __restrict__ spec = p.spec + (size_t)b * nfreq;
  int logn = 0;
  while ((1 << logn) < N) ++logn;
  // Hermitian extension + bit reversal (iftr, greens.cpp:149-152)
  const int nact = p.nact;
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    const int k = (i <= N / 2) ? i : N - i;
    cd v = mk(0.0, 0.0);
    if (k < nact) v = (i <= N / 2) ? spec[k] : cconj(spec[k]);
    int r = (int)(__brev((unsigned)i) >> (32 - logn));
    x[r] = v;
  }
  // twiddles exp(+i pi k / (N/2)), k < N/2, once per CTA (the reference recomputes exp() inside
  // its butterfly loops, fork.cpp:40-55; same values)
  cd* tw = x + N;
  for (int k = threadIdx.x; k < N / 2; k += blockDim.x) {
    double sw, cw;
    sincospi((double)k / (double)(N / 2), &sw, &cw);
    tw[k] = mk(cw, sw);
  }
  __syncthreads();
  for (int l = 1, sh = logn - 1; l < N; l <<= 1, --sh) {
    for (int t = threadIdx.x; t < N / 2; t += blockDim.x) {
      int m = t & (l - 1);
      int i = ((t - m) << 1) + m;
      cd wv = tw[m << sh];                      // exp(i pi m / l) = tw[m * (N/2) / l]
      cd a = x[i], bb = wv * x[i + l];
      x[i] = a + bb;
      x[i + l] = a - bb;
    }
    __syncthreads();
  }
  const double scale = 1.0 / (double)N;
  double* __restrict__ out = p.out + (size_t)b * p.out_stride + p.out_off;
  for (int i = threadIdx.x; i < p.ndata; i += blockDim.x) out[i] = x[i].re * scale;
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
  static bool carved = false;
  if (!carved) { bh_set_carveout(rf_spectrum_kernel); carved = true; }
  rf_spectrum_kernel<<<(int)blocks, threads, 0, st>>>(p);
}

void launch_rf_synth(const RfLaunch& p, cudaStream_t st) {
  if (p.B <= 0) return;
  const int N = p.k.nsamp;
  int threads = N / 2;
  if (threads > 512) threads = 512;
  if (threads < 32) threads = 32;
  const size_t smem = sizeof(cd) * ((size_t)N + (size_t)N / 2);     // samples + twiddles
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    cudaFuncSetAttribute(rf_synth_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    configured = smem;
  }
  static bool carved = false;
  if (!carved) { bh_set_carveout(rf_synth_kernel); carved = true; }
  rf_synth_kernel<<<p.B, threads, smem, st>>>(p);
}

}  // namespace bh
