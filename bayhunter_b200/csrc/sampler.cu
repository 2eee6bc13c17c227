// sampler.cu -- lock-step ensemble of BayHunter chains on one GPU (C ABI: bh_sampler_*).
//
// One iteration of all B chains is three steps on one stream:
//   sampler_propose_kernel   one thread per chain: choose a modification, perturb, order,
//                            check the priors, pack the engine rows      (SingleChain.iterate :511-547)
//   bh_engine_eval           forward models + log-likelihood of the B proposals (Targets.evaluate)
//   sampler_accept_kernel    one thread per chain: Metropolis-Hastings test, bookkeeping, append the
//                            accepted model to the chain arrays, proposal-width control (:549-589)
// Chains never interact (src/mcmcOptimizer.py:208-216), so nothing here synchronises across chains
// and the host enqueues iterations back to back without reading anything.  Chains whose proposal
// fails its prior check sit the evaluation out (nlay = 0 -> the engine skips the model), exactly
// like the reference's early return (:541-547).
#include <cuda_runtime.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/bayhunter_b200.h"
#include "sampler_core.cuh"

using namespace bh;

namespace bh { int bh_set_error_message(int code, const char* what); }   // engine.cu: sets bh_last_error()

struct bh_sampler {
  bh_engine* eng = nullptr;
  SamplerCfg cfg{};
  int B = 0, S = 0, T = 0, maxl = 0;
  long long first_chain = 0;
  // current state
  double *cur_model = nullptr, *cur_vpvs = nullptr, *cur_noise = nullptr, *cur_logL = nullptr,
         *cur_misfits = nullptr, *propdist = nullptr;
  long long *accepted = nullptr, *proposed = nullptr, *iiter = nullptr;
  int *cur_k = nullptr, *nstored = nullptr;
  // proposal
  double *prop_model = nullptr, *prop_vpvs = nullptr, *prop_noise = nullptr, *prop_dvs2 = nullptr,
         *rows = nullptr, *p_logL = nullptr, *p_misfits = nullptr, *forced = nullptr;
  int *prop_k = nullptr, *prop_valid = nullptr, *prop_modify = nullptr, *nlay = nullptr, *p_status = nullptr;
  int use_forced = 0;
  // chain arrays (float32, NaN padded like the reference's shared RawArrays, mcmcOptimizer.py:78-128)
  float *st_models = nullptr, *st_misfits = nullptr, *st_likes = nullptr, *st_noise = nullptr, *st_vpvs = nullptr;
  int* st_iter = nullptr;
  unsigned long long* overflow = nullptr;
  long long *ovf_count = nullptr, *ovf_iter = nullptr;   // per chain: accepted models that found the arrays full, iteration of the first
  std::vector<void*> owned;
  cudaStream_t st = nullptr;
};

namespace {

struct SamplerDev {      // by-value kernel argument
  SamplerCfg cfg;
  int B, S, T, maxl;
  long long first_chain;
  double *cur_model, *cur_vpvs, *cur_noise, *cur_logL, *cur_misfits, *propdist;
  long long *accepted, *proposed, *iiter;
  int *cur_k, *nstored;
  double *prop_model, *prop_vpvs, *prop_noise, *prop_dvs2, *rows, *p_logL, *p_misfits;
  const double* forced;
  int *prop_k, *prop_valid, *prop_modify, *nlay, *p_status;
  float *st_models, *st_misfits, *st_likes, *st_noise, *st_vpvs;
  int* st_iter;
  unsigned long long* overflow;
  long long *ovf_count, *ovf_iter;
};

__device__ __forceinline__ Draw chain_draw(const SamplerDev& p, int b, long long iiter) {
  if (p.forced) {
    Draw d;
    d.u_mod = p.forced[4 * b]; d.u_idx = p.forced[4 * b + 1];
    d.gauss = p.forced[4 * b + 2]; d.u_acc = p.forced[4 * b + 3];
    return d;
  }
  return sampler_draw(p.cfg.seed, (unsigned long long)(p.first_chain + b), iiter);
}

__global__ void __launch_bounds__(64)
sampler_propose_kernel(SamplerDev p) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= p.B) return;
  const int maxl = p.maxl, T2 = 2 * p.T;
  double vs[SMP_MAX_ROWS], z[SMP_MAX_ROWS], h[SMP_MAX_ROWS], noise[2 * SMP_MAX_TARGETS];
  int k = p.cur_k[b];
  const double* cm = p.cur_model + (size_t)b * 2 * maxl;
  for (int i = 0; i < k; ++i) { vs[i] = cm[i]; z[i] = cm[maxl + i]; }
  for (int i = 0; i < T2; ++i) noise[i] = p.cur_noise[(size_t)b * T2 + i];
  double vpvs = p.cur_vpvs[b];
  const long long iiter = p.iiter[b];
  const Draw d = chain_draw(p, b, iiter);
  int modify = 0;
  double dvs2 = 0.0;
  int valid = sampler_propose(p.cfg, iiter, p.propdist + (size_t)b * SMP_NPAR, d, vs, z, &k, &vpvs, noise,
                              &modify, &dvs2, h);
  if (k > maxl) valid = 0;         // cannot happen for a valid model (layers <= layers_max)
  p.prop_valid[b] = valid;
  p.prop_modify[b] = modify;
  p.prop_dvs2[b] = dvs2;
  p.prop_k[b] = k;
  p.prop_vpvs[b] = vpvs;
  for (int i = 0; i < T2; ++i) p.prop_noise[(size_t)b * T2 + i] = noise[i];
  p.nlay[b] = valid ? k : 0;
  if (valid) {
    double* pm = p.prop_model + (size_t)b * 2 * maxl;
    for (int i = 0; i < k; ++i) { pm[i] = vs[i]; pm[maxl + i] = z[i]; }
    double* rows = p.rows + (size_t)b * maxl * 4;
    // rows are written straight from registers/local arrays: 32 B per layer
    double r[4];
    bool mantle = false;
    double ztop = 0.0;
    for (int i = 0; i < k; ++i) {
      if (p.cfg.has_mantle && vs[i] >= p.cfg.mantle_vs) mantle = true;
      r[0] = vs[i]; r[1] = mantle ? p.cfg.mantle_vpvs : vpvs; r[2] = ztop; r[3] = h[i];
      *reinterpret_cast<double4*>(rows + 4 * i) = make_double4(r[0], r[1], r[2], r[3]);
      ztop = ztop + h[i];
    }
  }
}

__device__ __forceinline__ void store_row(const SamplerDev& p, int b, long long iiter) {
  const int n = p.nstored[b];
  if (n >= p.S) {
    // the chain arrays of this chain are full: count, and remember WHEN, so that the dwell time of the last
    // stored model ends here instead of absorbing every later iteration (save_chain_files)
    atomicAdd(p.overflow, 1ULL);
    if (p.ovf_count[b]++ == 0) p.ovf_iter[b] = iiter;
    return;
  }
  const int maxl = p.maxl, T = p.T, k = p.cur_k[b];
  const double* cm = p.cur_model + (size_t)b * 2 * maxl;
  float* m = p.st_models + ((size_t)b * p.S + n) * 2 * maxl;
  // chainmodels[n, :model.size] = model: vs then z, contiguous, NaN beyond (SingleChain.py:501)
  for (int i = 0; i < k; ++i) { m[i] = (float)cm[i]; m[k + i] = (float)cm[maxl + i]; }
  for (int i = 2 * k; i < 2 * maxl; ++i) m[i] = __int_as_float(0x7fc00000);
  for (int t = 0; t <= T; ++t)
    p.st_misfits[((size_t)b * p.S + n) * (T + 1) + t] = (float)p.cur_misfits[(size_t)b * (T + 1) + t];
  p.st_likes[(size_t)b * p.S + n] = (float)p.cur_logL[b];
  for (int i = 0; i < 2 * T; ++i)
    p.st_noise[((size_t)b * p.S + n) * 2 * T + i] = (float)p.cur_noise[(size_t)b * 2 * T + i];
  p.st_vpvs[(size_t)b * p.S + n] = (float)p.cur_vpvs[b];
  p.st_iter[(size_t)b * p.S + n] = (int)iiter;
  p.nstored[b] = n + 1;
}

__device__ __forceinline__ void take_proposal(const SamplerDev& p, int b) {
  const int maxl = p.maxl, T = p.T, k = p.prop_k[b];
  for (int i = 0; i < k; ++i) {
    p.cur_model[(size_t)b * 2 * maxl + i] = p.prop_model[(size_t)b * 2 * maxl + i];
    p.cur_model[(size_t)b * 2 * maxl + maxl + i] = p.prop_model[(size_t)b * 2 * maxl + maxl + i];
  }
  p.cur_k[b] = k;
  p.cur_vpvs[b] = p.prop_vpvs[b];
  for (int i = 0; i < 2 * T; ++i) p.cur_noise[(size_t)b * 2 * T + i] = p.prop_noise[(size_t)b * 2 * T + i];
  p.cur_logL[b] = p.p_logL[b];
  for (int t = 0; t <= T; ++t) p.cur_misfits[(size_t)b * (T + 1) + t] = p.p_misfits[(size_t)b * (T + 1) + t];
}

__global__ void __launch_bounds__(64)
sampler_accept_kernel(SamplerDev p) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= p.B) return;
  const long long iiter = p.iiter[b];
  if (!p.prop_valid[b]) { p.iiter[b] = iiter + 1; return; }      // :541-547
  const int modify = p.prop_modify[b];
  const int par = sampler_paridx(modify);
  long long* proposed = p.proposed + (size_t)b * SMP_NPAR;
  long long* accepted = p.accepted + (size_t)b * SMP_NPAR;
  double* propdist = p.propdist + (size_t)b * SMP_NPAR;
  proposed[par] += 1;
  const Draw d = chain_draw(p, b, iiter);
  const double u = log(d.u_acc);                                  // :556
  const double alpha = sampler_alpha(p.cfg, modify, propdist, p.prop_dvs2[b], p.p_logL[b], p.cur_logL[b]);
  if (u < alpha) {                                                // :560-564
    take_proposal(p, b);
    store_row(p, b, iiter);
    accepted[par] += 1;
  }
  if (iiter % 1000 == 0) {                                        // :585-587
    bool all = true;
    for (int i = 0; i < SMP_NPAR; ++i) all = all && proposed[i] != 0;
    if (all) sampler_adjust_propdist(p.cfg, propdist, accepted, proposed);
  }
  p.iiter[b] = iiter + 1;
}

// initial model of every chain: pack the rows of the current state for the first evaluation
__global__ void sampler_pack_current_kernel(SamplerDev p) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= p.B) return;
  const int maxl = p.maxl, k = p.cur_k[b];
  double vs[SMP_MAX_ROWS], z[SMP_MAX_ROWS], h[SMP_MAX_ROWS];
  for (int i = 0; i < k; ++i) { vs[i] = p.cur_model[(size_t)b * 2 * maxl + i]; z[i] = p.cur_model[(size_t)b * 2 * maxl + maxl + i]; }
  sampler_thickness(z, k, h);
  sampler_pack_rows(p.cfg, vs, h, k, p.cur_vpvs[b], p.rows + (size_t)b * maxl * 4);
  p.nlay[b] = k;
  p.prop_k[b] = k;
  p.prop_vpvs[b] = p.cur_vpvs[b];
  for (int i = 0; i < k; ++i) {
    p.prop_model[(size_t)b * 2 * maxl + i] = vs[i];
    p.prop_model[(size_t)b * 2 * maxl + maxl + i] = z[i];
  }
  for (int i = 0; i < 2 * p.T; ++i) p.prop_noise[(size_t)b * 2 * p.T + i] = p.cur_noise[(size_t)b * 2 * p.T + i];
}

// accept_as_currentmodel + append_currentmodel of the initial model (SingleChain.py:88-92)
__global__ void sampler_init_accept_kernel(SamplerDev p) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= p.B) return;
  take_proposal(p, b);
  store_row(p, b, p.iiter[b]);
}

template <class T>
int salloc(bh_sampler* s, T** ptr, size_t n) {
  *ptr = nullptr;
  if (n == 0) n = 1;
  cudaError_t ce = cudaMalloc((void**)ptr, n * sizeof(T));
  if (ce != cudaSuccess) return bh_set_error_message(BH_ERR_CUDA, cudaGetErrorString(ce));
  s->owned.push_back(*ptr);
  return BH_OK;
}

SamplerDev dev_view(const bh_sampler* s) {
  SamplerDev p;
  p.cfg = s->cfg; p.B = s->B; p.S = s->S; p.T = s->T; p.maxl = s->maxl; p.first_chain = s->first_chain;
  p.cur_model = s->cur_model; p.cur_vpvs = s->cur_vpvs; p.cur_noise = s->cur_noise; p.cur_logL = s->cur_logL;
  p.cur_misfits = s->cur_misfits; p.propdist = s->propdist; p.accepted = s->accepted; p.proposed = s->proposed;
  p.iiter = s->iiter; p.cur_k = s->cur_k; p.nstored = s->nstored;
  p.prop_model = s->prop_model; p.prop_vpvs = s->prop_vpvs; p.prop_noise = s->prop_noise; p.prop_dvs2 = s->prop_dvs2;
  p.rows = s->rows; p.p_logL = s->p_logL; p.p_misfits = s->p_misfits;
  p.forced = s->use_forced ? s->forced : nullptr;
  p.prop_k = s->prop_k; p.prop_valid = s->prop_valid; p.prop_modify = s->prop_modify; p.nlay = s->nlay;
  p.p_status = s->p_status;
  p.st_models = s->st_models; p.st_misfits = s->st_misfits; p.st_likes = s->st_likes; p.st_noise = s->st_noise;
  p.st_vpvs = s->st_vpvs; p.st_iter = s->st_iter; p.overflow = s->overflow; p.ovf_count = s->ovf_count; p.ovf_iter = s->ovf_iter;
  return p;
}

#define SMP_CUDA(call)                                                                   \
  do {                                                                                   \
    cudaError_t _e = (call);                                                             \
    if (_e != cudaSuccess) return bh_set_error_message(BH_ERR_CUDA, cudaGetErrorString(_e)); \
  } while (0)

}  // namespace

extern "C" {

void bh_sampler_destroy(bh_sampler* s) {
  if (!s) return;
  for (void* p : s->owned) cudaFree(p);
  if (s->st) cudaStreamDestroy(s->st);
  delete s;
}

int bh_sampler_create(bh_engine* e, const bh_sampler_config* c, int ntargets, int nchains, long long first_chain,
                      bh_sampler** out) {
  if (!out) return bh_set_error_message(BH_ERR_ARG, "out is null");
  *out = nullptr;
  if (!e || !c) return bh_set_error_message(BH_ERR_ARG, "null engine/config");
  if (ntargets < 1 || ntargets > BH_MAX_TARGETS) return bh_set_error_message(BH_ERR_ARG, "ntargets out of range");
  if (nchains < 1) return bh_set_error_message(BH_ERR_ARG, "nchains must be >= 1");
  if (c->layers_min < 0 || c->layers_max < c->layers_min || c->layers_max + 1 > BH_MAX_LAYERS)
    return bh_set_error_message(BH_ERR_ARG, "layers prior must satisfy 0 <= min <= max <= 99");
  if (c->max_accepted < 1) return bh_set_error_message(BH_ERR_ARG, "max_accepted must be >= 1");
  if (bh_device_count() < 1) return bh_set_error_message(BH_ERR_NO_DEVICE, "no CUDA device visible; this library has no CPU path");
  bh_sampler* s = new bh_sampler();
  s->eng = e; s->B = nchains; s->S = c->max_accepted; s->T = ntargets; s->maxl = c->layers_max + 1;
  s->first_chain = first_chain;
  s->cfg = sampler_cfg_from_public(*c, ntargets);
  bh_engine_set(e, "swd_autotune", 1);     // a sampler evaluates the same batch size thousands of times
  const size_t B = (size_t)nchains, L = (size_t)s->maxl, T = (size_t)ntargets, S = (size_t)s->S;
  int rc = BH_OK;
#define A(p, n) if (rc == BH_OK) rc = salloc(s, &s->p, (n))
  A(cur_model, B * 2 * L); A(cur_vpvs, B); A(cur_noise, B * 2 * T); A(cur_logL, B); A(cur_misfits, B * (T + 1));
  A(propdist, B * SMP_NPAR); A(accepted, B * SMP_NPAR); A(proposed, B * SMP_NPAR); A(iiter, B);
  A(cur_k, B); A(nstored, B);
  A(prop_model, B * 2 * L); A(prop_vpvs, B); A(prop_noise, B * 2 * T); A(prop_dvs2, B); A(rows, B * L * 4);
  A(p_logL, B); A(p_misfits, B * (T + 1)); A(forced, B * 4);
  A(prop_k, B); A(prop_valid, B); A(prop_modify, B); A(nlay, B); A(p_status, B);
  A(st_models, B * S * 2 * L); A(st_misfits, B * S * (T + 1)); A(st_likes, B * S); A(st_noise, B * S * 2 * T);
  A(st_vpvs, B * S); A(st_iter, B * S); A(overflow, 1); A(ovf_count, B); A(ovf_iter, B);
#undef A
  if (rc == BH_OK && cudaStreamCreateWithFlags(&s->st, cudaStreamNonBlocking) != cudaSuccess)
    rc = bh_set_error_message(BH_ERR_CUDA, "stream creation");
  if (rc == BH_OK) {
    // NaN-fill the chain arrays (mcmcOptimizer.py:92-125), zero the counters, seed propdist
    cudaMemsetAsync(s->st_models, 0xff, B * S * 2 * L * sizeof(float), s->st);
    cudaMemsetAsync(s->st_misfits, 0xff, B * S * (T + 1) * sizeof(float), s->st);
    cudaMemsetAsync(s->st_likes, 0xff, B * S * sizeof(float), s->st);
    cudaMemsetAsync(s->st_noise, 0xff, B * S * 2 * T * sizeof(float), s->st);
    cudaMemsetAsync(s->st_vpvs, 0xff, B * S * sizeof(float), s->st);
    cudaMemsetAsync(s->st_iter, 0, B * S * sizeof(int), s->st);
    cudaMemsetAsync(s->accepted, 0, B * SMP_NPAR * sizeof(long long), s->st);
    cudaMemsetAsync(s->proposed, 0, B * SMP_NPAR * sizeof(long long), s->st);
    cudaMemsetAsync(s->nstored, 0, B * sizeof(int), s->st);
    cudaMemsetAsync(s->overflow, 0, sizeof(unsigned long long), s->st);
    cudaMemsetAsync(s->ovf_count, 0, B * sizeof(long long), s->st);
    cudaMemsetAsync(s->ovf_iter, 0, B * sizeof(long long), s->st);
    cudaMemsetAsync(s->cur_model, 0, B * 2 * L * sizeof(double), s->st);
    cudaMemsetAsync(s->prop_model, 0, B * 2 * L * sizeof(double), s->st);
    std::vector<double> pd(B * SMP_NPAR);
    std::vector<long long> it(B, -(long long)c->iter_burnin);
    for (size_t b = 0; b < B; ++b) for (int i = 0; i < SMP_NPAR; ++i) pd[b * SMP_NPAR + i] = c->propdist[i];
    if (cudaMemcpyAsync(s->propdist, pd.data(), pd.size() * sizeof(double), cudaMemcpyHostToDevice, s->st) != cudaSuccess ||
        cudaMemcpyAsync(s->iiter, it.data(), it.size() * sizeof(long long), cudaMemcpyHostToDevice, s->st) != cudaSuccess ||
        cudaStreamSynchronize(s->st) != cudaSuccess)
      rc = bh_set_error_message(BH_ERR_CUDA, "sampler initialisation");
  }
  if (rc != BH_OK) { std::string keep = bh_last_error(); bh_sampler_destroy(s); bh_set_error_message(rc, keep.c_str()); return rc; }
  *out = s;
  return BH_OK;
}

int bh_sampler_set_state(bh_sampler* s, const double* models, const int* k, const double* vpvs, const double* noise,
                         const double* logL, const double* misfits, const double* propdist,
                         const long long* accepted, const long long* proposed, const long long* iiter) {
  if (!s || !models || !k || !vpvs || !noise) return bh_set_error_message(BH_ERR_ARG, "null argument");
  const size_t B = s->B, L = s->maxl, T = s->T;
  for (size_t b = 0; b < B; ++b)
    if (k[b] < 1 || k[b] > (int)L) return bh_set_error_message(BH_ERR_ARG, "k must be 1..layers_max+1");
  SMP_CUDA(cudaMemcpyAsync(s->cur_model, models, B * 2 * L * sizeof(double), cudaMemcpyHostToDevice, s->st));
  SMP_CUDA(cudaMemcpyAsync(s->cur_k, k, B * sizeof(int), cudaMemcpyHostToDevice, s->st));
  SMP_CUDA(cudaMemcpyAsync(s->cur_vpvs, vpvs, B * sizeof(double), cudaMemcpyHostToDevice, s->st));
  SMP_CUDA(cudaMemcpyAsync(s->cur_noise, noise, B * 2 * T * sizeof(double), cudaMemcpyHostToDevice, s->st));
  if (logL) SMP_CUDA(cudaMemcpyAsync(s->cur_logL, logL, B * sizeof(double), cudaMemcpyHostToDevice, s->st));
  if (misfits) SMP_CUDA(cudaMemcpyAsync(s->cur_misfits, misfits, B * (T + 1) * sizeof(double), cudaMemcpyHostToDevice, s->st));
  if (propdist) SMP_CUDA(cudaMemcpyAsync(s->propdist, propdist, B * SMP_NPAR * sizeof(double), cudaMemcpyHostToDevice, s->st));
  if (accepted) SMP_CUDA(cudaMemcpyAsync(s->accepted, accepted, B * SMP_NPAR * sizeof(long long), cudaMemcpyHostToDevice, s->st));
  if (proposed) SMP_CUDA(cudaMemcpyAsync(s->proposed, proposed, B * SMP_NPAR * sizeof(long long), cudaMemcpyHostToDevice, s->st));
  if (iiter) SMP_CUDA(cudaMemcpyAsync(s->iiter, iiter, B * sizeof(long long), cudaMemcpyHostToDevice, s->st));
  SMP_CUDA(cudaStreamSynchronize(s->st));
  return BH_OK;
}

int bh_sampler_init(bh_sampler* s, const double* models, const int* k, const double* vpvs, const double* noise) {
  int rc = bh_sampler_set_state(s, models, k, vpvs, noise, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
  if (rc != BH_OK) return rc;
  const SamplerDev p = dev_view(s);
  const int threads = 64, blocks = (s->B + threads - 1) / threads;
  sampler_pack_current_kernel<<<blocks, threads, 0, s->st>>>(p);
  rc = bh_engine_eval(s->eng, s->rows, s->nlay, s->prop_noise, nullptr, s->B, s->maxl, s->p_logL, s->p_misfits,
                      s->p_status, nullptr, s->st);
  if (rc != BH_OK) return rc;
  sampler_init_accept_kernel<<<blocks, threads, 0, s->st>>>(p);
  SMP_CUDA(cudaGetLastError());
  SMP_CUDA(cudaStreamSynchronize(s->st));
  return BH_OK;
}

int bh_sampler_set_forced_draws(bh_sampler* s, const double* draws) {
  if (!s) return bh_set_error_message(BH_ERR_ARG, "null sampler");
  if (!draws) { s->use_forced = 0; return BH_OK; }
  SMP_CUDA(cudaMemcpyAsync(s->forced, draws, (size_t)s->B * 4 * sizeof(double), cudaMemcpyHostToDevice, s->st));
  SMP_CUDA(cudaStreamSynchronize(s->st));
  s->use_forced = 1;
  return BH_OK;
}

// One lock-step iteration, enqueued on s->st: propose, evaluate, accept.
static int enqueue_iteration(bh_sampler* s, const SamplerDev& p) {
  const int threads = 64, blocks = (s->B + threads - 1) / threads;
  sampler_propose_kernel<<<blocks, threads, 0, s->st>>>(p);
  int rc = bh_engine_eval(s->eng, s->rows, s->nlay, s->prop_noise, nullptr, s->B, s->maxl, s->p_logL,
                          s->p_misfits, s->p_status, nullptr, s->st);
  if (rc != BH_OK) return rc;
  sampler_accept_kernel<<<blocks, threads, 0, s->st>>>(p);
  return BH_OK;
}

// An iteration is ~15 dependent launches on three streams; replayed from a CUDA graph the gaps between them go
// (tutorial ensemble of 512 chains: 1.39 -> ~1.1 ms per iteration).  Every kGraphChunk iterations one is enqueued
// plainly, which lets the engine refresh what it adapts (record capacity, layout) before the next capture.
// BH_SAMPLER_GRAPH=0 switches the replay off.
int bh_sampler_run(bh_sampler* s, int niter) {
  if (!s || niter < 0) return bh_set_error_message(BH_ERR_ARG, "bad argument");
  const SamplerDev p = dev_view(s);
  static const bool graphs = []() { const char* v = getenv("BH_SAMPLER_GRAPH"); return !(v && v[0] == '0'); }();
  constexpr int kGraphChunk = 128;
  std::vector<cudaGraphExec_t> execs;
  std::vector<cudaGraph_t> captured;
  int rc = BH_OK;
  cudaError_t launch_err = cudaSuccess;
  for (int it = 0; it < niter && rc == BH_OK;) {
    rc = enqueue_iteration(s, p);
    ++it;
    if (rc != BH_OK) break;
    // while the engine is still timing its models-per-warp candidates (the first ~10 iterations of a
    // batch size) let each iteration finish, so that its timing is read before the next one is enqueued
    if (bh_engine_is_tuning(s->eng)) { SMP_CUDA(cudaStreamSynchronize(s->st)); continue; }
    const int n = niter - it < kGraphChunk ? niter - it : kGraphChunk;
    if (!graphs || n < 8 || bh_engine_capture_begin(s->eng) != BH_OK) continue;
    cudaGraph_t g = nullptr;
    cudaGraphExec_t ex = nullptr;
    cudaError_t ce = cudaStreamBeginCapture(s->st, cudaStreamCaptureModeThreadLocal);
    if (ce == cudaSuccess) {
      rc = enqueue_iteration(s, p);
      ce = cudaStreamEndCapture(s->st, &g);
    }
    bh_engine_capture_end(s->eng);
    if (ce == cudaSuccess && rc == BH_OK && g) ce = cudaGraphInstantiate(&ex, g, 0);
    if (g) captured.push_back(g);
    if (ce != cudaSuccess || rc != BH_OK || !ex) {       // no graph: the plain path goes on
      cudaGetLastError();
      if (rc != BH_OK) break;
      continue;
    }
    execs.push_back(ex);
    for (int k = 0; k < n && ce == cudaSuccess; ++k) ce = cudaGraphLaunch(ex, s->st);
    if (ce != cudaSuccess) { launch_err = ce; break; }
    it += n;
  }
  cudaError_t ce = launch_err != cudaSuccess ? launch_err : cudaGetLastError();
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(s->st);
  else cudaStreamSynchronize(s->st);
  for (cudaGraphExec_t ex : execs) cudaGraphExecDestroy(ex);
  for (cudaGraph_t g : captured) cudaGraphDestroy(g);
  if (rc != BH_OK) return rc;
  SMP_CUDA(ce);
  return BH_OK;
}

int bh_sampler_get_state(bh_sampler* s, double* models, int* k, double* vpvs, double* noise, double* logL,
                         double* misfits, double* propdist, long long* accepted, long long* proposed,
                         long long* iiter, int* nstored, long long* overflow) {
  if (!s) return bh_set_error_message(BH_ERR_ARG, "null sampler");
  const size_t B = s->B, L = s->maxl, T = s->T;
  cudaStream_t st = s->st;
#define G(dst, src, n) if (dst) SMP_CUDA(cudaMemcpyAsync(dst, s->src, (n), cudaMemcpyDeviceToHost, st))
  G(models, cur_model, B * 2 * L * sizeof(double)); G(k, cur_k, B * sizeof(int)); G(vpvs, cur_vpvs, B * sizeof(double));
  G(noise, cur_noise, B * 2 * T * sizeof(double)); G(logL, cur_logL, B * sizeof(double));
  G(misfits, cur_misfits, B * (T + 1) * sizeof(double)); G(propdist, propdist, B * SMP_NPAR * sizeof(double));
  G(accepted, accepted, B * SMP_NPAR * sizeof(long long)); G(proposed, proposed, B * SMP_NPAR * sizeof(long long));
  G(iiter, iiter, B * sizeof(long long)); G(nstored, nstored, B * sizeof(int));
  G(overflow, overflow, sizeof(long long));
#undef G
  SMP_CUDA(cudaStreamSynchronize(st));
  return BH_OK;
}

int bh_sampler_get_overflow(bh_sampler* s, long long* count, long long* first_iter) {
  if (!s) return bh_set_error_message(BH_ERR_ARG, "null sampler");
  const size_t B = s->B;
  if (count) SMP_CUDA(cudaMemcpyAsync(count, s->ovf_count, B * sizeof(long long), cudaMemcpyDeviceToHost, s->st));
  if (first_iter) SMP_CUDA(cudaMemcpyAsync(first_iter, s->ovf_iter, B * sizeof(long long), cudaMemcpyDeviceToHost, s->st));
  SMP_CUDA(cudaStreamSynchronize(s->st));
  return BH_OK;
}

int bh_sampler_get_proposal(bh_sampler* s, double* models, int* k, double* vpvs, double* noise, int* valid,
                            int* modify, double* dvs2, double* logL, double* misfits) {
  if (!s) return bh_set_error_message(BH_ERR_ARG, "null sampler");
  const size_t B = s->B, L = s->maxl, T = s->T;
  cudaStream_t st = s->st;
#define G(dst, src, n) if (dst) SMP_CUDA(cudaMemcpyAsync(dst, s->src, (n), cudaMemcpyDeviceToHost, st))
  G(models, prop_model, B * 2 * L * sizeof(double)); G(k, prop_k, B * sizeof(int)); G(vpvs, prop_vpvs, B * sizeof(double));
  G(noise, prop_noise, B * 2 * T * sizeof(double)); G(valid, prop_valid, B * sizeof(int));
  G(modify, prop_modify, B * sizeof(int)); G(dvs2, prop_dvs2, B * sizeof(double));
  G(logL, p_logL, B * sizeof(double)); G(misfits, p_misfits, B * (T + 1) * sizeof(double));
#undef G
  SMP_CUDA(cudaStreamSynchronize(st));
  return BH_OK;
}

int bh_sampler_get_chains(bh_sampler* s, int chain0, int nchain, float* models, float* misfits, float* likes,
                          float* noise, float* vpvs, int* iters) {
  if (!s || chain0 < 0 || nchain < 0 || chain0 + nchain > s->B) return bh_set_error_message(BH_ERR_ARG, "chain range");
  const size_t c0 = chain0, n = nchain, L = s->maxl, T = s->T, S = s->S;
  cudaStream_t st = s->st;
  if (models) SMP_CUDA(cudaMemcpyAsync(models, s->st_models + c0 * S * 2 * L, n * S * 2 * L * sizeof(float), cudaMemcpyDeviceToHost, st));
  if (misfits) SMP_CUDA(cudaMemcpyAsync(misfits, s->st_misfits + c0 * S * (T + 1), n * S * (T + 1) * sizeof(float), cudaMemcpyDeviceToHost, st));
  if (likes) SMP_CUDA(cudaMemcpyAsync(likes, s->st_likes + c0 * S, n * S * sizeof(float), cudaMemcpyDeviceToHost, st));
  if (noise) SMP_CUDA(cudaMemcpyAsync(noise, s->st_noise + c0 * S * 2 * T, n * S * 2 * T * sizeof(float), cudaMemcpyDeviceToHost, st));
  if (vpvs) SMP_CUDA(cudaMemcpyAsync(vpvs, s->st_vpvs + c0 * S, n * S * sizeof(float), cudaMemcpyDeviceToHost, st));
  if (iters) SMP_CUDA(cudaMemcpyAsync(iters, s->st_iter + c0 * S, n * S * sizeof(int), cudaMemcpyDeviceToHost, st));
  SMP_CUDA(cudaStreamSynchronize(st));
  return BH_OK;
}

}  // extern "C"
