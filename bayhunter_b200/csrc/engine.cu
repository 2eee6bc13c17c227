// engine.cu -- C ABI of libbayhunter_b200.so (see include/bayhunter_b200.h).
//
// Owns the device-resident constants of a joint target set, the per-batch
// scratch, and the launch schedule of one joint evaluation:
//
//   stream S  : prepare(SWD rows) -> swd_kernel ----------------------+
//   stream A  : prepare(RF tables) -> rf_spectrum -> rf_synth --(join)-+-> loglik
//
// The SWD kernel is latency bound (few, long, serial searches); the RF spectrum
// kernel is throughput bound ((model, frequency) items by the million).  Forking
// them onto two streams lets the RF warps fill the issue slots the SWD warps
// leave idle.  There is no CPU fallback anywhere in this file.
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>
#include <stdio.h>
#include <string.h>

#include <mutex>
#include <string>
#include <vector>

#include "../../include/bayhunter_b200.h"
#include "kernels.h"
#include "bh_math.cuh"

using namespace bh;

namespace {

thread_local std::string g_err;

int set_err(int code, const char* what, cudaError_t ce = cudaSuccess) {
  g_err = what;
  if (ce != cudaSuccess) {
    g_err += ": ";
    g_err += cudaGetErrorString(ce);
  }
  return code;
}

}  // namespace
namespace bh {
// shared with sampler.cu: record a message for bh_last_error() and hand the code back
int bh_set_error_message(int code, const char* what) { return set_err(code, what); }
}  // namespace bh
namespace {

#define BH_CUDA(call)                                                        \
  do {                                                                       \
    cudaError_t _e = (call);                                                 \
    if (_e != cudaSuccess) return set_err(BH_ERR_CUDA, #call, _e);           \
  } while (0)

template <class T>
int dev_alloc(T** p, size_t n) {
  *p = nullptr;
  if (n == 0) return BH_OK;
  BH_CUDA(cudaMalloc((void**)p, n * sizeof(T)));
  return BH_OK;
}

int odd_stride(int lmax) { return (lmax & 1) ? lmax : lmax + 1; }

bool is_swd(int ref) { return ref >= BH_REF_RDISPPH && ref <= BH_REF_LDISPGR; }
bool is_rf(int ref) { return ref == BH_REF_PRF || ref == BH_REF_SRF; }

}  // namespace

struct bh_engine {
  TargetSet ts{};
  int max_batch = 0, max_layers = 0;
  std::vector<void*> owned;   // device allocations freed on destroy
  // scratch
  PrepOut prep{};
  cd* spec = nullptr;
  double* curves = nullptr;
  double* roots = nullptr;
  int curve_stride = 0;
  int curve_off[kMaxTargets] = {0};
  double* rfsynth = nullptr;
  int* tstatus = nullptr;
  bool has_generic = false;   // a target without a forward model of this library: bh_engine_loglik_host only
  double* gauss_phi = nullptr;
  double *gauss_res = nullptr, *gauss_part = nullptr;
  unsigned long long* counters = nullptr;
  int* swd_queue = nullptr;   // work-item counters of the mixed dispersion launch
  int* swd_perm = nullptr;    // [max_batch] models ordered by layer count (ragged batches)
  // Largest layer count of the batches seen lately: read back asynchronously (never waited for) and
  // used to size the dispersion kernel's shared-memory records for the NEXT evaluations; models that
  // exceed it are caught by a second launch with full capacity, so a stale value is never wrong.
  int* d_maxn = nullptr;
  int* h_maxn = nullptr;      // pinned
  cudaEvent_t ev_maxn = nullptr;
  bool maxn_pending = false;
  int last_maxn = -1;
  int adaptive_lcap = 1;
  // models-per-warp autotuning ("swd_autotune"): candidates around the rule's pick, timed in place
  struct Tune {
    int B = -1, base = -1, ncand = 0, cur = 0, locked = -1;
    int cand[3] = {0, 0, 0}, trials[3] = {0, 0, 0};
    float best_ms[3] = {0.f, 0.f, 0.f};
    bool pending = false;
  } tune;
  int autotune = 0;
  cudaEvent_t ev_tune[2] = {nullptr, nullptr};
  int sort_layers = 1;        // deal models to dispersion warps in that order
  // The RF stream is released once this share of the dispersion warps has retired (0: no gate).  The
  // evaluation is bound by the SUM of fp64 work, but the RF kernels must not get onto the SMs BEFORE
  // the dispersion warps (which stream wins that race depends on what ran before): they would hold the
  // registers the latency-critical searches need.
  int rf_gate_pct = 25;
  int* swd_done = nullptr;    // retired dispersion warps of the current evaluation
  // Spectral bins whose Gauss-filter weight exp(-(w/2a)^2) is below this are not computed: next to a
  // trace peak of order 0.1-1 they are below the resolution of fp64 (1e-30 vs 2e-16).  0 = all bins.
  double rf_floor = 1e-20;    // Gauss-filter weight below which a spectral bin is not computed (4 orders below fp64 resolution of the trace)
  int nsm = 0;
  int max_nfreq = 0;
  // Host-pointer entry points: two slots of pinned staging + device mirrors, so that the copies of call
  // k + 1 (stream s_h2d) and of call k (stream s_d2h) run beside the kernels of the call in between.
  struct HostSlot {
    double *h_model = nullptr, *h_noise = nullptr, *h_rho = nullptr, *h_logL = nullptr, *h_misfits = nullptr,
           *h_synth = nullptr;
    int *h_nlay = nullptr, *h_status = nullptr;
    double *d_model = nullptr, *d_noise = nullptr, *d_rho = nullptr, *d_logL = nullptr, *d_misfits = nullptr,
           *d_synth = nullptr;
    int *d_nlay = nullptr, *d_status = nullptr;
    cudaEvent_t ev_in = nullptr, ev_kernels = nullptr, ev_done = nullptr;
    bool busy = false;
    long long ticket = 0;
    // delivery of a staged result
    double *out_logL = nullptr, *out_misfits = nullptr, *out_synth = nullptr;
    int* out_status = nullptr;
    int B = 0;
    bool staged_out = false;
  } slot[2];
  bool slots_ready = false;
  long long next_ticket = 1;
  cudaStream_t s_h2d = nullptr, s_d2h = nullptr;
  cudaStream_t s_own = nullptr, s_aux = nullptr, s_aux2 = nullptr, s_aux3 = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_join2 = nullptr, ev_fork2 = nullptr, ev_join3 = nullptr, ev_fork3 = nullptr;
  // tunables
  int searches_per_warp = 0;  // phase-velocity curves; 0 = auto
  int group_spw = 0;          // group-velocity curves; 0 = half of the above
  int max_spec = 16;   // candidates per walking chain and round (8 -> 16: -3..5 % at B <= 512, neutral for a full batch)
  int spw_curve[4] = {0, 0, 0, 0};   // per curve type override: Rayleigh group, Rayleigh phase, Love group, Love phase
  int split_waves = 0;        // 1: Rayleigh and Love curves in separate launches (two streams)
  int rayleigh_sm_pct = 0;    // mixed launch: share of the SMs dedicated to the Rayleigh items (0 = no partition)
  int direct = 0;             // 0 never, 1 when warps are full of chains, 2 always
  int lockstep = 0;           // swd_lockstep_kernel (every lane owns a chain, pairwise guesses): 0 off, 1 on.  Measured
                              // equal or slower than swd_kernel on every BASELINE configuration (profiles/r02_swd_restructure.txt)
  int graph_safe = 0;         // between bh_engine_capture_begin / _end: evaluations enqueue stream-ordered work only
  int pool = -1;              // swd_pool_kernel (a CTA's 128 lanes dealt over the chains of ~28 models): 0 off, 1 on, -1 rule (full batches)
  int pool_models = 0;        // models per CTA of the pool kernel (0 = rule)
  int ls_spw[2] = {0, 0};     // lockstep kernel: models per warp of group / phase curves (0 = rule)
  int concurrent = 1;
  // optional per-kernel timing (bh_engine_set "profile"): event pairs around
  // each launch, recorded on the stream the kernel is launched on
  int profile = 0;
  cudaEvent_t pev[2 * BH_NUM_KERNELS] = {nullptr};
  bool pev_used[BH_NUM_KERNELS] = {false};
};

namespace {
const char* const kRangeNames[BH_NUM_KERNELS] = {"bh:prepare_swd", "bh:swd", "bh:prepare_rf", "bh:rf_spectrum",
                                                 "bh:rf_synth", "bh:loglik", "bh:swd_love", "bh:swd_general",
                                                 "bh:swd_pool", "bh:swd_pool_love"};
struct KTimer {   // an NVTX range around the enqueue of every kernel group; with profiling on also a start / stop event pair
  bh_engine* e; int k; cudaStream_t st;
  KTimer(bh_engine* e_, int k_, cudaStream_t st_) : e(e_), k(k_), st(st_) {
    nvtxRangePushA(kRangeNames[k]);
    if (e->profile) { cudaEventRecord(e->pev[2 * k], st); e->pev_used[k] = true; }
  }
  ~KTimer() { if (e->profile) cudaEventRecord(e->pev[2 * k + 1], st); nvtxRangePop(); }
};
struct Range { explicit Range(const char* n) { nvtxRangePushA(n); } ~Range() { nvtxRangePop(); } };
}  // namespace

static int upload(bh_engine* e, const double* host, size_t n, const double** dev) {
  double* d = nullptr;
  int rc = dev_alloc(&d, n);
  if (rc != BH_OK) return rc;
  e->owned.push_back(d);
  BH_CUDA(cudaMemcpy(d, host, n * sizeof(double), cudaMemcpyHostToDevice));
  *dev = d;
  return BH_OK;
}

template <class T>
static int scratch(bh_engine* e, T** p, size_t n) {
  int rc = dev_alloc(p, n);
  if (rc == BH_OK && *p) e->owned.push_back(*p);
  return rc;
}

extern "C" {

int bh_abi_version(void) { return BH_ABI_VERSION; }

const char* bh_last_error(void) { return g_err.c_str(); }

int bh_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

int bh_set_device(int device) {
  if (bh_device_count() < 1) return set_err(BH_ERR_NO_DEVICE, "no CUDA device visible; this library has no CPU path");
  BH_CUDA(cudaSetDevice(device));
  return BH_OK;
}

void bh_engine_destroy(bh_engine* e) {
  if (!e) return;
  for (void* p : e->owned) cudaFree(p);
  for (cudaEvent_t ev : e->pev) if (ev) cudaEventDestroy(ev);
  if (e->ev_fork) cudaEventDestroy(e->ev_fork);
  if (e->ev_join) cudaEventDestroy(e->ev_join);
  if (e->ev_join2) cudaEventDestroy(e->ev_join2);
  if (e->ev_fork2) cudaEventDestroy(e->ev_fork2);
  if (e->ev_join3) cudaEventDestroy(e->ev_join3);
  if (e->ev_fork3) cudaEventDestroy(e->ev_fork3);
  if (e->ev_maxn) cudaEventDestroy(e->ev_maxn);
  for (cudaEvent_t ev : e->ev_tune) if (ev) cudaEventDestroy(ev);
  if (e->h_maxn) cudaFreeHost(e->h_maxn);
  for (auto& sl : e->slot) {
    void* hp[] = {sl.h_model, sl.h_noise, sl.h_rho, sl.h_logL, sl.h_misfits, sl.h_synth, sl.h_nlay, sl.h_status};
    for (void* q : hp) if (q) cudaFreeHost(q);
    void* dp[] = {sl.d_model, sl.d_noise, sl.d_rho, sl.d_logL, sl.d_misfits, sl.d_synth, sl.d_nlay, sl.d_status};
    for (void* q : dp) if (q) cudaFree(q);
    if (sl.ev_in) cudaEventDestroy(sl.ev_in);
    if (sl.ev_kernels) cudaEventDestroy(sl.ev_kernels);
    if (sl.ev_done) cudaEventDestroy(sl.ev_done);
  }
  if (e->s_h2d) cudaStreamDestroy(e->s_h2d);
  if (e->s_d2h) cudaStreamDestroy(e->s_d2h);
  if (e->s_own) cudaStreamDestroy(e->s_own);
  if (e->s_aux) cudaStreamDestroy(e->s_aux);
  if (e->s_aux2) cudaStreamDestroy(e->s_aux2);
  if (e->s_aux3) cudaStreamDestroy(e->s_aux3);
  delete e;
}

int bh_engine_create(const bh_target* targets, int ntargets, int max_batch, int max_layers,
                     bh_engine** out) {
  if (!out) return set_err(BH_ERR_ARG, "out is null");
  *out = nullptr;
  if (!targets || ntargets < 1 || ntargets > BH_MAX_TARGETS)
    return set_err(BH_ERR_ARG, "ntargets must be 1..BH_MAX_TARGETS");
  if (max_batch < 1 || max_layers < 1 || max_layers > BH_MAX_LAYERS)
    return set_err(BH_ERR_ARG, "max_batch >= 1 and 1 <= max_layers <= 100 required");
  if (bh_device_count() < 1)
    return set_err(BH_ERR_NO_DEVICE, "no CUDA device visible; this library has no CPU path");

  bh_engine* e = new bh_engine();
  e->max_batch = max_batch;
  e->max_layers = max_layers;
  e->ts.ntargets = ntargets;
  int off = 0, coff = 0, nrf = 0;
  int rc = BH_OK;
  for (int t = 0; t < ntargets && rc == BH_OK; ++t) {
    const bh_target& s = targets[t];
    TargetDev& d = e->ts.t[t];
    memset(&d, 0, sizeof(d));
    if (s.n < 1 || !s.x || !s.y) { rc = set_err(BH_ERR_ARG, "target needs n >= 1, x and y"); break; }
    if (!is_swd(s.ref) && !is_rf(s.ref) && s.ref != BH_REF_GENERIC) { rc = set_err(BH_ERR_ARG, "unknown target ref"); break; }
    d.ref = s.ref; d.n = s.n; d.cov = s.cov; d.synth_off = off;
    off += s.n;
    if ((rc = upload(e, s.x, s.n, &d.x)) != BH_OK) break;
    if ((rc = upload(e, s.y, s.n, &d.y)) != BH_OK) break;
    if (s.cov == BH_COV_WHITE_SCALED) {
      if (!s.yerr) { rc = set_err(BH_ERR_ARG, "BH_COV_WHITE_SCALED needs yerr"); break; }
      // scaled_err = yerr / yerr.min(); log(prod(scaled_err))  (Targets.py:125-128)
      std::vector<double> se(s.n);
      double mn = s.yerr[0];
      for (int i = 1; i < s.n; ++i) if (s.yerr[i] < mn) mn = s.yerr[i];
      double prod = 1.0;
      for (int i = 0; i < s.n; ++i) { se[i] = s.yerr[i] / mn; prod *= se[i]; }
      d.log_serr_prod = log(prod);
      if ((rc = upload(e, se.data(), s.n, &d.serr)) != BH_OK) break;
    } else if (s.cov == BH_COV_GAUSS) {
      if (!s.corr_inv) { rc = set_err(BH_ERR_ARG, "BH_COV_GAUSS needs corr_inv"); break; }
      // d^T A d = d^T (A + A^T)/2 d: the kernel walks the upper triangle of the symmetric part
      std::vector<double> sym((size_t)s.n * s.n);
      for (int i = 0; i < s.n; ++i)
        for (int j = 0; j < s.n; ++j)
          sym[(size_t)i * s.n + j] = 0.5 * (s.corr_inv[(size_t)i * s.n + j] + s.corr_inv[(size_t)j * s.n + i]);
      if ((rc = upload(e, sym.data(), (size_t)s.n * s.n, &d.corr_inv)) != BH_OK) break;
      d.logcorr_det = s.logcorr_det;
    } else if (s.cov != BH_COV_EXP && s.cov != BH_COV_WHITE) {
      rc = set_err(BH_ERR_ARG, "unknown covariance law"); break;
    }
    e->curve_off[t] = coff;
    if (is_swd(s.ref)) {
      if (s.mode < 1) { rc = set_err(BH_ERR_ARG, "mode must be >= 1 (1 = fundamental)"); break; }
      if (s.flsph != 0 && s.flsph != 1) { rc = set_err(BH_ERR_ARG, "flsph must be 0 (flat) or 1 (spherical)"); break; }
      d.mode = s.mode; d.flsph = s.flsph;
      d.wave = (s.ref == BH_REF_RDISPPH || s.ref == BH_REF_RDISPGR) ? 2 : 1;   // surf96_modsw.py:48-59
      d.igr = (s.ref == BH_REF_RDISPGR || s.ref == BH_REF_LDISPGR) ? 1 : 0;
      if (s.n > BH_MAX_PERIODS) {
        // surf96_modsw.py:35-43: 60-point linspace over [min, max], np.interp back
        double mn = s.x[0], mx = s.x[0];
        for (int i = 1; i < s.n; ++i) { if (s.x[i] < mn) mn = s.x[i]; if (s.x[i] > mx) mx = s.x[i]; }
        std::vector<double> pi(BH_MAX_PERIODS);
        const double step = (mx - mn) / (BH_MAX_PERIODS - 1);
        for (int i = 0; i < BH_MAX_PERIODS; ++i) pi[i] = mn + i * step;   // numpy.linspace
        pi[BH_MAX_PERIODS - 1] = mx;
        d.kmax = BH_MAX_PERIODS;
        if ((rc = upload(e, pi.data(), BH_MAX_PERIODS, &d.periods)) != BH_OK) break;
      } else {
        d.kmax = s.n;
        d.periods = d.x;
      }
      coff += d.kmax;
    } else if (s.ref == BH_REF_GENERIC) {
      e->has_generic = true;
    } else {
      // rfmini_modrf.py:41-62: fsamp, tshft, nsamp from the observed time axis
      if (s.n < 2) { rc = set_err(BH_ERR_ARG, "RF target needs >= 2 samples"); break; }
      double dt = round((s.x[1] - s.x[0]) * 1e4) / 1e4;
      for (int i = 2; i < s.n; ++i) {
        double di = round((s.x[i] - s.x[i - 1]) * 1e4) / 1e4;
        if (di != dt) { rc = set_err(BH_ERR_ARG, "RF sampling rate must be constant"); break; }
      }
      if (rc != BH_OK) break;
      d.fsamp = 1.0 / dt;
      d.tshift = -s.x[0];
      int ns = 1;
      while (ns < 2 * s.n) ns <<= 1;        // 2**ceil(log2(2*ndata))
      d.nsamp = ns;
      d.waveno = (s.ref == BH_REF_SRF) ? 1 : 0;
      d.gauss = s.gauss; d.p = s.p; d.nsv = s.nsv;
      d.qp = s.qp > 0 ? s.qp : 500.0;      // rfmini_modrf.py:119-120
      d.qs = s.qs > 0 ? s.qs : 225.0;
      if (ns / 2 + 1 > e->max_nfreq) e->max_nfreq = ns / 2 + 1;
      ++nrf;
    }
  }
  e->ts.synth_stride = off;
  e->curve_stride = coff > 0 ? coff : 1;
  const size_t B = (size_t)max_batch, L = (size_t)max_layers;
  e->prep.swd_stride = odd_stride(max_layers);
  if (rc == BH_OK) rc = scratch(e, &e->prep.swd_rows, B * e->prep.swd_stride);
  if (rc == BH_OK && nrf) rc = scratch(e, &e->prep.rf_lay, B * L);
  if (rc == BH_OK && nrf) rc = scratch(e, &e->prep.rf_coef, B * L * 4);
  if (rc == BH_OK && nrf) rc = scratch(e, &e->prep.rf_mc, B * 16);
  if (rc == BH_OK && nrf) rc = scratch(e, &e->spec, B * e->max_nfreq);
  if (rc == BH_OK) rc = scratch(e, &e->curves, B * e->curve_stride);
  if (rc == BH_OK) rc = scratch(e, &e->roots, 2 * B * e->curve_stride);
  if (rc == BH_OK) rc = scratch(e, &e->rfsynth, B * (size_t)off);
  if (rc == BH_OK) rc = scratch(e, &e->tstatus, B * kMaxTargets);
  if (rc == BH_OK) rc = scratch(e, &e->gauss_phi, B * kMaxTargets);
  {
    int gn = 0;
    for (int t = 0; t < ntargets; ++t)
      if (targets[t].cov == BH_COV_GAUSS && targets[t].n > gn) gn = targets[t].n;
    if (rc == BH_OK && gn > 0) rc = scratch(e, &e->gauss_res, B * 32 * (size_t)gauss_tile_rows(gn));
    if (rc == BH_OK && gn > 0) rc = scratch(e, &e->gauss_part, B * (size_t)gauss_tile_rows(gn));
  }
  if (rc == BH_OK) rc = scratch(e, &e->counters, BH_NUM_COUNTERS);
  if (rc == BH_OK) rc = scratch(e, &e->swd_queue, 2 + 1024);
  if (rc == BH_OK) rc = scratch(e, &e->swd_perm, B);
  if (rc == BH_OK) rc = scratch(e, &e->d_maxn, 1);
  if (rc == BH_OK) rc = scratch(e, &e->swd_done, 1);
  if (rc == BH_OK && (cudaMallocHost((void**)&e->h_maxn, sizeof(int)) != cudaSuccess ||
                      cudaEventCreateWithFlags(&e->ev_maxn, cudaEventDisableTiming) != cudaSuccess))
    rc = set_err(BH_ERR_CUDA, "pinned readback buffer");
  if (rc == BH_OK) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&e->nsm, cudaDevAttrMultiProcessorCount, dev);
  }
  if (rc == BH_OK) {
    cudaError_t ce;
    if ((ce = cudaStreamCreateWithFlags(&e->s_own, cudaStreamNonBlocking)) != cudaSuccess ||
        (ce = cudaStreamCreateWithFlags(&e->s_aux, cudaStreamNonBlocking)) != cudaSuccess ||
        (ce = cudaStreamCreateWithFlags(&e->s_aux2, cudaStreamNonBlocking)) != cudaSuccess ||
        (ce = cudaStreamCreateWithFlags(&e->s_aux3, cudaStreamNonBlocking)) != cudaSuccess ||
        (ce = cudaEventCreateWithFlags(&e->ev_join3, cudaEventDisableTiming)) != cudaSuccess ||
        (ce = cudaEventCreateWithFlags(&e->ev_fork3, cudaEventDisableTiming)) != cudaSuccess ||
        (ce = cudaEventCreateWithFlags(&e->ev_join2, cudaEventDisableTiming)) != cudaSuccess ||
        (ce = cudaEventCreateWithFlags(&e->ev_fork2, cudaEventDisableTiming)) != cudaSuccess ||
        (ce = cudaEventCreateWithFlags(&e->ev_fork, cudaEventDisableTiming)) != cudaSuccess ||
        (ce = cudaEventCreateWithFlags(&e->ev_join, cudaEventDisableTiming)) != cudaSuccess)
      rc = set_err(BH_ERR_CUDA, "stream/event creation", ce);
  }
  if (rc == BH_OK && cudaMemset(e->counters, 0, BH_NUM_COUNTERS * sizeof(unsigned long long)) != cudaSuccess)
    rc = set_err(BH_ERR_CUDA, "cudaMemset(counters)");
  if (rc != BH_OK) { std::string keep = g_err; bh_engine_destroy(e); g_err = keep; return rc; }
  *out = e;
  return BH_OK;
}

int bh_engine_synth_stride(const bh_engine* e) { return e ? e->ts.synth_stride : BH_ERR_ARG; }

int bh_engine_is_tuning(const bh_engine* e) { return (e && e->autotune && e->tune.locked < 0) ? 1 : 0; }

// CUDA-graph capture of evaluations: between _begin and _end an evaluation polls no event, reads nothing back and keeps
// its launch layout (the models-per-warp pick and the record capacity of the last plain evaluation; deeper models are
// still caught by the second launch), so that the caller may capture it with cudaStreamBeginCapture on its stream.
int bh_engine_capture_begin(bh_engine* e) {
  if (!e) return set_err(BH_ERR_ARG, "null engine");
  if (e->profile) return set_err(BH_ERR_UNSUPPORTED, "profiling records event pairs that cannot be read from a graph: set profile = 0");
  if (bh_engine_is_tuning(e)) return set_err(BH_ERR_UNSUPPORTED, "the engine is still timing layouts for this batch size (bh_engine_is_tuning)");
  e->graph_safe = 1;
  return BH_OK;
}
int bh_engine_capture_end(bh_engine* e) {
  if (!e) return set_err(BH_ERR_ARG, "null engine");
  e->graph_safe = 0;
  return BH_OK;
}

int bh_engine_set(bh_engine* e, const char* key, int value) {
  if (!e || !key) return set_err(BH_ERR_ARG, "null engine/key");
  if (!strcmp(key, "swd_searches_per_warp")) {
    if (value < 0 || value > 32 || (value & (value - 1))) return set_err(BH_ERR_ARG, "searches_per_warp must be 0 (auto) or a power of two <= 32");
    e->searches_per_warp = value;
  } else if (!strcmp(key, "swd_group_searches_per_warp")) {
    if (value < 0 || value > 32 || (value & (value - 1))) return set_err(BH_ERR_ARG, "group searches_per_warp must be 0 (auto) or a power of two <= 32");
    e->group_spw = value;
  } else if (!strncmp(key, "swd_spw_", 8) && strlen(key) == 10) {
    // swd_spw_rg / swd_spw_rp / swd_spw_lg / swd_spw_lp: models per warp of one curve type (0 = rule above)
    const int w = key[8] == 'r' ? 0 : key[8] == 'l' ? 2 : -1, g = key[9] == 'g' ? 0 : key[9] == 'p' ? 1 : -1;
    if (w < 0 || g < 0) return set_err(BH_ERR_ARG, "unknown tunable");
    if (value < 0 || value > (g == 0 ? 16 : 32)) return set_err(BH_ERR_ARG, "models per warp must be 0..32 (<= 16 for group curves)");
    e->spw_curve[w + g] = value;
  } else if (!strcmp(key, "swd_max_spec")) {
    if (value < 1 || value > 32) return set_err(BH_ERR_ARG, "swd_max_spec must be 1..32");
    e->max_spec = value;
  } else if (!strcmp(key, "swd_rayleigh_sm_pct")) {
    if (value < 0 || value > 100) return set_err(BH_ERR_ARG, "swd_rayleigh_sm_pct must be 0..100");
    e->rayleigh_sm_pct = value;
  } else if (!strcmp(key, "rf_prune_exp10")) {
    // bins with a Gauss weight below 10^-value are skipped; 0 disables the pruning
    if (value < 0 || value > 300) return set_err(BH_ERR_ARG, "rf_prune_exp10 must be 0 (off) or 1..300");
    e->rf_floor = value ? pow(10.0, -(double)value) : 0.0;
  } else if (!strcmp(key, "rf_gate_pct")) {
    if (value < 0 || value > 100) return set_err(BH_ERR_ARG, "rf_gate_pct must be 0..100");
    e->rf_gate_pct = value;
  } else if (!strcmp(key, "swd_autotune")) {
    e->autotune = value ? 1 : 0;
    e->tune = bh_engine::Tune();
    if (e->autotune && !e->ev_tune[0]) {
      BH_CUDA(cudaEventCreate(&e->ev_tune[0]));
      BH_CUDA(cudaEventCreate(&e->ev_tune[1]));
    }
  } else if (!strcmp(key, "swd_adaptive_capacity")) {
    e->adaptive_lcap = value ? 1 : 0;
    e->last_maxn = -1;
  } else if (!strcmp(key, "swd_sort_layers")) {
    e->sort_layers = value ? 1 : 0;
  } else if (!strcmp(key, "swd_split_waves")) {
    e->split_waves = value ? 1 : 0;
  } else if (!strcmp(key, "swd_direct")) {
    if (value < 0 || value > 2) return set_err(BH_ERR_ARG, "swd_direct must be 0, 1 or 2");
    e->direct = value;
  } else if (!strcmp(key, "swd_lockstep")) {
    if (value < 0 || value > 1) return set_err(BH_ERR_ARG, "swd_lockstep must be 0 or 1");
    e->lockstep = value;
  } else if (!strcmp(key, "swd_pool")) {
    if (value < -1 || value > 1) return set_err(BH_ERR_ARG, "swd_pool must be -1 (rule), 0 or 1");
    e->pool = value;
  } else if (!strcmp(key, "swd_pool_models")) {
    if (value < 0 || value > 128) return set_err(BH_ERR_ARG, "swd_pool_models must be in [0, 128]");
    e->pool_models = value;
  } else if (!strcmp(key, "swd_ls_spw_group") || !strcmp(key, "swd_ls_spw_phase")) {
    const int g = key[11] == 'g' ? 0 : 1;
    if (value < 0 || value > (g == 0 ? 16 : 32)) return set_err(BH_ERR_ARG, "models per warp must be 0..32 (<= 16 for group curves)");
    e->ls_spw[g] = value;
  } else if (!strcmp(key, "concurrent")) {
    e->concurrent = value ? 1 : 0;

  } else if (!strcmp(key, "profile")) {
    e->profile = value ? 1 : 0;
    if (e->profile && !e->pev[0])
      for (int i = 0; i < 2 * BH_NUM_KERNELS; ++i) BH_CUDA(cudaEventCreate(&e->pev[i]));
  } else {
    return set_err(BH_ERR_ARG, "unknown tunable");
  }
  return BH_OK;
}

int bh_engine_eval(bh_engine* e, const double* model, const int* nlay, const double* noise,
                   const double* rho, int B, int lmax, double* logL, double* misfits, int* status,
                   double* synth, void* stream) {
  if (!e || !model || !nlay || !noise || !logL || !misfits || !status)
    return set_err(BH_ERR_ARG, "null argument");
  if (B < 0 || B > e->max_batch) return set_err(BH_ERR_ARG, "B exceeds max_batch");
  if (lmax < 1 || lmax > e->max_layers) return set_err(BH_ERR_ARG, "lmax exceeds max_layers");
  if (B == 0) return BH_OK;
  if (e->has_generic)
    return set_err(BH_ERR_UNSUPPORTED, "a BH_REF_GENERIC target has no forward model here: use bh_engine_loglik_host");
  cudaStream_t st = (cudaStream_t)stream;
  const TargetSet& ts = e->ts;

  // one launch per wave type (Rayleigh, Love); inside a launch the curves with the
  // longer serial chains come first (group before phase)
  SwdLaunch swl[2] = {};
  int first_rf = -1;
  for (int pass = 0; pass < 4; ++pass) {
    const int want_wave = pass < 2 ? 2 : 1, want_igr = (pass & 1) ? 0 : 1;
    SwdLaunch& sw = swl[pass < 2 ? 0 : 1];
    for (int t = 0; t < ts.ntargets; ++t) {
      const TargetDev& d = ts.t[t];
      if (!is_swd(d.ref) || d.wave != want_wave || d.igr != want_igr) continue;
      if (d.mode != 1 || d.flsph != 0) continue;       // non-default branches: general kernel below
      int c = sw.ncurves++;
      sw.target_id[c] = t; sw.wave[c] = d.wave; sw.igr[c] = d.igr; sw.kmax[c] = d.kmax;
      sw.periods[c] = d.periods; sw.curve_off[c] = e->curve_off[t]; sw.synth_off[c] = d.synth_off;
    }
  }
  const int nswd = swl[0].ncurves + swl[1].ncurves;
  SwdGeneralLaunch gen{};
  for (int t = 0; t < ts.ntargets; ++t) {
    const TargetDev& d = ts.t[t];
    if (!is_swd(d.ref)) continue;
    // the packed rows carry vp as vs * vp/vs, so a water layer cannot occur in a batch
    // (only bh_surfdisp96 sees one); default settings take the fast kernel
    if (d.mode == 1 && d.flsph == 0) continue;
    int c = gen.ncurves++;
    gen.target_id[c] = t; gen.wave[c] = d.wave; gen.igr[c] = d.igr; gen.kmax[c] = d.kmax;
    gen.mode[c] = d.mode; gen.flsph[c] = d.flsph;
    gen.periods[c] = d.periods; gen.curve_off[c] = e->curve_off[t];
  }
  // The pool kernel (swd_pool.cu) against swd_kernel, joint5 unless noted (profiles/r02_swd_restructure.txt sections 12,
  // 15, 16, 19): B = 2560 2.18 -> 1.94 ms, 4096 2.55 -> 2.36, 8192 3.64 -> 3.36, 16384 7.14 -> 6.47; slower at B <= 2048.
  // With one wave type: deep models (transd3) from ~1800 models, 2048: 5.12 -> 4.54, 4096: 6.34 -> 5.58, 8192: 9.49 -> 7.61;
  // shallow models (swd2) only between ~4 k and ~8 k models at 14 per CTA (4096: 1.47 -> 1.42, 8192: 2.08 -> 1.97;
  // 2048 and 16384: swd_kernel is faster).
  int pool_cpc = 0;      // chains per CTA of the pool kernel; 0: swd_kernel
  if (e->pool != 0 && nswd > 0) {
    const int nl = (swl[0].ncurves > 0) + (swl[1].ncurves > 0);
    const bool deep = lmax > 12;
    bool fits = (nl == 1 || e->concurrent) && !e->lockstep;
    for (int w = 0; w < 2; ++w) fits = fits && (swl[w].ncurves == 0 || swd_pool_fits(swl[w]));
    // counted in chains (3 per model with a group and a phase curve): as many per CTA as fill four CTAs per SM, within
    // 26..96 with both wave types, 24..96 for deep models with one, 42 for shallow models with one
    long long chains = 0;
    int cpm_max = 1;
    for (int w = 0; w < 2; ++w) {
      int cpm = 0;
      for (int c = 0; c < swl[w].ncurves; ++c) cpm += swl[w].igr[c] ? 2 : 1;
      chains += (long long)B * cpm;
      if (cpm > cpm_max) cpm_max = cpm;
    }
    const long long slots = 4LL * e->nsm;
    const int clo = nl == 2 ? 26 : (deep ? 24 : 42);
    int cpc = slots > 0 ? (int)((chains + slots - 1) / slots) : 84;
    cpc = cpc < clo ? clo : (cpc > 96 ? 96 : cpc);
    bool pays = false;
    if (e->nsm > 0 && e->searches_per_warp == 0) {
      if (nl == 2) pays = chains * 100 >= slots * clo * 95;
      else if (deep) pays = chains * 100 >= slots * clo * 38;           // ~5.4 k chains on 148 SMs
      else { pays = chains * 100 >= slots * 42 * 46 && chains <= slots * 44; cpc = 42; }
      pays = pays && swd_pool_smem_bytes(lmax, (cpc + cpm_max - 1) / cpm_max) * 4 <= (size_t)220 * 1024;
    }
    if (fits && (e->pool == 1 || pays)) pool_cpc = cpc;
  }
  if (!e->split_waves && pool_cpc == 0 && swl[0].ncurves > 0 && swl[1].ncurves > 0) {
    // one mixed launch: append the Love curves to the Rayleigh launch
    SwdLaunch& a = swl[0];
    const SwdLaunch& b = swl[1];
    for (int c = 0; c < b.ncurves; ++c) {
      int k = a.ncurves++;
      a.target_id[k] = b.target_id[c]; a.wave[k] = b.wave[c]; a.igr[k] = b.igr[c]; a.kmax[k] = b.kmax[c];
      a.periods[k] = b.periods[c]; a.curve_off[k] = b.curve_off[c]; a.synth_off[k] = b.synth_off[c];
    }
    swl[1].ncurves = 0;
  }
  swl[1].counter_base = swl[0].ncurves;
  for (int t = 0; t < ts.ntargets; ++t)
    if (is_rf(ts.t[t].ref)) { first_rf = t; break; }
  const bool have_rf = first_rf >= 0;
  PrepOut prep = e->prep;
  prep.swd_stride = odd_stride(lmax);   // rows of this batch; buffer is sized for max_layers
  BH_CUDA(cudaMemsetAsync(e->counters, 0, BH_NUM_COUNTERS * sizeof(unsigned long long), st));
  const bool gated = e->rf_gate_pct > 0 && have_rf && nswd > 0 && e->concurrent;
  if (gated) BH_CUDA(cudaMemsetAsync(e->swd_done, 0, sizeof(int), st));

  cudaStream_t st_rf = st;
  if (have_rf && nswd > 0 && e->concurrent) {
    BH_CUDA(cudaEventRecord(e->ev_fork, st));
    BH_CUDA(cudaStreamWaitEvent(e->s_aux, e->ev_fork, 0));
    st_rf = e->s_aux;
  }

  for (bool& u : e->pev_used) u = false;
  // receiver-function launches of this evaluation
  auto launch_rf = [&](cudaStream_t st_rf) {
  for (int t = 0; t < ts.ntargets; ++t) {
    const TargetDev& d = ts.t[t];
    if (!is_rf(d.ref)) continue;
    { KTimer kt(e, BH_K_PREP_RF, st_rf);
      launch_prepare(model, nlay, rho, B, lmax, false, true, d.p, d.nsv, d.qp, d.qs, prep, st_rf); }
    RfLaunch rf{};
    rf.lay = prep.rf_lay; rf.coef = prep.rf_coef; rf.mc = prep.rf_mc; rf.nlay = nlay;
    rf.B = B; rf.lmax = lmax;
    rf.k.dw = 2.0 * RF_PI * d.fsamp / d.nsamp;
    rf.k.wref = 2.0 * RF_PI * 1.0;
    rf.k.a = d.gauss; rf.k.tshift = d.tshift;
    rf.k.qn = sqrt(RF_PI) * d.fsamp / d.gauss;
    rf.k.u = d.p * RF_DEG_PER_KM;
    rf.k.waveno = d.waveno; rf.k.nsamp = d.nsamp;
    rf.spec = e->spec;
    rf.nact = rf_active_frequencies(rf.k, e->rf_floor);
    rf.out = e->rfsynth; rf.out_stride = ts.synth_stride; rf.out_off = d.synth_off; rf.ndata = d.n;
    rf.tstatus = e->tstatus; rf.target_id = t;
    { KTimer kt(e, BH_K_RF_SPECTRUM, st_rf); launch_rf_spectrum(rf, st_rf); }
    { KTimer kt(e, BH_K_RF_SYNTH, st_rf); launch_rf_synth(rf, st_rf); }
  }
  };
  bool love_forked = false;
  int gate_warps = 0;
  if (nswd > 0 || gen.ncurves > 0) {
    // record capacity of the main dispersion launch: the layer counts seen lately (+2), not lmax
    int cap = lmax;
    if (e->adaptive_lcap && e->sort_layers && nswd > 0) {
      if (!e->graph_safe && e->maxn_pending && cudaEventQuery(e->ev_maxn) == cudaSuccess) {
        e->last_maxn = *e->h_maxn;
        e->maxn_pending = false;
      }
      if (e->last_maxn > 0 && e->last_maxn + 2 < lmax) cap = e->last_maxn + 2;
    }
    { KTimer kt(e, BH_K_PREP_SWD, st);
      launch_prepare(model, nlay, rho, B, lmax, true, false, 0, 0, 0, 0, prep, st);
      if (e->sort_layers && nswd > 0) {
        launch_layer_order(nlay, B, e->swd_perm, e->d_maxn, st);
      } }
    // models per warp: phase curves S (one chain per model), group curves S_g <= 16
    // (two chains per model).  Fewer models per warp = more spare lanes for bracket
    // speculation = fewer rounds, but more warps to issue.  Rule fitted to sweeps on B200
    // (profiles/r01_swd_sweep.txt, r01_variants.txt): about 14 chains per warp until one resident
    // wave (~2400 warps) is full.  With "swd_autotune" the neighbours of that pick are timed on the
    // first evaluations (events, never waited for) and the fastest is kept: results do not depend
    // on the choice, only the time does.
    int S = e->searches_per_warp, Sg = e->group_spw;
    bool tuning_now = false;
    if (S == 0) {
      long long nph = 0, ngr = 0;
      for (int w = 0; w < 2; ++w)
        for (int c = 0; c < swl[w].ncurves; ++c) (swl[w].igr[c] ? ngr : nph) += 1;
      const long long chains = (long long)B * (nph + 2 * ngr);
      long long target = chains / 14;
      if (target > 2400) target = 2400;
      // below a full machine, lanes beyond the chains' own pay twice -- bracket steps and refinement guesses
      // (swd_core.cuh: deal_lanes): about 4 chains per warp up to ~11 warps per SM (profiles/r02_swd_restructure.txt
      // section 14: joint5 B = 1024 1.75 -> 1.38 ms, B = 2048 2.03 -> 1.76, swd2 B = 4096 1.68 -> 1.48)
      {
        long long t2 = chains / 4;
        // (deep models -- long evaluations, relatively cheap bookkeeping -- take lanes up to 14 warps per SM: transd3
        //  B = 4096 6.79 -> 6.46 ms)
        const long long cap2 = e->nsm > 0 ? (lmax > 12 ? 14LL : 11LL) * e->nsm + e->nsm / 8 : 1650;
        if (t2 > cap2) t2 = cap2;
        if (target < t2) target = t2;
        // and with the guess trees up to 30 lanes deep, 1.5 chains per warp up to 7.5 warps per SM
        // (section 17: joint5 B = 256 1.04 -> 0.89 ms, swd2 B = 1024 0.95 -> 0.85)
        long long t3 = chains * 2 / 3;
        const long long cap3 = e->nsm > 0 ? 7LL * e->nsm + e->nsm / 2 : 1100;
        if (t3 > cap3) t3 = cap3;
        if (target < t3) target = t3;
      }
      // small batches: the chains are the critical path and the machine is far from full -- one warp per SM
      // sub-partition, down to one chain per warp (its idle lanes walk 16 steps a round and carry the refinement
      // guesses), before chains are packed
      // (profiles/r01_variants.txt: joint5 B = 256 / 512 / 1024: 1.81 -> 1.49 / 1.58 / 1.72 ms; r02 section 17:
      //  one chain instead of two per warp, 512-chain tutorial ensemble 1.11 -> 1.04 ms per iteration)
      if (e->nsm > 0) {
        long long fill = 4LL * e->nsm;
        if (fill > chains) fill = chains;
        if (target < fill) target = fill;
      }
      static const int cand[][2] = {{32, 16}, {16, 16}, {16, 8}, {8, 8}, {8, 4}, {4, 4}, {4, 2}, {2, 2}, {2, 1}, {1, 1}};
      int pick = 9;
      long long bestd = -1;
      for (int i = 0; i < 10; ++i) {
        const long long warps = nph * ((B + cand[i][0] - 1) / cand[i][0]) + ngr * ((B + cand[i][1] - 1) / cand[i][1]);
        const long long d = warps > target ? warps - target : target - warps;
        if (bestd < 0 || d < bestd) { bestd = d; pick = i; }
      }
      if (e->nsm > 0) {
        // a pick that overflows even the dense build's resident wave (16 warps per SM) runs in two waves:
        // prefer the largest candidate that, with its phase curves lifted to <= 24 models per warp, fits
        // under the roomy build's line (B = 7168, four curves: 5.1 -> 4.1 ms)
        auto grid = [&](int i, int sph) {
          return nph * ((B + sph - 1) / sph) + ngr * ((B + cand[i][1] - 1) / cand[i][1]);
        };
        const long long line = 12LL * e->nsm - e->nsm / 4;
        if (grid(pick, cand[pick][0]) > 16LL * e->nsm) {
          long long bestw = -1;
          for (int i = 0; i < 10; ++i) {
            int sph = cand[i][0];
            while (sph < 24 && grid(i, sph) > line) ++sph;
            const long long w = grid(i, sph);
            if (w <= line && w > bestw) { bestw = w; pick = i; }
          }
        }
      }
      if (e->autotune && Sg == 0) {
        bh_engine::Tune& t = e->tune;
        if (t.B != B || t.base != pick) {              // new problem size: start over
          t = bh_engine::Tune();
          t.B = B; t.base = pick;
          for (int i = pick - 1; i <= pick + 1; ++i) if (i >= 0 && i < 10) t.cand[t.ncand++] = i;
        }
        if (t.locked < 0 && !e->graph_safe) {
          if (t.pending && cudaEventQuery(e->ev_tune[1]) == cudaSuccess) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, e->ev_tune[0], e->ev_tune[1]) == cudaSuccess) {
              if (t.trials[t.cur] == 0 || ms < t.best_ms[t.cur]) t.best_ms[t.cur] = ms;
              t.trials[t.cur] += 1;
            }
            t.pending = false;
            int next = -1;
            for (int k = 1; k <= t.ncand; ++k) {       // round robin over candidates short of 3 timings
              const int j = (t.cur + k) % t.ncand;
              if (t.trials[j] < 3) { next = j; break; }
            }
            if (next < 0) {
              int b = 0;
              for (int j = 1; j < t.ncand; ++j) if (t.best_ms[j] < t.best_ms[b]) b = j;
              t.locked = t.cand[b];
            } else {
              t.cur = next;
            }
          }
          if (t.locked < 0 && !t.pending) {            // time ONE evaluation with the candidate under test;
            pick = t.cand[t.cur];                      // while that measurement is in flight (the host may be
            tuning_now = true;                         // far ahead of the device) keep the rule's pick
          }
        }
        if (t.locked >= 0) pick = t.locked;
      }
      S = cand[pick][0];
      if (Sg == 0) Sg = cand[pick][1];
    }
    if (Sg == 0) Sg = S > 16 ? 16 : S;
    if (Sg > 16) Sg = 16;
    // A grid a little above 12 warps per SM would take the 128-register build of the kernel (~10 % slower
    // code).  Phase curves have spare lanes: a few more models per warp (up to 24, not a power of two) bring
    // the grid back under that line -- 16 -> 23 at B = 8192 with four curves: 4.6 -> 4.34 ms.
    if (e->searches_per_warp == 0 && e->nsm > 0) {
      long long nph = 0, wgr = 0;
      for (int w = 0; w < 2; ++w)
        for (int c = 0; c < swl[w].ncurves; ++c) {
          if (swl[w].igr[c]) wgr += (B + Sg - 1) / Sg; else nph += 1;
        }
      // a quarter of a warp per SM of slack: a grid that fills every slot leaves no room for the one-thread
      // gate kernel and runs in two waves every other time (4.3 / 5.2 ms bimodal at 1770 of 1776 slots,
      // stable 4.34 ms at 1738; profiles/r01_variants.txt)
      const long long line = 12LL * e->nsm - e->nsm / 4;
      const long long now = wgr + nph * ((B + S - 1) / S);
      if (nph > 0 && now > line && S < 24) {
        for (int s2 = S + 1; s2 <= 24; ++s2)
          if (wgr + nph * ((B + s2 - 1) / s2) <= line) { S = s2; break; }
      }
    }
    // keep one warp's records + mailbox within ~24 KB of shared memory
    auto fit = [](int s_, int lc) { while (s_ > 1 && swd_smem_bytes(lc, s_) > 24 * 1024) s_ >>= 1; return s_; };
    // The reduced capacity exists to keep many warps resident when a warp holds many models.  With few models per warp
    // the full-capacity records are small anyway, and a model deeper than the capacity would cost a whole second search
    // BEHIND the first one -- twice the latency of a launch whose chains are the critical path (512-chain tutorial
    // ensemble: 1.39 -> 1.09 ms per iteration with the full capacity).
    if (cap < lmax && pool_cpc == 0 && (size_t)SWD_REC_FIELDS * lmax * (S > Sg ? S : Sg) * sizeof(double) <= 6 * 1024) cap = lmax;
    if (swl[0].ncurves > 0 && swl[1].ncurves > 0 && e->concurrent) {
      // Love chains share the SMs with the Rayleigh chains: own stream, forked after
      // the row preparation and before the Rayleigh launch
      BH_CUDA(cudaEventRecord(e->ev_fork2, st));
      BH_CUDA(cudaStreamWaitEvent(e->s_aux2, e->ev_fork2, 0));
      love_forked = true;
    }
    // the second search (models deeper than the record capacity) does not depend on the first: own stream, so that
    // its few deep chains run beside the main launch instead of behind it (8192-chain tutorial ensemble: 3.09 -> ~2.2 ms)
    bool deep_forked = false;
    if (cap < lmax && e->concurrent) {
      BH_CUDA(cudaEventRecord(e->ev_fork3, st));
      BH_CUDA(cudaStreamWaitEvent(e->s_aux3, e->ev_fork3, 0));
      deep_forked = true;
    }
    if (tuning_now) BH_CUDA(cudaEventRecord(e->ev_tune[0], st));
    for (int w = 0; w < 2; ++w) {
      SwdLaunch& sw = swl[w];
      if (sw.ncurves == 0) continue;
      sw.rows = prep.swd_rows; sw.row_stride = prep.swd_stride; sw.nlay = nlay; sw.B = B;
      sw.perm = e->sort_layers ? e->swd_perm : nullptr;
      sw.curves = e->curves; sw.roots = e->roots; sw.curve_stride = e->curve_stride;
      sw.tstatus = e->tstatus; sw.counters = e->counters;
      sw.max_spec = e->max_spec;
      sw.done = gated ? e->swd_done : nullptr;
      cudaStream_t sst = st;
      if (w == 1 && love_forked) sst = e->s_aux2;
      bool mixed = false;
      for (int c = 1; c < sw.ncurves; ++c) mixed |= sw.wave[c] != sw.wave[0];
      // pass 0: models with at most `cap` rows, records sized for cap; pass 1 (only when cap < lmax):
      // the deeper models with full capacity.  Warps of pass 1 without such a model exit at once.
      const cudaStream_t sst0 = sst;
      for (int pass = 0; pass < (cap < lmax ? 2 : 1); ++pass) {
        const int lc = pass == 0 ? cap : lmax;
        sst = (pass == 1 && deep_forked) ? e->s_aux3 : sst0;
        sw.lcap = lc;
        sw.nlay_lo = pass == 0 ? -1 : cap;
        sw.nlay_hi = (pass == 0 && cap < lmax) ? cap : 0x7fffffff;
        for (int c = 0; c < sw.ncurves; ++c) {
          sw.spw[c] = fit(sw.igr[c] ? Sg : S, lc);
          const int ov = e->spw_curve[(sw.wave[c] == 2 ? 0 : 2) + (sw.igr[c] ? 0 : 1)];
          if (ov > 0) sw.spw[c] = fit(ov, lc);
        }
        bool full = true;
        for (int c = 0; c < sw.ncurves; ++c) full = full && sw.spw[c] == (sw.igr[c] ? 16 : 32);
        sw.direct = (e->direct == 2 || (e->direct == 1 && full)) ? 1 : 0;
        sw.lockstep = e->lockstep;       // every lane owns a chain (swd_lockstep.cu)
        if (sw.lockstep) {
          for (int c = 0; c < sw.ncurves; ++c) {
            int s_ = e->ls_spw[sw.igr[c] ? 0 : 1];
            if (s_ <= 0) s_ = sw.igr[c] ? 16 : 32;
            while (s_ > 1 && (size_t)SWD_REC_FIELDS * lc * s_ * sizeof(double) > 24 * 1024) s_ >>= 1;
            sw.spw[c] = s_;
          }
        }
        sw.queue = nullptr;
        if (pass == 0 && mixed && e->rayleigh_sm_pct > 0 && e->nsm > 1 && e->nsm <= 1024) {
          // dedicate SMs [0, split) to the Rayleigh items, the rest to the Love items
          long long wr = 0, wl = 0;
          for (int c = 0; c < sw.ncurves; ++c) (sw.wave[c] == 2 ? wr : wl) += (B + sw.spw[c] - 1) / sw.spw[c];
          int split = (e->nsm * e->rayleigh_sm_pct / 100) & ~1;      // whole TPCs
          if (split < 2) split = 2;
          if (split > e->nsm - 2) split = e->nsm - 2;
          sw.queue = e->swd_queue; sw.sm_split = split;
          sw.type_quota[0] = (int)((wr + split - 1) / split);
          sw.type_quota[1] = (int)((wl + (e->nsm - split) - 1) / (e->nsm - split));
          BH_CUDA(cudaMemsetAsync(e->swd_queue, 0, (2 + e->nsm) * sizeof(int), sst));
        }
        const bool pool = pool_cpc > 0;
        int pm = 0;
        if (pool) {
          int cpm = 0;
          for (int c = 0; c < sw.ncurves; ++c) cpm += sw.igr[c] ? 2 : 1;
          const int want = e->pool_models > 0 ? e->pool_models : (pool_cpc + cpm - 1) / cpm;
          pm = swd_pool_models(sw, lc, want);
        }
        if (pool) sw.queue = nullptr;
        gate_warps += pool ? swd_pool_warp_count(sw, pm) : swd_warp_count(sw);
        auto go = [&]() {
          if (pool) launch_swd_pool(sw, pm, sst);
          else if (sw.lockstep) launch_swd_lockstep(sw, sst);
          else launch_swd(sw, sst);
        };
        if (pass == 0) {
          KTimer kt(e, pool ? (w == 0 ? BH_K_SWD_POOL : BH_K_SWD_POOL_LOVE) : (w == 0 ? BH_K_SWD : BH_K_SWD_LOVE), sst);
          go();
        } else go();
      }
    }
    if (deep_forked) {
      BH_CUDA(cudaEventRecord(e->ev_join3, e->s_aux3));
      BH_CUDA(cudaStreamWaitEvent(st, e->ev_join3, 0));
    }
    if (tuning_now) {
      if (love_forked) {                               // the Love launch is part of what is timed
        BH_CUDA(cudaEventRecord(e->ev_join2, e->s_aux2));
        BH_CUDA(cudaStreamWaitEvent(st, e->ev_join2, 0));
      }
      BH_CUDA(cudaEventRecord(e->ev_tune[1], st));
      e->tune.pending = true;
    }
    if (e->adaptive_lcap && e->sort_layers && nswd > 0 && !e->graph_safe) {
      // read the batch's largest layer count back behind the dispersion kernel (for LATER evaluations;
      // enqueued after the launch so that it cannot delay it)
      BH_CUDA(cudaMemcpyAsync(e->h_maxn, e->d_maxn, sizeof(int), cudaMemcpyDeviceToHost, st));
      BH_CUDA(cudaEventRecord(e->ev_maxn, st));
      e->maxn_pending = true;
    }
  }
  LoglikLaunch ll{};
  ll.ts = ts;
  ll.curves = e->curves; ll.curve_stride = e->curve_stride;
  for (int t = 0; t < ts.ntargets; ++t) ll.curve_off[t] = e->curve_off[t];
  ll.rfsynth = e->rfsynth; ll.tstatus = e->tstatus; ll.noise = noise; ll.B = B;
  ll.logL = logL; ll.misfits = misfits; ll.status = status; ll.synth = synth; ll.gauss_phi = e->gauss_phi;
  ll.gauss_res = e->gauss_res; ll.gauss_part = e->gauss_part;
  // Gauss-law contractions: a single one on a receiver-function target runs right behind its traces on the
  // RF stream (in the dispersion kernel's shadow); otherwise they share one scratch and run before loglik
  int ngauss = 0, gauss_t = -1;
  for (int t = 0; t < ts.ntargets; ++t)
    if (ts.t[t].cov == BH_COV_GAUSS) { ++ngauss; gauss_t = t; }
  const bool gauss_on_rf = ngauss == 1 && is_rf(ts.t[gauss_t].ref);
  if (gated && st_rf != st && gate_warps > 0)
    launch_swd_gate(e->swd_done, (int)((long long)gate_warps * e->rf_gate_pct / 100), st_rf);
  launch_rf(st_rf);
  if (gauss_on_rf) { Range nv("bh:gauss_quadform"); launch_gauss_quadform(ll, gauss_t, st_rf); }
  if (st_rf != st) {
    BH_CUDA(cudaEventRecord(e->ev_join, st_rf));
    BH_CUDA(cudaStreamWaitEvent(st, e->ev_join, 0));
  }
  if (love_forked) {
    BH_CUDA(cudaEventRecord(e->ev_join2, e->s_aux2));
    BH_CUDA(cudaStreamWaitEvent(st, e->ev_join2, 0));
  }
  if (gen.ncurves > 0) {
    gen.rows = prep.swd_rows; gen.row_stride = prep.swd_stride; gen.lcap = lmax; gen.nlay = nlay; gen.B = B;
    gen.curves = e->curves; gen.curve_stride = e->curve_stride; gen.tstatus = e->tstatus;
    gen.counters = e->counters;
    KTimer kt(e, BH_K_SWD_GENERAL, st);
    launch_swd_general(gen, st);
  }
  { KTimer kt(e, BH_K_LOGLIK, st);
    if (!gauss_on_rf)
      for (int t = 0; t < ts.ntargets; ++t) launch_gauss_quadform(ll, t, st);
    launch_loglik(ll, st); }
  BH_CUDA(cudaGetLastError());
  return BH_OK;
}

int bh_engine_last_kernel_ms(bh_engine* e, float* ms) {
  if (!e || !ms) return set_err(BH_ERR_ARG, "null argument");
  if (!e->profile) return set_err(BH_ERR_ARG, "profiling is off: bh_engine_set(e, \"profile\", 1)");
  for (int k = 0; k < BH_NUM_KERNELS; ++k) {
    ms[k] = -1.0f;
    if (!e->pev_used[k]) continue;
    BH_CUDA(cudaEventSynchronize(e->pev[2 * k + 1]));
    BH_CUDA(cudaEventElapsedTime(&ms[k], e->pev[2 * k], e->pev[2 * k + 1]));
  }
  return BH_OK;
}

}  // extern "C"

namespace {

bool is_pinned(const void* p) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return at.type == cudaMemoryTypeHost;
}

template <class T>
int pinned_alloc(T** p, size_t n) {
  if (*p || n == 0) return BH_OK;
  BH_CUDA(cudaMallocHost((void**)p, n * sizeof(T)));
  return BH_OK;
}
template <class T>
int device_alloc_once(T** p, size_t n) {
  if (*p || n == 0) return BH_OK;
  BH_CUDA(cudaMalloc((void**)p, n * sizeof(T)));
  return BH_OK;
}

int slots_init(bh_engine* e) {
  if (e->slots_ready) return BH_OK;
  BH_CUDA(cudaStreamCreateWithFlags(&e->s_h2d, cudaStreamNonBlocking));
  BH_CUDA(cudaStreamCreateWithFlags(&e->s_d2h, cudaStreamNonBlocking));
  const size_t B = (size_t)e->max_batch, L = (size_t)e->max_layers, T = (size_t)e->ts.ntargets;
  for (auto& sl : e->slot) {
    int rc;
    if ((rc = device_alloc_once(&sl.d_model, B * L * 4)) != BH_OK) return rc;
    if ((rc = device_alloc_once(&sl.d_nlay, B)) != BH_OK) return rc;
    if ((rc = device_alloc_once(&sl.d_noise, B * 2 * T)) != BH_OK) return rc;
    if ((rc = device_alloc_once(&sl.d_logL, B)) != BH_OK) return rc;
    if ((rc = device_alloc_once(&sl.d_misfits, B * (T + 1))) != BH_OK) return rc;
    if ((rc = device_alloc_once(&sl.d_status, B)) != BH_OK) return rc;
    BH_CUDA(cudaEventCreateWithFlags(&sl.ev_in, cudaEventDisableTiming));
    BH_CUDA(cudaEventCreateWithFlags(&sl.ev_kernels, cudaEventDisableTiming));
    BH_CUDA(cudaEventCreateWithFlags(&sl.ev_done, cudaEventDisableTiming));
  }
  e->slots_ready = true;
  return BH_OK;
}

// Deliver a finished job: wait for its copies, move staged results to the caller's buffers.
int slot_finish(bh_engine* e, bh_engine::HostSlot& sl) {
  if (!sl.busy) return BH_OK;
  BH_CUDA(cudaEventSynchronize(sl.ev_done));
  if (sl.staged_out) {
    const size_t nb = (size_t)sl.B, T = (size_t)e->ts.ntargets;
    memcpy(sl.out_logL, sl.h_logL, nb * sizeof(double));
    memcpy(sl.out_misfits, sl.h_misfits, nb * (T + 1) * sizeof(double));
    memcpy(sl.out_status, sl.h_status, nb * sizeof(int));
    if (sl.out_synth) memcpy(sl.out_synth, sl.h_synth, nb * e->ts.synth_stride * sizeof(double));
  }
  sl.busy = false;
  return BH_OK;
}

}  // namespace

extern "C" {

int bh_engine_eval_host_async(bh_engine* e, const double* model, const int* nlay, const double* noise,
                              const double* rho, int B, int lmax, double* logL, double* misfits,
                              int* status, double* synth, long long* ticket) {
  if (!e || !model || !nlay || !noise || !logL || !misfits || !status || !ticket)
    return set_err(BH_ERR_ARG, "null argument");
  if (B < 0 || B > e->max_batch) return set_err(BH_ERR_ARG, "B exceeds max_batch");
  if (lmax < 1 || lmax > e->max_layers) return set_err(BH_ERR_ARG, "lmax exceeds max_layers");
  *ticket = 0;
  if (B == 0) return BH_OK;
  Range nv("bh:eval_host");
  int rc = slots_init(e);
  if (rc != BH_OK) return rc;
  const size_t T = (size_t)e->ts.ntargets, nb = (size_t)B, MB = (size_t)e->max_batch, ML = (size_t)e->max_layers;
  bh_engine::HostSlot& sl = e->slot[e->next_ticket & 1];
  if ((rc = slot_finish(e, sl)) != BH_OK) return rc;       // the job that used this slot two calls ago
  // ---- inputs: straight from pinned caller memory, else through this slot's pinned staging ----
  const double* src_model = model; const int* src_nlay = nlay; const double* src_noise = noise; const double* src_rho = rho;
  if (!is_pinned(model)) {
    if ((rc = pinned_alloc(&sl.h_model, MB * ML * 4)) != BH_OK) return rc;
    memcpy(sl.h_model, model, nb * lmax * 4 * sizeof(double)); src_model = sl.h_model;
  }
  if (!is_pinned(nlay)) {
    if ((rc = pinned_alloc(&sl.h_nlay, MB)) != BH_OK) return rc;
    memcpy(sl.h_nlay, nlay, nb * sizeof(int)); src_nlay = sl.h_nlay;
  }
  if (!is_pinned(noise)) {
    if ((rc = pinned_alloc(&sl.h_noise, MB * 2 * T)) != BH_OK) return rc;
    memcpy(sl.h_noise, noise, nb * 2 * T * sizeof(double)); src_noise = sl.h_noise;
  }
  if (rho) {
    if ((rc = device_alloc_once(&sl.d_rho, MB * ML)) != BH_OK) return rc;
    if (!is_pinned(rho)) {
      if ((rc = pinned_alloc(&sl.h_rho, MB * ML)) != BH_OK) return rc;
      memcpy(sl.h_rho, rho, nb * lmax * sizeof(double)); src_rho = sl.h_rho;
    }
  }
  if (synth && (rc = device_alloc_once(&sl.d_synth, MB * (size_t)e->ts.synth_stride)) != BH_OK) return rc;
  BH_CUDA(cudaMemcpyAsync(sl.d_model, src_model, nb * lmax * 4 * sizeof(double), cudaMemcpyHostToDevice, e->s_h2d));
  BH_CUDA(cudaMemcpyAsync(sl.d_nlay, src_nlay, nb * sizeof(int), cudaMemcpyHostToDevice, e->s_h2d));
  BH_CUDA(cudaMemcpyAsync(sl.d_noise, src_noise, nb * 2 * T * sizeof(double), cudaMemcpyHostToDevice, e->s_h2d));
  if (rho) BH_CUDA(cudaMemcpyAsync(sl.d_rho, src_rho, nb * lmax * sizeof(double), cudaMemcpyHostToDevice, e->s_h2d));
  BH_CUDA(cudaEventRecord(sl.ev_in, e->s_h2d));
  // ---- kernels (the engine's scratch is single: evaluations run one after the other on s_own) ----
  BH_CUDA(cudaStreamWaitEvent(e->s_own, sl.ev_in, 0));
  rc = bh_engine_eval(e, sl.d_model, sl.d_nlay, sl.d_noise, rho ? sl.d_rho : nullptr, B, lmax,
                      sl.d_logL, sl.d_misfits, sl.d_status, synth ? sl.d_synth : nullptr, e->s_own);
  if (rc != BH_OK) return rc;
  BH_CUDA(cudaEventRecord(sl.ev_kernels, e->s_own));
  // ---- results: straight into pinned caller memory, else staged and delivered by bh_engine_wait ----
  sl.staged_out = !(is_pinned(logL) && is_pinned(misfits) && is_pinned(status) && (!synth || is_pinned(synth)));
  double *dst_logL = logL, *dst_mis = misfits, *dst_synth = synth;
  int* dst_stat = status;
  if (sl.staged_out) {
    if ((rc = pinned_alloc(&sl.h_logL, MB)) != BH_OK) return rc;
    if ((rc = pinned_alloc(&sl.h_misfits, MB * (T + 1))) != BH_OK) return rc;
    if ((rc = pinned_alloc(&sl.h_status, MB)) != BH_OK) return rc;
    if (synth && (rc = pinned_alloc(&sl.h_synth, MB * (size_t)e->ts.synth_stride)) != BH_OK) return rc;
    dst_logL = sl.h_logL; dst_mis = sl.h_misfits; dst_stat = sl.h_status; dst_synth = synth ? sl.h_synth : nullptr;
  }
  BH_CUDA(cudaStreamWaitEvent(e->s_d2h, sl.ev_kernels, 0));
  BH_CUDA(cudaMemcpyAsync(dst_logL, sl.d_logL, nb * sizeof(double), cudaMemcpyDeviceToHost, e->s_d2h));
  BH_CUDA(cudaMemcpyAsync(dst_mis, sl.d_misfits, nb * (T + 1) * sizeof(double), cudaMemcpyDeviceToHost, e->s_d2h));
  BH_CUDA(cudaMemcpyAsync(dst_stat, sl.d_status, nb * sizeof(int), cudaMemcpyDeviceToHost, e->s_d2h));
  if (synth)
    BH_CUDA(cudaMemcpyAsync(dst_synth, sl.d_synth, nb * e->ts.synth_stride * sizeof(double), cudaMemcpyDeviceToHost, e->s_d2h));
  BH_CUDA(cudaEventRecord(sl.ev_done, e->s_d2h));
  sl.out_logL = logL; sl.out_misfits = misfits; sl.out_status = status; sl.out_synth = synth; sl.B = B;
  sl.busy = true;
  sl.ticket = e->next_ticket++;
  *ticket = sl.ticket;
  return BH_OK;
}

int bh_engine_wait(bh_engine* e, long long ticket) {
  if (!e) return set_err(BH_ERR_ARG, "null engine");
  if (ticket <= 0) return BH_OK;
  for (auto& sl : e->slot)
    if (sl.busy && sl.ticket == ticket) return slot_finish(e, sl);
  return BH_OK;                // already delivered (its slot was reused, which finishes the older job first)
}

int bh_engine_eval_host(bh_engine* e, const double* model, const int* nlay, const double* noise,
                        const double* rho, int B, int lmax, double* logL, double* misfits,
                        int* status, double* synth) {
  long long ticket = 0;
  int rc = bh_engine_eval_host_async(e, model, nlay, noise, rho, B, lmax, logL, misfits, status, synth, &ticket);
  if (rc != BH_OK) return rc;
  return bh_engine_wait(e, ticket);
}

int bh_engine_loglik_host(bh_engine* e, const double* synth, const int* tvalid, const double* noise, int B,
                          double* logL, double* misfits, int* status) {
  if (!e || !synth || !tvalid || !noise || !logL || !misfits || !status) return set_err(BH_ERR_ARG, "null argument");
  if (B < 0 || B > e->max_batch) return set_err(BH_ERR_ARG, "B exceeds max_batch");
  if (B == 0) return BH_OK;
  int rc = slots_init(e);
  if (rc != BH_OK) return rc;
  const TargetSet& ts = e->ts;
  const size_t T = (size_t)ts.ntargets, nb = (size_t)B, stride = (size_t)ts.synth_stride;
  bh_engine::HostSlot& sl = e->slot[0];
  for (auto& s2 : e->slot) if ((rc = slot_finish(e, s2)) != BH_OK) return rc;
  if ((rc = device_alloc_once(&sl.d_synth, (size_t)e->max_batch * stride)) != BH_OK) return rc;
  cudaStream_t st = e->s_own;
  std::vector<int> tv(nb * kMaxTargets, 1);
  for (size_t b = 0; b < nb; ++b)
    for (size_t t = 0; t < T; ++t) tv[b * kMaxTargets + t] = tvalid[b * T + t] ? 1 : 0;
  BH_CUDA(cudaMemcpyAsync(sl.d_synth, synth, nb * stride * sizeof(double), cudaMemcpyHostToDevice, st));
  BH_CUDA(cudaMemcpyAsync(e->tstatus, tv.data(), tv.size() * sizeof(int), cudaMemcpyHostToDevice, st));
  BH_CUDA(cudaMemcpyAsync(sl.d_noise, noise, nb * 2 * T * sizeof(double), cudaMemcpyHostToDevice, st));
  LoglikLaunch ll{};
  ll.ts = ts;
  ll.curves = e->curves; ll.curve_stride = e->curve_stride;
  for (int t = 0; t < ts.ntargets; ++t) ll.curve_off[t] = e->curve_off[t];
  ll.rfsynth = e->rfsynth; ll.given = sl.d_synth; ll.tstatus = e->tstatus; ll.noise = sl.d_noise; ll.B = B;
  ll.logL = sl.d_logL; ll.misfits = sl.d_misfits; ll.status = sl.d_status; ll.synth = nullptr;
  ll.gauss_phi = e->gauss_phi; ll.gauss_res = e->gauss_res; ll.gauss_part = e->gauss_part;
  for (int t = 0; t < ts.ntargets; ++t) launch_gauss_quadform(ll, t, st);
  launch_loglik(ll, st);
  BH_CUDA(cudaMemcpyAsync(logL, sl.d_logL, nb * sizeof(double), cudaMemcpyDeviceToHost, st));
  BH_CUDA(cudaMemcpyAsync(misfits, sl.d_misfits, nb * (T + 1) * sizeof(double), cudaMemcpyDeviceToHost, st));
  BH_CUDA(cudaMemcpyAsync(status, sl.d_status, nb * sizeof(int), cudaMemcpyDeviceToHost, st));
  BH_CUDA(cudaStreamSynchronize(st));
  BH_CUDA(cudaGetLastError());
  return BH_OK;
}

int bh_engine_last_counts(bh_engine* e, long long* nsec) {
  if (!e || !nsec) return set_err(BH_ERR_ARG, "null argument");
  unsigned long long h[BH_NUM_COUNTERS];
  BH_CUDA(cudaMemcpy(h, e->counters, sizeof(h), cudaMemcpyDeviceToHost));
  for (int i = 0; i < BH_NUM_COUNTERS; ++i) nsec[i] = (long long)h[i];
  return BH_OK;
}

// ---------------------------------------------------------------------------
// single-model shims (reference FFI semantics, host pointers)
// ---------------------------------------------------------------------------
namespace {
struct Shim {
  std::mutex mu;
  bool ready = false;
  LayerRow* rows = nullptr;
  double* periods = nullptr;
  double* curve = nullptr;
  double* roots = nullptr;
  int* nlay = nullptr;
  int* tstatus = nullptr;
  double* model6 = nullptr;      // z, vp, vs, rho, qp, qs  (6 x 100)
  RfLayer* rf_lay = nullptr;
  cm2* rf_coef = nullptr;
  double* rf_mc = nullptr;
  cd* spec = nullptr;
  double* trace = nullptr;
  int spec_cap = 0;
  cudaStream_t st = nullptr;
};
Shim g_shim;

int shim_init() {
  if (g_shim.ready) return BH_OK;
  if (bh_device_count() < 1)
    return set_err(BH_ERR_NO_DEVICE, "no CUDA device visible; this library has no CPU path");
  BH_CUDA(cudaMalloc((void**)&g_shim.rows, sizeof(LayerRow) * 101));
  BH_CUDA(cudaMalloc((void**)&g_shim.periods, sizeof(double) * BH_MAX_PERIODS));
  BH_CUDA(cudaMalloc((void**)&g_shim.curve, sizeof(double) * BH_MAX_PERIODS));
  BH_CUDA(cudaMalloc((void**)&g_shim.roots, sizeof(double) * 2 * BH_MAX_PERIODS));
  BH_CUDA(cudaMalloc((void**)&g_shim.nlay, sizeof(int)));
  BH_CUDA(cudaMalloc((void**)&g_shim.tstatus, sizeof(int) * kMaxTargets));
  BH_CUDA(cudaMalloc((void**)&g_shim.model6, sizeof(double) * 6 * BH_MAX_LAYERS));
  BH_CUDA(cudaMalloc((void**)&g_shim.rf_lay, sizeof(RfLayer) * BH_MAX_LAYERS));
  BH_CUDA(cudaMalloc((void**)&g_shim.rf_coef, sizeof(cm2) * 4 * BH_MAX_LAYERS));
  BH_CUDA(cudaMalloc((void**)&g_shim.rf_mc, sizeof(double) * 16));
  BH_CUDA(cudaStreamCreateWithFlags(&g_shim.st, cudaStreamNonBlocking));
  g_shim.ready = true;
  return BH_OK;
}
}  // namespace

int bh_surfdisp96(const float* thkm, const float* vpm, const float* vsm, const float* rhom,
                  int nlayer, int iflsph, int iwave, int mode, int igr, int kmax, const double* t,
                  double* cg, int* err) {
  if (!thkm || !vpm || !vsm || !rhom || !t || !cg || !err) return set_err(BH_ERR_ARG, "null argument");
  if (nlayer < 1 || nlayer > BH_MAX_LAYERS || kmax < 1 || kmax > BH_MAX_PERIODS)
    return set_err(BH_ERR_ARG, "nlayer must be 1..100 and kmax 1..60");
  if (iwave != 1 && iwave != 2) return set_err(BH_ERR_ARG, "iwave must be 1 (Love) or 2 (Rayleigh)");
  if (mode < 1) return set_err(BH_ERR_ARG, "mode must be >= 1");
  if (iflsph != 0 && iflsph != 1) return set_err(BH_ERR_ARG, "iflsph must be 0 or 1");
  std::lock_guard<std::mutex> lock(g_shim.mu);
  int rc = shim_init();
  if (rc != BH_OK) return rc;
  Shim& s = g_shim;
  std::vector<LayerRow> rows(nlayer);
  bool fluid = false;
  for (int i = 0; i < nlayer; ++i) {
    rows[i].x = thkm[i]; rows[i].y = vpm[i]; rows[i].z = vsm[i]; rows[i].w = rhom[i];
    fluid |= vsm[i] <= 0.01f;
  }
  BH_CUDA(cudaMemcpyAsync(s.rows, rows.data(), sizeof(LayerRow) * nlayer, cudaMemcpyHostToDevice, s.st));
  BH_CUDA(cudaMemcpyAsync(s.periods, t, sizeof(double) * kmax, cudaMemcpyHostToDevice, s.st));
  BH_CUDA(cudaMemcpyAsync(s.nlay, &nlayer, sizeof(int), cudaMemcpyHostToDevice, s.st));
  BH_CUDA(cudaMemsetAsync(s.curve, 0, sizeof(double) * BH_MAX_PERIODS, s.st));
  if (mode != 1 || iflsph != 0 || fluid) {
    SwdGeneralLaunch g{};
    g.rows = s.rows; g.row_stride = odd_stride(nlayer); g.lcap = nlayer; g.nlay = s.nlay; g.B = 1; g.ncurves = 1;
    g.target_id[0] = 0; g.wave[0] = iwave; g.igr[0] = igr > 0 ? 1 : 0; g.kmax[0] = kmax; g.mode[0] = mode;
    g.flsph[0] = iflsph; g.periods[0] = s.periods; g.curves = s.curve;
    g.curve_stride = BH_MAX_PERIODS; g.curve_off[0] = 0; g.tstatus = s.tstatus; g.counters = nullptr;
    launch_swd_general(g, s.st);
  } else {
    SwdLaunch sw{};
    sw.rows = s.rows; sw.row_stride = odd_stride(nlayer); sw.nlay = s.nlay; sw.B = 1; sw.ncurves = 1;
    sw.target_id[0] = 0; sw.wave[0] = iwave; sw.igr[0] = igr > 0 ? 1 : 0; sw.kmax[0] = kmax;
    sw.periods[0] = s.periods; sw.curves = s.curve; sw.roots = s.roots; sw.curve_stride = BH_MAX_PERIODS; sw.curve_off[0] = 0;
    sw.tstatus = s.tstatus; sw.counters = nullptr; sw.spw[0] = 1; sw.lcap = nlayer; sw.max_spec = 32; sw.direct = 0;
    sw.nlay_lo = -1; sw.nlay_hi = 0x7fffffff;
    launch_swd(sw, s.st);
  }
  int ok = 0;
  std::vector<double> out(kmax);
  BH_CUDA(cudaMemcpyAsync(out.data(), s.curve, sizeof(double) * kmax, cudaMemcpyDeviceToHost, s.st));
  BH_CUDA(cudaMemcpyAsync(&ok, s.tstatus, sizeof(int), cudaMemcpyDeviceToHost, s.st));
  BH_CUDA(cudaStreamSynchronize(s.st));
  BH_CUDA(cudaGetLastError());
  *err = ok ? 0 : 1;
  for (int k = 0; k < kmax; ++k) cg[k] = out[k];
  return BH_OK;
}

int bh_synrf(int nsamp, double fsamp, double tshift, double p, double a, double nsv, double sigma,
             int waveno, int nlay, const double* z, const double* vp, const double* vs,
             const double* rh, const double* qp, const double* qs, double* fz, double* fr,
             double* rf) {
  if (!z || !vp || !vs || !rh || !qp || !qs || !rf) return set_err(BH_ERR_ARG, "null argument");
  if (nlay < 2 || nlay > BH_MAX_LAYERS) return set_err(BH_ERR_ARG, "nlay must be 2..100");
  if (nsamp < 2 || (nsamp & (nsamp - 1))) return set_err(BH_ERR_ARG, "nsamp must be a power of two");
  if (waveno != 0 && waveno != 1) return set_err(BH_ERR_UNSUPPORTED, "waveno must be 0 (P) or 1 (SV)");
  std::lock_guard<std::mutex> lock(g_shim.mu);
  int rc = shim_init();
  if (rc != BH_OK) return rc;
  Shim& s = g_shim;
  const int nfreq = nsamp / 2 + 1;
  if (nfreq > s.spec_cap) {
    if (s.spec) { cudaFree(s.spec); cudaFree(s.trace); s.spec = nullptr; s.trace = nullptr; s.spec_cap = 0; }
    BH_CUDA(cudaMalloc((void**)&s.spec, sizeof(cd) * nfreq));
    BH_CUDA(cudaMalloc((void**)&s.trace, sizeof(double) * nsamp));
    s.spec_cap = nfreq;
  }
  const double* src[6] = {z, vp, vs, rh, qp, qs};
  for (int k = 0; k < 6; ++k)
    BH_CUDA(cudaMemcpyAsync(s.model6 + k * BH_MAX_LAYERS, src[k], sizeof(double) * nlay, cudaMemcpyHostToDevice, s.st));
  BH_CUDA(cudaMemcpyAsync(s.nlay, &nlay, sizeof(int), cudaMemcpyHostToDevice, s.st));
  PrepOut prep{};
  prep.rf_lay = s.rf_lay; prep.rf_coef = s.rf_coef; prep.rf_mc = s.rf_mc;
  double* m = s.model6;
  launch_prepare_rf_explicit(m, m + BH_MAX_LAYERS, m + 2 * BH_MAX_LAYERS, m + 3 * BH_MAX_LAYERS,
                             m + 4 * BH_MAX_LAYERS, m + 5 * BH_MAX_LAYERS, nlay, p, nsv, sigma, prep, s.st);
  RfLaunch rfl{};
  rfl.lay = s.rf_lay; rfl.coef = s.rf_coef; rfl.mc = s.rf_mc; rfl.nlay = s.nlay; rfl.B = 1; rfl.lmax = nlay;
  rfl.k.dw = 2.0 * RF_PI * fsamp / nsamp; rfl.k.wref = 2.0 * RF_PI; rfl.k.a = a; rfl.k.tshift = tshift;
  rfl.k.qn = sqrt(RF_PI) * fsamp / a; rfl.k.u = p * RF_DEG_PER_KM; rfl.k.waveno = waveno; rfl.k.nsamp = nsamp;
  rfl.spec = s.spec; rfl.nact = rf_active_frequencies(rfl.k, 1e-30); rfl.out = s.trace; rfl.out_stride = nsamp; rfl.out_off = 0; rfl.ndata = nsamp;
  rfl.tstatus = nullptr; rfl.target_id = 0;
  launch_rf_spectrum(rfl, s.st);
  launch_rf_synth(rfl, s.st);
  BH_CUDA(cudaMemcpyAsync(rf, s.trace, sizeof(double) * nsamp, cudaMemcpyDeviceToHost, s.st));
  BH_CUDA(cudaStreamSynchronize(s.st));
  BH_CUDA(cudaGetLastError());
  if (fz) memset(fz, 0, sizeof(double) * nsamp);
  if (fr) memset(fr, 0, sizeof(double) * nsamp);
  return BH_OK;
}

// ---------------------------------------------------------------------------
// Link-level drop-in symbols: the raw native entry points of the reference, so that its own glue code
// (the f2py module generated from surfdisp96.f, the Cython module rfmini.pyx) links against this library
// instead of surfdisp96.o / librfmini objects without a change.
// ---------------------------------------------------------------------------
// subroutine surfdisp96(thkm,vpm,vsm,rhom,nlayer,iflsph,iwave,mode,igr,kmax,t,cg,err)
// (src/extensions/surfdisp96.f:55-56; real*4 model arrays of NL = 100, double precision t(60), cg(60),
// integer scalars, err intent(out) :82-86,101): the gfortran symbol, every argument by reference.
void surfdisp96_(const float* thkm, const float* vpm, const float* vsm, const float* rhom, const int* nlayer,
                 const int* iflsph, const int* iwave, const int* mode, const int* igr, const int* kmax,
                 const double* t, double* cg, int* err) {
  int e = 1;
  const int rc = bh_surfdisp96(thkm, vpm, vsm, rhom, *nlayer, *iflsph, *iwave, *mode, *igr, *kmax, t, cg, &e);
  if (rc != BH_OK) {
    // no error channel besides err: report like "no root found" (the plugin returns (nan, nan)) and say why
    fprintf(stderr, "surfdisp96_ (libbayhunter_b200): %s\n", bh_last_error());
    e = 1;
    for (int k = 0; k < *kmax && k < BH_MAX_PERIODS; ++k) cg[k] = 0.0;
  }
  *err = e;
}

// extern "C" int synrf_cwrap(...) (src/extensions/rfmini/wrap.cpp:26-31, 57-80): always returns 1; fz / fr are
// the Z / R traces BayHunter discards (zero-filled here).  A failure (no device) has no channel in this
// prototype: the trace is filled with NaN, which BayHunter rejects (a NaN likelihood fails `u < alpha`).
int synrf_cwrap(int nsamp, double fsamp, double tshift, double p, double a, double nsv, double sigma, int waveno,
                int nlay, double* z, double* vp, double* vs, double* rh, double* qp, double* qs, double* fz,
                double* fr, double* rf) {
  const int rc = bh_synrf(nsamp, fsamp, tshift, p, a, nsv, sigma, waveno, nlay, z, vp, vs, rh, qp, qs, fz, fr, rf);
  if (rc != BH_OK) {
    fprintf(stderr, "synrf_cwrap (libbayhunter_b200): %s\n", bh_last_error());
    if (rf) for (int i = 0; i < nsamp; ++i) rf[i] = nan("");
  }
  return 1;
}

// ---------------------------------------------------------------------------
// diagnostics: the straight-line elementary functions of bh_math.cuh, evaluated
// on the device for a host vector (tests compare them with libm)
// ---------------------------------------------------------------------------
}  // extern "C"

namespace {
__global__ void debug_math_kernel(const double* __restrict__ x, double* __restrict__ out, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double v = x[i], s, c, sq, rsq;
  fm::sincos_cw(v, &s, &c);
  fm::sqrt_rsqrt(fabs(v), &sq, &rsq);
  out[i] = fm::exp_small(-fabs(v));
  out[n + i] = s;
  out[2 * n + i] = c;
  out[3 * n + i] = fm::rcp(v);
  out[4 * n + i] = sq;
  out[5 * n + i] = rsq;
  out[6 * n + i] = fm::div(1.0, v);
}
}  // namespace

namespace {
// 8 independent DFMA chains per thread: the fp64 pipe's issue rate, not its latency
__global__ void fp64_peak_kernel(double* out, double a, double b, int iters) {
  double x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = a + i + threadIdx.x;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
#pragma unroll
      for (int i = 0; i < 8; ++i) x[i] = fma(x[i], b, a);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i];
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}
}  // namespace

extern "C" int bh_measure_fp64_peak(double* tflops, double* sm_mhz_seen) {
  if (!tflops) return set_err(BH_ERR_ARG, "null argument");
  if (bh_device_count() < 1) return set_err(BH_ERR_NO_DEVICE, "no CUDA device visible");
  int dev = 0, nsm = 0, khz = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
  const int threads = 256, blocks = nsm * 4, iters = 4096;
  double* d = nullptr;
  BH_CUDA(cudaMalloc((void**)&d, sizeof(double) * threads * blocks));
  cudaEvent_t e0, e1;
  BH_CUDA(cudaEventCreate(&e0));
  BH_CUDA(cudaEventCreate(&e1));
  fp64_peak_kernel<<<blocks, threads>>>(d, 1.0e-3, 0.999, iters / 8);       // warm-up
  float best = 1e30f;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    fp64_peak_kernel<<<blocks, threads>>>(d, 1.0e-3, 0.999, iters);
    cudaEventRecord(e1);
    cudaError_t ce = cudaEventSynchronize(e1);
    if (ce != cudaSuccess) { cudaFree(d); return set_err(BH_ERR_CUDA, "fp64 peak probe", ce); }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
  const double flop = 2.0 * 8 * 16 * (double)iters * threads * blocks;
  *tflops = flop / (best * 1e-3) / 1e12;
  if (sm_mhz_seen) *sm_mhz_seen = khz / 1e3;
  return BH_OK;
}

extern "C" int bh_debug_math(int n, const double* x, double* out) {
  if (n < 1 || !x || !out) return set_err(BH_ERR_ARG, "bad argument");
  if (bh_device_count() < 1) return set_err(BH_ERR_NO_DEVICE, "no CUDA device visible");
  double *dx = nullptr, *dout = nullptr;
  BH_CUDA(cudaMalloc((void**)&dx, sizeof(double) * n));
  BH_CUDA(cudaMalloc((void**)&dout, sizeof(double) * 7 * n));
  BH_CUDA(cudaMemcpy(dx, x, sizeof(double) * n, cudaMemcpyHostToDevice));
  debug_math_kernel<<<(n + 127) / 128, 128>>>(dx, dout, n);
  cudaError_t ce = cudaMemcpy(out, dout, sizeof(double) * 7 * n, cudaMemcpyDeviceToHost);
  cudaFree(dx); cudaFree(dout);
  if (ce != cudaSuccess) return set_err(BH_ERR_CUDA, "debug_math", ce);
  return BH_OK;
}
