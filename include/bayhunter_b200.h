/*
 * bayhunter_b200.h -- C ABI of the B200-native forward-model + likelihood engine
 * for BayHunter's hot path (libbayhunter_b200.so, sm_100a).
 *
 * Plain C: pointers and sizes only, no torch / C++ types.  Device pointers are
 * whatever the caller owns (e.g. torch.Tensor.data_ptr()); the library never
 * allocates caller-visible memory and keeps no caller pointer after a call
 * returns (same ownership rule as the reference FFI).  Internal
 * scratch lives inside the opaque engine handle.
 *
 * What each entry point replaces in the reference (paths relative to the
 * BayHunter tree):
 *
 *   bh_surfdisp96        <- Fortran `surfdisp96` called through f2py
 *                           (src/extensions/surfdisp96.f:55-56, called from
 *                           src/surf96_modsw.py:116-117)
 *   bh_synrf             <- extern "C" `synrf_cwrap`
 *                           (src/extensions/rfmini/wrap.cpp:57-80, called from
 *                           src/extensions/rfmini/rfmini.pyx:111-112 and
 *                           src/rfmini_modrf.py:134-137)
 *   bh_engine_eval       <- the per-chain, per-iteration body of
 *                           JointTarget.evaluate (src/Targets.py:314-347) incl.
 *                           Model.get_vp_vs_h's vp rule (src/Models.py:40-52),
 *                           the plugins' run_model (src/surf96_modsw.py:84-126,
 *                           src/rfmini_modrf.py:99-154) and the covariance laws
 *                           (src/Targets.py:105-183), for B chains per call
 *   bh_engine_eval_host  <- same, host buffers in / host buffers out (what a
 *                           ctypes / cgo / JNI binding with numpy-like arrays calls)
 *
 * All functions return BH_OK (0) or a negative error code; no exceptions and no
 * process exit cross this boundary.  Numerical failure is reported in-band the
 * way the reference does: status[b] = 0, logL[b] = -1e15, misfits[b][*] = 1e15
 * (src/Targets.py:325-328).
 */
#ifndef BAYHUNTER_B200_H
#define BAYHUNTER_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define BH_ABI_VERSION 1
#define BH_MAX_TARGETS 8
#define BH_MAX_PERIODS 60    /* NP, surfdisp96.f:62 */
#define BH_MAX_LAYERS 100    /* NL, surfdisp96.f:60 */

/* error codes */
#define BH_OK 0
#define BH_ERR_ARG (-1)          /* bad argument (null pointer, size out of range) */
#define BH_ERR_CUDA (-2)         /* CUDA runtime error; see bh_last_error() */
#define BH_ERR_UNSUPPORTED (-3)  /* feature outside this engine's scope */
#define BH_ERR_NO_DEVICE (-4)    /* no CUDA device: this library has no CPU path */

/* target kinds: BayHunter `ref` strings (src/Targets.py:52-53) */
#define BH_REF_RDISPPH 0
#define BH_REF_RDISPGR 1
#define BH_REF_LDISPPH 2
#define BH_REF_LDISPGR 3
#define BH_REF_PRF 4
#define BH_REF_SRF 5
#define BH_REF_GENERIC 6   /* a data set whose forward model is the caller's (templates/myfwd.py plugins,
                              per-layer Q arrays): only bh_engine_loglik_host evaluates such a target */

/* covariance laws bound per target at chain init (src/SingleChain.py:159-205) */
#define BH_COV_EXP 0            /* Valuation.get_covariance_exp              */
#define BH_COV_WHITE 1          /* Valuation.get_covariance_nocorr           */
#define BH_COV_WHITE_SCALED 2   /* Valuation.get_covariance_nocorr_scalederr */
#define BH_COV_GAUSS 3          /* Valuation.get_covariance_gauss            */

/* One observed data set + its forward-model parameters.  All pointers are
 * HOST pointers, read during bh_engine_create only. */
typedef struct bh_target {
  int ref;                 /* BH_REF_*                                        */
  int n;                   /* number of observed samples                      */
  const double* x;         /* [n] periods (s) or time axis (s)                */
  const double* y;         /* [n] observed data                               */
  const double* yerr;      /* [n] or NULL (only BH_COV_WHITE_SCALED reads it) */
  int cov;                 /* BH_COV_*                                        */
  const double* corr_inv;  /* [n*n] row-major R^-1 for BH_COV_GAUSS else NULL */
  double logcorr_det;      /* slogdet(R) for BH_COV_GAUSS                     */
  /* SurfDisp.set_modelparams keys (src/surf96_modsw.py:28-31) */
  int mode;                /* number of modes searched, 1 = fundamental; the curve
                              returned is the highest one (surfdisp96.f:223-311)  */
  int flsph;               /* 0 = flat earth, 1 = earth-flattening (sphere)    */
  /* RFminiModRF.set_modelparams keys (src/rfmini_modrf.py:26-31) */
  double gauss;            /* Gauss parameter a                               */
  double p;                /* slowness, s/deg                                 */
  double nsv;              /* near-surface vs for the rotation; <= 0: vs[0]   */
  double qp, qs;           /* layer-independent Q (<= 0: 500 / 225)           */
} bh_target;

typedef struct bh_engine bh_engine;

/* Library / device probes (no device work). */
int bh_abi_version(void);
const char* bh_last_error(void);
int bh_device_count(void);
/* Device of the calling thread for everything this library creates afterwards (one process per GPU:
 * pass LOCAL_RANK).  Without it the library adopts the thread's current CUDA context, else device 0. */
int bh_set_device(int device);

/*
 * Create an engine for a fixed joint target set.  Uploads observed data,
 * periods and R^-1 once; allocates scratch for up to max_batch models of up
 * to max_layers rows (half-space included) on the current CUDA device.
 */
int bh_engine_create(const bh_target* targets, int ntargets, int max_batch,
                     int max_layers, bh_engine** out);
void bh_engine_destroy(bh_engine* e);

/* Size of one model's synthetic-data row: sum of n over targets. */
int bh_engine_synth_stride(const bh_engine* e);

/* 1 while "swd_autotune" is still timing candidates for the current batch size (a caller that enqueues
 * evaluations without ever waiting can synchronise its stream between them until this returns 0, so
 * that every timing is read back before the next candidate is tried). */
int bh_engine_is_tuning(const bh_engine* e);

/* CUDA-graph capture.  Between bh_engine_capture_begin and bh_engine_capture_end, bh_engine_eval enqueues stream-ordered
 * work only (no event polling, no read-back; the launch layout of the last plain evaluation is kept and deeper models are
 * still handled by the second dispersion launch), so the caller may capture it on its stream with cudaStreamBeginCapture /
 * cudaStreamEndCapture and replay the graph; the forked streams join the capture through events.  _begin fails with
 * BH_ERR_UNSUPPORTED while "profile" is on or bh_engine_is_tuning() is 1.  bh_sampler_run uses this for its iterations. */
int bh_engine_capture_begin(bh_engine* e);
int bh_engine_capture_end(bh_engine* e);

/* Tunables (call before eval; all have working defaults).
 *   key "swd_searches_per_warp"  1..32   phase-velocity curves (default chosen from the batch size)
 *   key "swd_group_searches_per_warp" 1..32  group-velocity curves (default: half of the above)
 *   key "swd_spw_rg" / "_rp" / "_lg" / "_lp"  models per warp of one curve type (Rayleigh/Love,
 *                                        group/phase); 0 = the two keys above (default)
 *   key "swd_sort_layers"        0/1     deal models to the dispersion warps sorted by layer count, so
 *                                        that a warp's lanes run layer loops of similar length (default 1;
 *                                        results do not depend on it)
 *   key "swd_adaptive_capacity"  0/1     size the dispersion kernel's per-warp layer records by the layer
 *                                        counts of recent batches instead of lmax; deeper models are
 *                                        handled by a second launch (default 1; results do not depend on it)
 *   key "swd_autotune"           0/1     time the neighbours of the models-per-warp rule's pick on the
 *                                        first evaluations of a batch size and keep the fastest (default 0;
 *                                        the device sampler switches it on; results do not depend on it)
 *   key "swd_direct"             0/1/2   chains evaluate only their own candidate: never (default) /
 *                                        when warps are full of chains / always
 *   key "swd_rayleigh_sm_pct"    0..100  one launch: share of the SMs whose CTAs take the Rayleigh work
 *                                        items (the rest take Love; 0 = no partition, the default)
 *   key "swd_split_waves"        0/1     Rayleigh and Love curves in one launch (default) or two
 *   key "swd_max_spec"           1..32   speculative bracket candidates per search
 *   key "swd_pool"               -1/0/1  dispersion by swd_pool_kernel (a CTA's 128 lanes dealt over all chains of M
 *                                        models of one wave type, Rayleigh and Love launches side by side): by rule
 *                                        (default: from ~2.5 k models with both wave types, ~1.8 k deep models or 4-8 k shallow ones with one), never, always
 *                                        (results do not depend on it)
 *   key "swd_pool_models"        0..128  M of the above (0 = rule: as many chains per CTA as fill four CTAs per SM)
 *   key "swd_lockstep"           0/1     dispersion by swd_lockstep_kernel (every lane owns a chain; a measured
 *                                        negative result at these batch sizes, default 0)
 *   key "rf_prune_exp10"         0..300  receiver function: spectral bins whose Gauss-filter weight
 *                                        exp(-(w/2a)^2) is below 10^-value are not computed (they enter the
 *                                        inverse transform as 0); default 20, i.e. 1e-20 of the passband --
 *                                        four orders of magnitude below fp64 resolution of the trace; 0 computes
 *                                        every bin
 *   key "rf_gate_pct"            0..100  the forked RF stream starts once this share of the dispersion
 *                                        warps has retired (default 25; 0: no gate, the streams race)
 *   key "concurrent"             0/1     run SWD and RF kernels on forked streams
 *   key "profile"                0/1     record per-kernel event timings */
int bh_engine_set(bh_engine* e, const char* key, int value);

/*
 * Evaluate B layered models (one per chain), DEVICE pointers.
 *   model   [B][lmax][4] fp64 rows (vs, vp/vs, z_top, h); rows >= nlay[b] ignored
 *   nlay    [B] int32, rows per model incl. half-space (h of last row ignored)
 *   noise   [B][2T] fp64 (corr_0, sigma_0, corr_1, sigma_1, ...)  (Targets.py:335)
 *   rho     [B][lmax] fp64 or NULL (NULL: rho = 0.32*vp + 0.77, Targets.py:319)
 *   logL    [B] fp64 out
 *   misfits [B][T+1] fp64 out (last = joint)
 *   status  [B] int32 out (1 valid, 0 invalid -> sentinels written)
 *   synth   [B][synth_stride] fp64 out or NULL (modelled data, targets back to back)
 *   stream  cudaStream_t (as void*); work is enqueued, not synchronised
 */
int bh_engine_eval(bh_engine* e, const double* model, const int* nlay,
                   const double* noise, const double* rho, int B, int lmax,
                   double* logL, double* misfits, int* status, double* synth,
                   void* stream);

/* Same with HOST pointers (what a ctypes / cgo / JNI binding holding numpy-like arrays calls).
 * The engine owns two slots of pinned staging buffers and device mirrors: pageable caller memory is
 * copied into the slot's pinned buffers on the calling thread (pinned caller memory -- cudaHostAlloc /
 * cudaHostRegister / torch pin_memory -- is used in place), the H2D copies, the kernels and the D2H
 * copies run on three streams, and results land in pinned memory first.
 *
 * bh_engine_eval_host        submit + wait: returns with the outputs written.
 * bh_engine_eval_host_async  submits and returns a ticket; the copies of this call overlap the
 *                            kernels of the previous one.  The caller's input AND output buffers must
 *                            stay valid until bh_engine_wait(ticket) returns; at most two calls are in
 *                            flight -- a third submit first completes (and delivers) the oldest.
 * bh_engine_wait             blocks until that call's outputs are in the caller's buffers. */
int bh_engine_eval_host(bh_engine* e, const double* model, const int* nlay,
                        const double* noise, const double* rho, int B, int lmax,
                        double* logL, double* misfits, int* status, double* synth);
int bh_engine_eval_host_async(bh_engine* e, const double* model, const int* nlay,
                              const double* noise, const double* rho, int B, int lmax,
                              double* logL, double* misfits, int* status, double* synth,
                              long long* ticket);
int bh_engine_wait(bh_engine* e, long long ticket);

/* Likelihood only, HOST pointers: the modelled data of every target are the caller's (a forward-model
 * plugin of its own, src/templates/myfwd.py; or this library's shims called with extra arguments) and the
 * engine evaluates what JointTarget.evaluate does with them (src/Targets.py:325-347): validity, RMS misfits,
 * the bound covariance laws, the joint log-likelihood, the sentinels.
 *   synth  [B][synth_stride] modelled data, targets back to back;  tvalid [B][T] int32: 0 = that target's
 *   synthetic was rejected (SingleTarget._moddata_valid, src/Targets.py:204-214) -> sentinels for the model. */
int bh_engine_loglik_host(bh_engine* e, const double* synth, const int* tvalid, const double* noise,
                          int B, double* logL, double* misfits, int* status);

/* Per-kernel device time of the last eval, measured with CUDA events on the
 * launching streams; needs bh_engine_set(e, "profile", 1) before that eval.
 * ms[BH_NUM_KERNELS], index BH_K_*; -1 for kernels that did not run.  With several
 * RF targets the RF entries hold the last one. Synchronises on the stop events. */
#define BH_K_PREP_SWD 0
#define BH_K_SWD 1
#define BH_K_PREP_RF 2
#define BH_K_RF_SPECTRUM 3
#define BH_K_RF_SYNTH 4
#define BH_K_LOGLIK 5
#define BH_K_SWD_LOVE 6   /* BH_K_SWD is the Rayleigh launch */
#define BH_K_SWD_GENERAL 7 /* higher modes / flsph = 1 / water-layer models */
#define BH_K_SWD_POOL 8      /* swd_pool_kernel, Rayleigh launch (full batches; replaces BH_K_SWD) */
#define BH_K_SWD_POOL_LOVE 9 /* swd_pool_kernel, Love launch, concurrent with the Rayleigh one */
#define BH_NUM_KERNELS 10
int bh_engine_last_kernel_ms(bh_engine* e, float* ms);

/* Work counters of the last eval, nsec[BH_NUM_COUNTERS] host ints:
 *   [0] secular-function values consumed by the searches, [1] evaluated (>= [0]
 *   with speculation); then per dispersion curve c (in launch order: Rayleigh
 *   group, Rayleigh phase, Love group, Love phase): [2+2c] sum over warps of the
 *   evaluation rounds, [3+2c] the maximum over warps. */
#define BH_NUM_COUNTERS (2 + 2 * BH_MAX_TARGETS)
int bh_engine_last_counts(bh_engine* e, long long* nsec);

/*
 * Single-model shims with the argument meaning of the reference FFI (HOST
 * pointers).  They run the same CUDA kernels with B = 1 (mode > 1, iflsph = 1 and
 * models with a water layer through the general dispersion kernel).
 *
 * bh_surfdisp96: thkm/vpm/vsm/rhom REAL*4 [nlayer]; t, cg fp64 [kmax];
 *   iwave 1 Love / 2 Rayleigh; igr 0 phase / >0 group; *err = 0 ok, 1 no root.
 * bh_synrf: as synrf_cwrap; fz/fr may be NULL (BayHunter discards them; when
 *   given they are zero-filled), rf [nsamp]; returns BH_OK instead of 1.
 */
int bh_surfdisp96(const float* thkm, const float* vpm, const float* vsm,
                  const float* rhom, int nlayer, int iflsph, int iwave, int mode,
                  int igr, int kmax, const double* t, double* cg, int* err);
int bh_synrf(int nsamp, double fsamp, double tshift, double p, double a,
             double nsv, double sigma, int waveno, int nlay, const double* z,
             const double* vp, const double* vs, const double* rh,
             const double* qp, const double* qs, double* fz, double* fr,
             double* rf);

/*
 * The reference's raw native symbols, for LINKING its own glue unchanged:
 *   surfdisp96_   gfortran name of `subroutine surfdisp96` (src/extensions/surfdisp96.f:55-56, :82-86,
 *                 :101): every argument by reference, real*4 thkm/vpm/vsm/rhom(100), double precision
 *                 t(60), cg(60), integer err out.  What the f2py-generated module calls.
 *   synrf_cwrap   src/extensions/rfmini/wrap.cpp:26-31, :57-80: same prototype, returns 1.  What
 *                 rfmini.pyx calls.
 * Both forward to the shims above; a library error (no device) is reported on stderr and in-band
 * (err = 1 / NaN trace), as neither prototype has another channel.
 */
void surfdisp96_(const float* thkm, const float* vpm, const float* vsm, const float* rhom,
                 const int* nlayer, const int* iflsph, const int* iwave, const int* mode,
                 const int* igr, const int* kmax, const double* t, double* cg, int* err);
int synrf_cwrap(int nsamp, double fsamp, double tshift, double p, double a, double nsv,
                double sigma, int waveno, int nlay, double* z, double* vp, double* vs,
                double* rh, double* qp, double* qs, double* fz, double* fr, double* rf);

/* ------------------------------------------------------------------------
 * Lock-step chain ensemble: the sampler around the hot path, on the device.
 *
 *   bh_sampler_*  <- SingleChain.iterate / run_chain (src/SingleChain.py:511-589,
 *                    591-612) for B chains at once, and the chain arrays of
 *                    MCMC_Optimizer._init_shareddata (src/mcmcOptimizer.py:78-128).
 *
 * Every chain keeps BayHunter's state (Voronoi nuclei, vp/vs, noise parameters,
 * proposal widths, accepted/proposed counters, iteration counter) in device
 * memory; one iteration = propose kernel -> bh_engine_eval -> accept kernel, with
 * no host round trip.  Random variates come from Philox4x32-10 keyed by
 * (seed, global chain index, iteration): a chain's trajectory does not depend on
 * the batch it runs in or on the GPU count.
 * ------------------------------------------------------------------------ */
typedef struct bh_sampler_config {
  /* priors (src/defaults/defaults.ini [modelpriors]) */
  int layers_min, layers_max;          /* priors['layers']; models hold up to layers_max + 1 nuclei */
  double vs_min, vs_max;               /* priors['vs'] */
  double z_min, z_max;                 /* priors['z'] */
  int vpvs_fixed;                      /* priors['vpvs'] given as a float: never perturbed */
  double vpvs_min, vpvs_max;
  int has_mantle;                      /* priors['mantle'] = (vs_m, vpvs_m) */
  double mantle_vs, mantle_vpvs;
  int noise_fixed[2 * BH_MAX_TARGETS]; /* per target (corr, sigma): prior given as a float */
  double noise_min[2 * BH_MAX_TARGETS], noise_max[2 * BH_MAX_TARGETS];
  /* initparams */
  double thickmin;
  int has_lvz, has_hvz;                /* initparams['lvz'] / ['hvz'] not None */
  double lvz, hvz;
  double propdist[5];                  /* vs, z, birth/death, noise, vpvs */
  double acceptance[2];                /* percent */
  int iter_burnin, iter_main;
  int max_accepted;                    /* rows of the chain arrays per chain (reference:
                                          iterations * max(acceptance) / 100); further accepted
                                          models are counted in `overflow` and not stored */
  unsigned long long seed;
} bh_sampler_config;

typedef struct bh_sampler bh_sampler;

/* nchains chains with global indices first_chain .. first_chain + nchains - 1 on the
 * engine's device.  ntargets must equal the engine's.  The engine must outlive the sampler
 * and have max_batch >= nchains, max_layers >= layers_max + 1. */
int bh_sampler_create(bh_engine* e, const bh_sampler_config* cfg, int ntargets, int nchains,
                      long long first_chain, bh_sampler** out);
void bh_sampler_destroy(bh_sampler* s);

/* Initial state, HOST pointers: models [B][2*(layers_max+1)] with vs of the k[b] nuclei at
 * [0, k) of the first half and their depths at [0, k) of the second half, ordered by depth;
 * vpvs [B]; noise [B][2T].  Evaluates the models and makes them row 0 of the chain arrays
 * (SingleChain._init_model_and_currentvalues, src/SingleChain.py:70-92). */
int bh_sampler_init(bh_sampler* s, const double* models, const int* k, const double* vpvs,
                    const double* noise);

/* niter lock-step iterations of every chain; returns when they are done. */
int bh_sampler_run(bh_sampler* s, int niter);

/* Current state to HOST buffers (any pointer may be NULL): models/k/vpvs/noise as in
 * bh_sampler_init, logL [B], misfits [B][T+1], propdist [B][5], accepted/proposed [B][5],
 * iiter [B], nstored [B] rows used in the chain arrays, overflow [1]. */
int bh_sampler_get_state(bh_sampler* s, double* models, int* k, double* vpvs, double* noise,
                         double* logL, double* misfits, double* propdist, long long* accepted,
                         long long* proposed, long long* iiter, int* nstored, long long* overflow);

/* Per chain (HOST, [B] each, either may be NULL): accepted models that did not fit the chain arrays
 * (max_accepted rows) and the iteration of the first of them.  A chain that overflowed has a complete,
 * correctly weighted record up to that iteration only. */
int bh_sampler_get_overflow(bh_sampler* s, long long* count, long long* first_iter);

/* Chain arrays of chains [chain0, chain0 + nchain) to HOST buffers (float32, NaN padded, the
 * layout of the reference's shared arrays): models [n][S][2*(layers_max+1)] (2k values, then
 * NaN), misfits [n][S][T+1], likes [n][S], noise [n][S][2T], vpvs [n][S]; iters [n][S] int32 is
 * SingleChain.chainiter (iteration at which the row was accepted). */
int bh_sampler_get_chains(bh_sampler* s, int chain0, int nchain, float* models, float* misfits,
                          float* likes, float* noise, float* vpvs, int* iters);

/* Testing / resuming: overwrite the whole state (NULL = keep), read the last proposal, and
 * replace the Philox variates of the following iterations by draws [B][4] =
 * (u_mod, u_idx, gauss, u_acc) (NULL switches back to Philox). */
int bh_sampler_set_state(bh_sampler* s, const double* models, const int* k, const double* vpvs,
                         const double* noise, const double* logL, const double* misfits,
                         const double* propdist, const long long* accepted,
                         const long long* proposed, const long long* iiter);
int bh_sampler_get_proposal(bh_sampler* s, double* models, int* k, double* vpvs, double* noise,
                            int* valid, int* modify, double* dvs2, double* logL, double* misfits);
int bh_sampler_set_forced_draws(bh_sampler* s, const double* draws);

/* Correlated-noise realisations for synthetic observations (src/SynthObs.py:136-155), HOST out [B][n]:
 * B draws of N(0, sigma^2 R), R_ij = corr^|i-j| (law BH_COV_EXP, compute_expnoise; exact AR(1) recursion) or
 * R_ij = corr^((i-j)^2) (law BH_COV_GAUSS, compute_gaussnoise; `factor` [n][n] HOST with factor^T factor = R,
 * computed by the caller like numpy.random.multivariate_normal does: sqrt(s) v of an SVD).  Philox4x32-10
 * keyed by (seed, realisation): a realisation does not depend on B. */
int bh_correlated_noise(int law, int n, int B, double corr, double sigma, unsigned long long seed,
                        const double* factor, double* out);

/* Diagnostics: evaluates the engine's straight-line fp64 elementary functions on
 * the device for n HOST values x; out[7][n] = exp(-|x|), sin x, cos x, 1/x,
 * sqrt|x|, 1/sqrt|x|, 1.0/x (faithful division).  Used by the accuracy tests. */
int bh_debug_math(int n, const double* x, double* out);

/* Measured fp64 peak of the current device: a kernel of independent DFMA chains on every SM,
 * best of three, in TFLOP/s (2 flop per DFMA); *sm_mhz_seen (may be NULL) = the device's nominal
 * clock.  The roofline denominator of bench.py. */
int bh_measure_fp64_peak(double* tflops, double* sm_mhz_seen);

#ifdef __cplusplus
}
#endif
#endif /* BAYHUNTER_B200_H */
