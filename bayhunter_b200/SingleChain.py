"""Lock-step chain ensemble: host side of the device sampler (bh_sampler_* in the C ABI).

The reference runs one `SingleChain` per OS process (src/SingleChain.py, src/mcmcOptimizer.py:
219-252).  Here all chains of a rank advance together on the GPU; this module

* turns BayHunter's `priors` / `initparams` dictionaries into `bh_sampler_config`,
* binds the covariance law of every target the way SingleChain.set_target_covariance does
  (src/SingleChain.py:159-205),
* draws the initial model / vpvs / noise of every chain with the reference's own sequence of
  numpy RandomState calls (draw_initvpvs :151-156, draw_initmodel :94-123, draw_initnoiseparams
  :125-149), so that chain i starts where the reference's chain i with the same seed starts,
* turns the device chain arrays into the weighted, thinned per-chain files the reference saves
  (get_weightedvalues :646-666, save_finalmodels :668-690).

The iterations themselves (proposal, prior check, forward models, likelihood, acceptance,
proposal-width control) are CUDA kernels; nothing here computes a likelihood.
"""
import ctypes
import os.path as op

import numpy as np

from . import _lib
from .Models import Model

PAR_MAP = {'vsmod': 0, 'zvmod': 1, 'birth': 2, 'death': 2, 'noise': 3, 'vpvs': 4}
MODIFICATIONS = ('vsmod', 'zvmod', 'birth', 'death', 'noise', 'vpvs')

DEFAULT_PRIORS = dict(mantle=None, vpvs=(1.5, 2.1), layers=(1, 20), vs=(1, 5), z=(0, 60), mohoest=None,
                      rfnoise_corr=(0.35, 0.75), rfnoise_sigma=(1e-5, 0.05), swdnoise_corr=0.,
                      swdnoise_sigma=(1e-5, 0.1))                  # src/defaults/defaults.ini
DEFAULT_INITPARAMS = dict(nchains=3, iter_burnin=2048 * 2, iter_main=2048 * 1,
                          propdist=(0.025, 0.025, 0.015, 0.005, 0.005), acceptance=(40, 45), thickmin=0.,
                          lvz=None, hvz=None, rcond=None, station='test', savepath='results/',
                          maxmodels=50000)


def _is_fixed(prior):
    return isinstance(prior, (int, float, np.floating, np.integer))


def noise_priors(targets, priors):
    """[(corr prior), (sigma prior)] per target, in the order of the noise vector
    (SingleChain.draw_initnoiseparams, src/SingleChain.py:125-149)."""
    out = []
    for t in targets.targets:
        for name in ('noise_corr', 'noise_sigma'):
            out.append(priors[t.noiseref + name])
    return out


def set_target_covariance(targets, priors, rcond=None):
    """Bind `target.get_covariance` per target (src/SingleChain.py:159-205)."""
    np_ = noise_priors(targets, priors)
    for i, target in enumerate(targets.targets):
        corr = np_[2 * i]
        if not _is_fixed(corr):
            target.get_covariance = target.valuation.get_covariance_exp
        elif corr == 0 and np.any(np.isnan(target.obsdata.yerr)):
            target.get_covariance = target.valuation.get_covariance_nocorr
        elif corr == 0:
            target.get_covariance = target.valuation.get_covariance_nocorr_scalederr
        elif target.noiseref == 'rf':
            target.valuation.init_covariance_gauss(corr, target.obsdata.x.size, rcond=rcond)
            target.get_covariance = target.valuation.get_covariance_gauss
        else:
            target.get_covariance = target.valuation.get_covariance_exp


def make_config(targets, priors, initparams, seed=0, max_accepted=None, nchains=None, max_chain_bytes=32 << 30):
    """BayHunter dictionaries -> struct bh_sampler_config."""
    c = _lib.BhSamplerConfig()
    c.layers_min, c.layers_max = int(priors['layers'][0]), int(priors['layers'][1])
    c.vs_min, c.vs_max = float(priors['vs'][0]), float(priors['vs'][1])
    c.z_min, c.z_max = float(priors['z'][0]), float(priors['z'][1])
    if _is_fixed(priors['vpvs']):
        c.vpvs_fixed, c.vpvs_min, c.vpvs_max = 1, float(priors['vpvs']), float(priors['vpvs'])
    else:
        c.vpvs_fixed, c.vpvs_min, c.vpvs_max = 0, float(priors['vpvs'][0]), float(priors['vpvs'][1])
    if priors.get('mantle') is not None:
        c.has_mantle, c.mantle_vs, c.mantle_vpvs = 1, float(priors['mantle'][0]), float(priors['mantle'][1])
    for i, p in enumerate(noise_priors(targets, priors)):
        if _is_fixed(p):
            c.noise_fixed[i], c.noise_min[i], c.noise_max[i] = 1, float(p), float(p)
        else:
            c.noise_fixed[i], c.noise_min[i], c.noise_max[i] = 0, float(p[0]), float(p[1])
    for i in range(2 * targets.ntargets, 2 * _lib.MAX_TARGETS):
        c.noise_fixed[i] = 1
    c.thickmin = float(initparams['thickmin'])
    if initparams.get('lvz') is not None:
        c.has_lvz, c.lvz = 1, float(initparams['lvz'])
    if initparams.get('hvz') is not None:
        c.has_hvz, c.hvz = 1, float(initparams['hvz'])
    for i, v in enumerate(initparams['propdist']):
        c.propdist[i] = float(v)
    c.acceptance[0], c.acceptance[1] = float(initparams['acceptance'][0]), float(initparams['acceptance'][1])
    c.iter_burnin, c.iter_main = int(initparams['iter_burnin']), int(initparams['iter_main'])
    if max_accepted is None:
        # The reference sizes its chain arrays as iterations * max(acceptance) / 100
        # (mcmcOptimizer.py:86-88) and dies with an IndexError when a chain accepts more.  Device
        # memory is cheap: keep one row per iteration while the arrays stay below 16 GB, else the
        # reference's size plus a quarter -- but never more than `max_chain_bytes` (32 GB of the B200's
        # 180 GB) for the whole ensemble: beyond that accepted models are counted as overflow, not stored.
        iterations = c.iter_burnin + c.iter_main
        row_bytes = 4 * (2 * (c.layers_max + 1) + 3 * targets.ntargets + 4)
        n = int(nchains) if nchains is not None else int(initparams.get('nchains', 1))
        if (iterations + 1) * row_bytes * n <= (16 << 30):
            max_accepted = iterations + 1
        else:
            max_accepted = int(1.25 * iterations * np.max(initparams['acceptance']) / 100.) + 16
            max_accepted = min(max_accepted, max(64, int(max_chain_bytes // (row_bytes * n))))
    c.max_accepted = max(int(max_accepted), 1)
    c.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    return c


class InitialStateDrawer(object):
    """The reference's initial draws of ONE chain, call for call (src/SingleChain.py:70-156)."""

    def __init__(self, targets, priors, initparams, random_seed):
        self.rstate = np.random.RandomState(random_seed)
        self.targets, self.priors, self.initparams = targets, priors, initparams

    def _validmodel(self, model, vpvs):
        p, ip = self.priors, self.initparams
        vp, vs, h = Model.get_vp_vs_h(model, vpvs, p.get('mantle'))
        layers = h.size - 1
        if not (p['layers'][0] <= layers <= p['layers'][1]):
            return False
        if np.any(h[:-1] < ip['thickmin']):
            return False
        if np.any(vs < p['vs'][0]) or np.any(vs > p['vs'][1]):
            return False
        z = np.cumsum(h)
        if np.any(z < p['z'][0]) or np.any(z > p['z'][1]):
            return False
        if ip.get('lvz') is not None:
            comp = vs[1:] - (vs[:-1] * (1 - ip['lvz']))
            if not comp.size == comp[comp > 0].size:
                return False
        if ip.get('hvz') is not None:
            comp = (vs[:-1] * (1 + ip['hvz'])) - vs[1:]
            if not comp.size == comp[comp > 0].size:
                return False
        return True

    def draw_initvpvs(self):
        if _is_fixed(self.priors['vpvs']):
            return self.priors['vpvs']
        return self.rstate.uniform(low=self.priors['vpvs'][0], high=self.priors['vpvs'][1])

    def draw_initmodel(self, vpvs):
        p = self.priors
        zmin, zmax = p['z']
        vsmin, vsmax = p['vs']
        layers = p['layers'][0] + 1
        while True:
            vs = self.rstate.uniform(low=vsmin, high=vsmax, size=layers)
            vs.sort()
            if p.get('mohoest') is not None and layers > 1:
                mean, std = p['mohoest']
                moho = self.rstate.normal(loc=mean, scale=std)
                tmp_z = self.rstate.uniform(1, np.min([5, moho]))
                tmp = [moho - tmp_z, moho + tmp_z]
                z = tmp if layers - 2 == 0 else np.concatenate(
                    (tmp, self.rstate.uniform(low=zmin, high=zmax, size=(layers - 2))))
                z = np.asarray(z, dtype=float)
            else:
                z = self.rstate.uniform(low=zmin, high=zmax, size=layers)
            z.sort()
            model = np.concatenate((vs, z))
            if self._validmodel(model, vpvs):
                return model

    def draw_initnoiseparams(self):
        pri = noise_priors(self.targets, self.priors)
        noise = np.ones(len(pri)) * np.nan
        for i, p in enumerate(pri):
            noise[i] = p if _is_fixed(p) else self.rstate.uniform(low=p[0], high=p[1])
        return noise

    def draw(self):
        vpvs = self.draw_initvpvs()
        model = self.draw_initmodel(vpvs)
        noise = self.draw_initnoiseparams()
        return model, float(vpvs), noise


class ChainEnsemble(object):
    """B chains advancing in lock step on one GPU.

    targets        bayhunter_b200.Targets.JointTarget (laws are bound here)
    priors, initparams  BayHunter dictionaries (defaults of defaults.ini filled in)
    nchains        chains on THIS device; first_chain = global index of the first one
    chain_seeds    one numpy seed per chain for the initial draws (reference: rstate.randint(1000)
                   per chain, src/mcmcOptimizer.py:136); the device stream is keyed by `seed`
    """

    def __init__(self, targets, priors=None, initparams=None, nchains=None, first_chain=0, seed=0,
                 chain_seeds=None, max_accepted=None):
        self._lib = _lib.require_device()
        self.priors = dict(DEFAULT_PRIORS); self.priors.update(priors or {})
        self.initparams = dict(DEFAULT_INITPARAMS); self.initparams.update(initparams or {})
        self.targets = targets
        self.nchains = int(nchains if nchains is not None else self.initparams['nchains'])
        self.first_chain = int(first_chain)
        self.ntargets = targets.ntargets
        self.maxlayers = int(self.priors['layers'][1]) + 1
        set_target_covariance(targets, self.priors, self.initparams.get('rcond'))
        self.config = make_config(targets, self.priors, self.initparams, seed=seed, max_accepted=max_accepted,
                                  nchains=self.nchains)
        self.nmodels = int(self.config.max_accepted)
        # an engine of its own: the device sampler keeps the raw handle, and JointTarget's cached engine is
        # closed and rebuilt whenever another caller asks for a larger batch or a re-bound law
        self.engine = targets.new_engine(self.nchains, self.maxlayers)
        h = ctypes.c_void_p()
        _lib.check(self._lib.bh_sampler_create(self.engine._h, ctypes.byref(self.config), self.ntargets,
                                               self.nchains, self.first_chain, ctypes.byref(h)))
        self._h = h
        if chain_seeds is None:
            chain_seeds = np.random.RandomState(seed).randint(1000, size=self.first_chain + self.nchains)[self.first_chain:]
        self.chain_seeds = np.asarray(chain_seeds)
        self._initialised = False

    def close(self):
        if getattr(self, "_h", None):
            self._lib.bh_sampler_destroy(self._h)
            self._h = None
        if getattr(self, "engine", None) is not None:
            self.engine.close()
            self.engine = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- state in / out -----------------------------------------------------
    def _pad_models(self, models):
        B, L = self.nchains, self.maxlayers
        out = np.zeros((B, 2 * L))
        k = np.zeros(B, dtype=np.int32)
        for b, m in enumerate(models):
            m = np.asarray(m, dtype=np.float64)
            m = m[~np.isnan(m)]
            n = m.size // 2
            out[b, :n] = m[:n]
            out[b, L:L + n] = m[n:]
            k[b] = n
        return out, k

    def draw_initial(self):
        """(models list, vpvs [B], noise [B,2T]) from the reference's initial-draw sequence."""
        models, vpvs, noise = [], np.zeros(self.nchains), np.zeros((self.nchains, 2 * self.ntargets))
        for b in range(self.nchains):
            m, v, n = InitialStateDrawer(self.targets, self.priors, self.initparams, int(self.chain_seeds[b])).draw()
            models.append(m); vpvs[b] = v; noise[b] = n
        return models, vpvs, noise

    def init(self, models=None, vpvs=None, noise=None):
        """Set and evaluate the initial state (SingleChain._init_model_and_currentvalues)."""
        if models is None:
            models, vpvs, noise = self.draw_initial()
        pm, k = self._pad_models(models)
        vpvs = np.ascontiguousarray(np.broadcast_to(vpvs, (self.nchains,)), dtype=np.float64)
        noise = np.ascontiguousarray(noise, dtype=np.float64).reshape(self.nchains, 2 * self.ntargets)
        _lib.check(self._lib.bh_sampler_init(self._h, pm.ctypes.data, k.ctypes.data, vpvs.ctypes.data, noise.ctypes.data))
        self._initialised = True

    def run(self, niter):
        if not self._initialised:
            self.init()
        _lib.check(self._lib.bh_sampler_run(self._h, int(niter)))

    def run_all(self, chunk=2048):
        """iter_burnin + iter_main iterations (SingleChain.run_chain's loop)."""
        total = self.config.iter_burnin + self.config.iter_main
        done = int(self.state()["iiter"][0]) + self.config.iter_burnin if self._initialised else 0
        while done < total:
            n = min(chunk, total - done)
            self.run(n)
            done += n

    def state(self):
        B, L, T = self.nchains, self.maxlayers, self.ntargets
        s = dict(models=np.zeros((B, 2 * L)), k=np.zeros(B, np.int32), vpvs=np.zeros(B), noise=np.zeros((B, 2 * T)),
                 logL=np.zeros(B), misfits=np.zeros((B, T + 1)), propdist=np.zeros((B, 5)),
                 accepted=np.zeros((B, 5), np.int64), proposed=np.zeros((B, 5), np.int64),
                 iiter=np.zeros(B, np.int64), nstored=np.zeros(B, np.int32), overflow=np.zeros(1, np.int64))
        order = ("models", "k", "vpvs", "noise", "logL", "misfits", "propdist", "accepted", "proposed", "iiter",
                 "nstored", "overflow")
        _lib.check(self._lib.bh_sampler_get_state(self._h, *[s[n].ctypes.data for n in order]))
        # per chain: accepted models that found the chain arrays full, and the iteration of the first one
        s["overflow_count"] = np.zeros(B, np.int64)
        s["overflow_iter"] = np.zeros(B, np.int64)
        _lib.check(self._lib.bh_sampler_get_overflow(self._h, s["overflow_count"].ctypes.data,
                                                     s["overflow_iter"].ctypes.data))
        return s

    def set_state(self, models, k, vpvs, noise, logL=None, misfits=None, propdist=None, accepted=None,
                  proposed=None, iiter=None):
        def ptr(a, dt):
            if a is None:
                return None
            a = np.ascontiguousarray(a, dtype=dt)
            keep.append(a)
            return a.ctypes.data
        keep = []
        _lib.check(self._lib.bh_sampler_set_state(
            self._h, ptr(models, np.float64), ptr(k, np.int32), ptr(vpvs, np.float64), ptr(noise, np.float64),
            ptr(logL, np.float64), ptr(misfits, np.float64), ptr(propdist, np.float64), ptr(accepted, np.int64),
            ptr(proposed, np.int64), ptr(iiter, np.int64)))
        self._initialised = True

    def proposal(self):
        B, L, T = self.nchains, self.maxlayers, self.ntargets
        s = dict(models=np.zeros((B, 2 * L)), k=np.zeros(B, np.int32), vpvs=np.zeros(B), noise=np.zeros((B, 2 * T)),
                 valid=np.zeros(B, np.int32), modify=np.zeros(B, np.int32), dvs2=np.zeros(B), logL=np.zeros(B),
                 misfits=np.zeros((B, T + 1)))
        order = ("models", "k", "vpvs", "noise", "valid", "modify", "dvs2", "logL", "misfits")
        _lib.check(self._lib.bh_sampler_get_proposal(self._h, *[s[n].ctypes.data for n in order]))
        return s

    def force_draws(self, draws):
        if draws is None:
            _lib.check(self._lib.bh_sampler_set_forced_draws(self._h, None))
            return
        d = np.ascontiguousarray(draws, dtype=np.float64).reshape(self.nchains, 4)
        _lib.check(self._lib.bh_sampler_set_forced_draws(self._h, d.ctypes.data))

    def chain_arrays(self, chain0=0, nchain=None):
        """The device chain arrays (float32, NaN padded) of chains [chain0, chain0 + nchain)."""
        n = self.nchains - chain0 if nchain is None else int(nchain)
        S, L, T = self.nmodels, self.maxlayers, self.ntargets
        a = dict(models=np.empty((n, S, 2 * L), np.float32), misfits=np.empty((n, S, T + 1), np.float32),
                 likes=np.empty((n, S), np.float32), noise=np.empty((n, S, 2 * T), np.float32),
                 vpvs=np.empty((n, S), np.float32), iters=np.empty((n, S), np.int32))
        order = ("models", "misfits", "likes", "noise", "vpvs", "iters")
        _lib.check(self._lib.bh_sampler_get_chains(self._h, int(chain0), n, *[a[k].ctypes.data for k in order]))
        return a


# ---- what the reference does with a finished chain ------------------------------------------
def weighted_phase(arr, n, phase, finaliter):
    """SingleChain.get_weightedvalues for one chain (src/SingleChain.py:646-666): rows accepted
    in `phase` (1: iteration < 0, 2: >= 0) repeated by their dwell time.  Returns the row index
    of every weighted sample (index into the n stored rows), i.e. np.repeat(rows, weights)."""
    it = arr["iters"][:n].astype(np.int64)
    pind = np.where(it < 0)[0] if phase == 1 else np.where(it >= 0)[0]
    if pind.size == 0:
        return None
    weights = np.diff(np.concatenate((it[pind], [finaliter])))
    return np.repeat(pind, weights.astype(int))


def save_chain_files(arrays, nstored, chainidx, savepath, maxmodels, final_iter):
    """Per-chain result files of SingleChain.save_finalmodels (src/SingleChain.py:668-690):
    c%03d_p{1,2}{models,likes,misfits,noise,vpvs}.npy, weighted and thinned.
    arrays: chain_arrays() of ONE chain (leading axis removed).  Returns the number of main-phase
    rows saved."""
    names = ['models', 'likes', 'misfits', 'noise', 'vpvs']
    idx1 = weighted_phase(arrays, nstored, 1, min(0, final_iter))    # final_iter < 0: a chain that overflowed in burn-in
    idx2 = weighted_phase(arrays, nstored, 2, final_iter)
    if idx2 is None:
        raise ValueError("chain %d accepted no main-phase model" % chainidx)   # reference: AttributeError
    thinning = int(np.ceil(float(idx2.size) / float(maxmodels)))
    saved = 0
    for phase, idx in ((1, idx1), (2, idx2)):
        if idx is None:
            continue
        sel = idx[::thinning]
        for name in names:
            data = arrays[name][sel].astype(np.float64) if name in ('models', 'misfits', 'noise') \
                else arrays[name][sel]
            np.save(op.join(savepath, 'c%.3d_p%d%s' % (chainidx, phase, name)), data)
        if phase == 2:
            saved = sel.size
    return saved
