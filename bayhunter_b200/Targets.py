"""Targets: the plugin surface of the hot path, GPU-backed.

Mirrors the public interface of BayHunter's src/Targets.py (ObservedData :16-30,
ModeledData :33-82, Valuation :85-183, SingleTarget :186-249, the six target
classes :252-297, JointTarget :300-347) so that SingleChain.iterate() and
MCMC_Optimizer can use these objects unchanged, but

* `JointTarget.evaluate(h, vp, vs, noise)` runs forward models AND likelihood in
  one call of the CUDA engine (batch of one), and
* `JointTarget.evaluate_batch(...)` evaluates thousands of chains per launch.

The covariance law of a target is whatever the chain bound to
`target.get_covariance` (src/SingleChain.py:159-205); the bound method's name
selects the device law.  The numpy bodies of `Valuation` exist because the chain
binds and BayWatch/SynthObs call them; `evaluate` never uses them.
"""
import logging

import numpy as np

from . import _lib
from .engine import Engine, TargetSpec
from .Models import pack_layers

logger = logging.getLogger()

RF_REFS = ("prf", "srf")
SWD_REFS = ("rdispph", "ldispph", "rdispgr", "ldispgr")


class ObservedData(object):
    """x (monotone increasing), y = y(x), optional yerr (NaN vector if unusable)."""

    def __init__(self, x, y, yerr=None):
        self.x = x
        self.y = y
        self.yerr = yerr
        if self.yerr is None or np.any(yerr <= 0.) or np.any(np.isnan(yerr)):
            self.yerr = np.ones(x.size) * np.nan


class ModeledData(object):
    """Holds the forward-modelling plugin and the latest synthetic (x, y)."""

    def __init__(self, obsx, ref):
        self.x = self.y = np.nan
        self.plugin, self.xlabel = None, "x"
        kind = "rf" if ref in RF_REFS else "swd" if ref in SWD_REFS else None
        if kind == "rf":
            from .rfmini_modrf import RFminiModRF as plugin_cls
        elif kind == "swd":
            from .surf96_modsw import SurfDisp as plugin_cls
        else:
            logger.info("No forward plugin is known for ref %r: attach one with "
                        "target.update_plugin(MyForwardClass())" % (ref,))
            return
        self.plugin = plugin_cls(obsx, ref)
        self.xlabel = {"rf": "Time in s", "swd": "Period in s"}[kind]

    def update(self, plugin):
        self.plugin = plugin

    def calc_synth(self, h, vp, vs, **kwargs):
        rho = kwargs.pop("rho")
        self.x, self.y = self.plugin.run_model(h, vp, vs, rho=rho, **kwargs)


class Valuation(object):
    """Covariance-law constructors + log-likelihood (host utilities).

    The chain binds one of the get_covariance_* methods per target; the device
    likelihood is selected from that binding (see `covariance_law`)."""

    def __init__(self):
        self.corr_inv = None
        self.logcorr_det = None
        self.misfit = None
        self.likelihood = None

    @staticmethod
    def get_rms(yobs, ymod):
        return np.sqrt(np.mean((ymod - yobs) ** 2))

    @staticmethod
    def get_covariance_nocorr(sigma, size, yerr=None, corr=0):
        return np.eye(size) / sigma ** 2, (2 * size) * np.log(sigma)

    @staticmethod
    def get_covariance_nocorr_scalederr(sigma, size, yerr, corr=0):
        scaled_err = yerr / yerr.min()          # not squared: reference quirk (Targets.py:125-127)
        return (np.eye(size) / (scaled_err * sigma ** 2),
                (2 * size) * np.log(sigma) + np.log(np.prod(scaled_err)))

    @staticmethod
    def get_corr_inv(corr, size):
        d = np.full(size, 1.0 + corr ** 2)
        d[0] = d[-1] = 1
        e = np.full(size - 1, -corr)
        return np.diag(d) + np.diag(e, k=1) + np.diag(e, k=-1)

    def get_covariance_exp(self, corr, sigma, size, yerr=None):
        c_inv = self.get_corr_inv(corr, size) / (sigma ** 2 * (1 - corr ** 2))
        return c_inv, (2 * size) * np.log(sigma) + (size - 1) * np.log(1 - corr ** 2)

    def init_covariance_gauss(self, corr, size, rcond=None):
        from .engine import gauss_corr_inverse
        self.corr_inv, self.logcorr_det = gauss_corr_inverse(corr, size, rcond)

    def get_covariance_gauss(self, sigma, size, yerr=None, corr=None):
        return self.corr_inv / sigma ** 2, (2 * size) * np.log(sigma) + self.logcorr_det

    @staticmethod
    def get_likelihood(yobs, ymod, c_inv, logc_det):
        ydiff = ymod - yobs
        madist = ydiff.dot(c_inv).dot(ydiff)
        return -0.5 * (yobs.size * np.log(2 * np.pi) + logc_det) - madist / 2.


_LAW_BY_METHOD = {"get_covariance_exp": "exp", "get_covariance_nocorr": "white",
                  "get_covariance_nocorr_scalederr": "white_scaled", "get_covariance_gauss": "gauss"}


class SingleTarget(object):
    """Observed + modelled data + valuation of one data set."""

    def __init__(self, x, y, ref, yerr=None):
        self.ref = ref
        self.obsdata = ObservedData(x=x, y=y, yerr=yerr)
        self.moddata = ModeledData(obsx=x, ref=ref)
        self.valuation = Valuation()
        self.get_covariance = None     # bound by the chain (SingleChain.py:175-205)
        logger.info("Initiated target: %s (ref: %s)" % (self.__class__.__name__, self.ref))

    def update_plugin(self, plugin):
        self.moddata.update(plugin)

    def covariance_law(self):
        """Name of the device law matching the bound get_covariance method."""
        name = getattr(self.get_covariance, "__name__", None)
        if name is None:
            return "exp"    # unbound: the law BayHunter uses for a free correlation
        try:
            return _LAW_BY_METHOD[name]
        except KeyError:
            raise ValueError("target %s: get_covariance is bound to %r, which has no device "
                             "equivalent" % (self.ref, name))

    def _moddata_valid(self):
        """The reference's acceptance test of a synthetic (src/Targets.py:204-214): an array of the
        observed length whose abscissa agrees in the (signed!) sum to 1e-5."""
        mx, my, ox = self.moddata.x, self.moddata.y, self.obsdata.x
        return (type(mx) == np.ndarray and len(mx) == len(ox) and bool(np.sum(ox - mx) <= 1e-5)
                and len(my) == len(self.obsdata.y))

    def calc_misfit(self):
        ok = self._moddata_valid()
        self.valuation.misfit = self.valuation.get_rms(self.obsdata.y, self.moddata.y) if ok else 1e15

    def calc_likelihood(self, c_inv, logc_det):
        ok = self._moddata_valid()
        self.valuation.likelihood = (self.valuation.get_likelihood(self.obsdata.y, self.moddata.y, c_inv, logc_det)
                                     if ok else -1e15)

    def to_spec(self, generic=False):
        """Engine description of this target (observed data, law, plugin parameters).  generic: the
        forward model stays with the plugin object (host side), the engine only evaluates the likelihood."""
        law = self.covariance_law()
        plugin = self.moddata.plugin
        params = {} if generic else dict(getattr(plugin, "modelparams", {}))
        params.pop("water", None)
        wtype = params.pop("wtype", None)
        if wtype is not None and {"P": "prf", "SV": "srf"}.get(wtype) != self.ref:
            raise ValueError("target %s: wtype %r contradicts the target reference" % (self.ref, wtype))
        kw = {}
        if law == "white_scaled":
            kw["yerr"] = self.obsdata.yerr
        if law == "gauss":
            kw["corr_inv"] = self.valuation.corr_inv
            kw["logcorr_det"] = self.valuation.logcorr_det
        y = self.obsdata.y if self.obsdata.y is not None else np.zeros(self.obsdata.x.size)
        return TargetSpec(self.ref, self.obsdata.x, y, cov=law, generic=generic, **kw, **params)


def _target_class(name, ref, noiseref):
    """Target class of BayHunter's name (src/Targets.py:252-297): fixes `ref` and the family of noise
    priors (`swd` / `rf`) its hyper-parameters are drawn from."""
    def __init__(self, x, y, yerr=None):
        SingleTarget.__init__(self, x, y, ref, yerr=yerr)
    cls = type(name, (SingleTarget,), {"__init__": __init__, "noiseref": noiseref, "__module__": __name__,
                                       "__doc__": "%s target (ref %r, noise priors %r)." % (name, ref, noiseref)})
    return cls


RayleighDispersionPhase = _target_class("RayleighDispersionPhase", "rdispph", "swd")
RayleighDispersionGroup = _target_class("RayleighDispersionGroup", "rdispgr", "swd")
LoveDispersionPhase = _target_class("LoveDispersionPhase", "ldispph", "swd")
LoveDispersionGroup = _target_class("LoveDispersionGroup", "ldispgr", "swd")
PReceiverFunction = _target_class("PReceiverFunction", "prf", "rf")
SReceiverFunction = _target_class("SReceiverFunction", "srf", "rf")


class JointTarget(object):
    """List of SingleTargets; joint log-likelihood of a model (or of a batch)."""

    def __init__(self, targets):
        self.targets = targets
        self.ntargets = len(targets)
        self._engine = None
        self._engine_key = None

    # pickling: engines hold device handles -> dropped, rebuilt lazily
    def __getstate__(self):
        d = dict(self.__dict__)
        d["_engine"] = None
        d["_engine_key"] = None
        return d

    def get_misfits(self):
        misfits = [target.valuation.misfit for target in self.targets]
        return np.concatenate((misfits, [np.sum(misfits)]))

    def _all_native(self):
        from .rfmini_modrf import RFminiModRF
        from .surf96_modsw import SurfDisp
        return all(isinstance(t.moddata.plugin, (RFminiModRF, SurfDisp)) for t in self.targets)

    @staticmethod
    def _spec_key(s):
        """Everything of a target the device constants are built from: a re-bound law, a new fixed Gauss
        correlation (R^-1), changed errors or abscissae must all lead to a new engine."""
        blob = lambda a: None if a is None else np.ascontiguousarray(a).tobytes()
        return (s.ref, s.generic, s.cov, s.n, tuple(sorted((k, str(v)) for k, v in s.params.items())),
                s.x.tobytes(), s.y.tobytes(), blob(s.yerr), blob(s.corr_inv), float(s.logcorr_det))

    def new_engine(self, max_batch, max_layers, generic=False):
        """A fresh Engine for the current laws / parameters, owned by the caller (a ChainEnsemble keeps the
        raw handle inside its device sampler, so it must not share the cached engine below)."""
        return Engine([t.to_spec(generic=generic) for t in self.targets], max_batch, max_layers)

    def engine(self, max_batch=1, max_layers=_lib.MAX_LAYERS, generic=False):
        """Cached engine for evaluate / evaluate_batch, bound to the current laws / parameters; rebuilt when
        any of them changes or a larger batch / deeper model arrives.  Never handed to a sampler."""
        specs = [t.to_spec(generic=generic) for t in self.targets]
        key = tuple(self._spec_key(s) for s in specs)
        if self._engine is None or self._engine_key != key or \
                self._engine.max_batch < max_batch or self._engine.max_layers < max_layers:
            if self._engine is not None:
                self._engine.close()
            self._engine = Engine(specs, max_batch, max_layers)
            self._engine_key = key
        return self._engine

    def evaluate(self, h, vp, vs, noise, **kwargs):
        """Single-model evaluation with BayHunter's semantics (src/Targets.py:314-347):
        leaves `proposallikelihood` (float) and `proposalmisfits` (length T+1).  Keyword arguments are
        forwarded to the plugins like the reference does (`rho`; `qp`, `qs` for rfmini)."""
        if not self._all_native() or any(k != "rho" for k in kwargs):
            return self._evaluate_through_plugins(h, vp, vs, noise, **kwargs)
        h = np.asarray(h, dtype=np.float64)
        rows = pack_layers(h, vp, vs)[None]
        nlay = np.array([h.size], dtype=np.int32)
        rho = kwargs.pop("rho", None)
        if rho is not None:
            rho = np.ascontiguousarray(rho, dtype=np.float64)[None]
        eng = self.engine(1, max(h.size, 2))
        logL, misfits, status, synth = eng.eval_host(
            rows, nlay, np.asarray(noise, dtype=np.float64)[None], rho=rho, want_synth=True)
        for target, ys in zip(self.targets, eng.split_synth(synth[0])):
            if status[0]:
                target.moddata.x, target.moddata.y = target.obsdata.x, ys.copy()
            else:
                target.moddata.x, target.moddata.y = np.nan, np.nan
        if not status[0]:
            self.proposallikelihood = -1e15
            self.proposalmisfits = [1e15] * (self.ntargets + 1)
            return
        for target, m in zip(self.targets, misfits[0]):
            target.valuation.misfit = m
        self.proposallikelihood = float(logL[0])
        self.proposalmisfits = misfits[0].copy()

    def _evaluate_through_plugins(self, h, vp, vs, noise, **kwargs):
        """The reference's own sequence (src/Targets.py:319-347) when the fused device evaluation does not
        apply: a user plugin (templates/myfwd.py contract) is attached to a target, or keyword arguments
        other than `rho` must reach the plugins (`qp`, `qs` arrays).  Every target's plugin models its data
        on the host side of the boundary (this package's plugins on the GPU through the shims, a user
        plugin wherever it likes); validity, misfits, covariance laws and the joint log-likelihood are
        evaluated on the device from those synthetics (bh_engine_loglik_host) -- there is no host
        likelihood path."""
        rho = kwargs.pop("rho", None)
        if rho is None:
            rho = np.asarray(vp) * 0.32 + 0.77           # Berteussen 1977 (src/Targets.py:319)
        tvalid = np.ones((1, self.ntargets), dtype=np.int32)
        synth = []
        for t, target in enumerate(self.targets):
            target.moddata.calc_synth(h, vp, vs, rho=rho, **kwargs)
            ok = target._moddata_valid()
            tvalid[0, t] = 1 if ok else 0
            synth.append(np.asarray(target.moddata.y, dtype=np.float64) if ok else np.zeros(target.obsdata.y.size))
        eng = self.engine(1, 2, generic=True)
        logL, misfits, status = eng.loglik_host(np.concatenate(synth)[None], tvalid,
                                                np.asarray(noise, dtype=np.float64)[None])
        if not status[0]:
            self.proposallikelihood = -1e15
            self.proposalmisfits = [1e15] * (self.ntargets + 1)
            return
        for target, m in zip(self.targets, misfits[0]):
            target.valuation.misfit = m
        self.proposallikelihood = float(logL[0])
        self.proposalmisfits = misfits[0].copy()

    def evaluate_batch(self, rows, nlay, noise, rho=None, want_synth=False):
        """Batched evaluation.  Accepts torch CUDA tensors (device path) or numpy
        arrays / CPU tensors (host path: copies happen inside the C call).
        rows [B,L,4] = (vs, vp/vs, z_top, h); nlay [B]; noise [B,2T].
        Returns (logL [B], misfits [B,T+1], status [B], synth | None)."""
        B, L = int(rows.shape[0]), int(rows.shape[1])
        eng = self.engine(B, L)
        if hasattr(rows, "is_cuda") and rows.is_cuda:
            return eng.eval(rows, nlay, noise, rho=rho, want_synth=want_synth)
        to_np = lambda a: a.numpy() if hasattr(a, "numpy") else a
        return eng.eval_host(to_np(rows), to_np(nlay), to_np(noise),
                             rho=None if rho is None else to_np(rho), want_synth=want_synth)
