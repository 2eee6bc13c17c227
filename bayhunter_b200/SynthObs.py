"""SynthObs with BayHunter's interface (src/SynthObs.py:18-155): synthetic "observed" data
from the GPU forward plugins, and the reference's correlated-noise generators.

compute_expnoise / compute_gaussnoise are the reference's numpy recipes (same module-level
RandomState(333), same dense covariance, same multivariate_normal call), so a script that seeds
nothing gets the reference's noise; device_noise draws any number of realisations of the same two
distributions on the GPU (batches of synthetic observations).
`compute_explike` (a BayWatch display helper) is not provided: this package has no host
likelihood; evaluate the model through JointTarget.evaluate instead.
"""
import logging
import os

import numpy as np

from . import Targets

logger = logging.getLogger()

rstate = np.random.RandomState(333)


def _forward(classes, x, h, vs, vpvs, **plugin_params):
    """ref -> array([x, y]) for each target class, from one layered model (rho by Berteussen)."""
    h, vs = np.array(h), np.array(vs)
    vp = vs * vpvs
    rho = vp * 0.32 + 0.77
    out = {}
    for cls in classes:
        target = cls(x=x, y=None)
        target.moddata.plugin.set_modelparams(**plugin_params)
        out[target.ref] = np.array(target.moddata.plugin.run_model(h=h, vp=vp, vs=vs, rho=rho))
    return out


def _correlated_noise(n, rmatrix, sigma):
    return rstate.multivariate_normal(np.zeros(n), sigma ** 2 * rmatrix)


def _lag_matrix(n):
    i = np.arange(n)
    return np.abs(i[:, None] - i[None, :]).astype(float)


class SynthObs():
    @staticmethod
    def return_swddata(h, vs, vpvs=1.73, pars=dict(), x=None):
        """The four dispersion curves of a model (src/SynthObs.py:25-58); pars: {'mode': n}."""
        x = np.linspace(1, 40, 20) if x is None else x
        data = _forward((Targets.RayleighDispersionPhase, Targets.RayleighDispersionGroup,
                         Targets.LoveDispersionPhase, Targets.LoveDispersionGroup), x, h, vs, vpvs,
                        mode=pars.get('mode', 1))
        logger.info('Compute SWD for %d periods, with model vp/vs %.2f.' % (x.size, vpvs))
        return data

    @staticmethod
    def return_rfdata(h, vs, vpvs=1.73, pars=dict(), x=None):
        """P and S receiver functions of a model (src/SynthObs.py:60-101); pars: gauss, water, p, nsv."""
        x = np.linspace(-5, 35, 201) if x is None else x
        kw = dict(gauss=pars.get('gauss', 1.0), water=pars.get('water', 0.001), p=pars.get('p', 6.4),
                  nsv=pars.get('nsv', None))
        data = _forward((Targets.PReceiverFunction, Targets.SReceiverFunction), x, h, vs, vpvs, **kw)
        logger.info('Compute RF with gauss: %.2f, waterlevel: %.4f, slowness: %.2f' % (kw['gauss'], kw['water'], kw['p']))
        return data

    @staticmethod
    def save_data(data, outfile=None):
        """One two-column ASCII file per reference, 4 decimals (src/SynthObs.py:103-118)."""
        outfile = 'syn_%s.dat' if outfile is None else outfile
        if '%s' not in outfile:
            stem, ext = os.path.splitext(outfile)
            outfile = stem + '_%s.' + ext
        for ref, (x, y) in data.items():
            np.savetxt(outfile % ref, np.column_stack((x, y)), fmt='%.4f', delimiter='\t')

    @staticmethod
    def save_model(h, vs, vpvs=1.73, outfile=None):
        """ASCII model table through the RF plugin's writer (src/SynthObs.py:120-135)."""
        vs = np.array(vs)
        vp = vs * vpvs
        writer = Targets.PReceiverFunction(x=np.arange(10), y=None).moddata.plugin
        writer.write_startmodel(np.array(h), vp, vs, vp * 0.32 + 0.77, 'syn_mod.dat' if outfile is None else outfile)

    @staticmethod
    def device_noise(law, size, corr=0.85, sigma=0.0125, nreal=1, seed=333):
        """`nreal` realisations [nreal, size] of the reference's correlated noise drawn on the GPU
        (bh_correlated_noise): law 'exp' (compute_expnoise) or 'gauss' (compute_gaussnoise).  Same
        distributions as the numpy recipes below, a different generator (Philox keyed by seed and
        realisation) -- use the recipes when the reference's RandomState(333) stream itself is wanted."""
        import ctypes
        from . import _lib
        lib = _lib.require_device()
        out = np.empty((int(nreal), int(size)))
        factor = None
        if law == 'gauss':
            # numpy.random.multivariate_normal: x = z . (sqrt(s)[:, None] * v) with (u, s, v) = svd(cov)
            _, s, v = np.linalg.svd(corr ** (_lag_matrix(size) ** 2))
            factor = np.ascontiguousarray(np.sqrt(s)[:, None] * v)
        elif law != 'exp':
            raise ValueError("law must be 'exp' or 'gauss'")
        _lib.check(lib.bh_correlated_noise(_lib.COV_GAUSS if law == 'gauss' else _lib.COV_EXP, int(size), int(nreal),
                                           float(corr), float(sigma), int(seed),
                                           None if factor is None else factor.ctypes.data_as(_lib.c_double_p),
                                           out.ctypes.data_as(_lib.c_double_p)))
        return out

    @staticmethod
    def compute_expnoise(data_obs, corr=0.85, sigma=0.0125):
        """Exponentially correlated noise, R_ij = corr^|i-j| (src/SynthObs.py:137-145)."""
        return _correlated_noise(data_obs.size, corr ** _lag_matrix(data_obs.size), sigma)

    @staticmethod
    def compute_gaussnoise(data_obs, corr=0.85, sigma=0.0125):
        """Gaussian correlated noise, R_ij = corr^(|i-j|^2) (src/SynthObs.py:147-155)."""
        return _correlated_noise(data_obs.size, corr ** (_lag_matrix(data_obs.size) ** 2), sigma)
