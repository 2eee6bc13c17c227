"""SynthObs with BayHunter's interface (src/SynthObs.py:18-155): synthetic "observed" data
from the GPU forward plugins, and the reference's correlated-noise generators.

The noise generators are data preparation, not hot path: they are the reference's numpy
recipes (same module-level RandomState(333), same dense covariance, same
multivariate_normal call), so a script that seeds nothing gets the reference's noise.
`compute_explike` (a BayWatch display helper) is not provided: this package has no host
likelihood; evaluate the model through JointTarget.evaluate instead.
"""
import logging
import os

import numpy as np

from . import Targets

logger = logging.getLogger()

rstate = np.random.RandomState(333)


class SynthObs():
    @staticmethod
    def return_swddata(h, vs, vpvs=1.73, pars=dict(), x=None):
        """Dictionary ref -> [x, y] of the four dispersion curves (src/SynthObs.py:25-58)."""
        if x is None:
            x = np.linspace(1, 40, 20)
        h = np.array(h)
        vs = np.array(vs)
        mode = pars.get('mode', 1)
        vp = vs * vpvs
        rho = vp * 0.32 + 0.77
        data = {}
        for cls in (Targets.RayleighDispersionPhase, Targets.RayleighDispersionGroup,
                    Targets.LoveDispersionPhase, Targets.LoveDispersionGroup):
            target = cls(x=x, y=None)
            target.moddata.plugin.set_modelparams(mode=mode)
            xmod, ymod = target.moddata.plugin.run_model(h=h, vp=vp, vs=vs, rho=rho)
            data[target.ref] = np.array([xmod, ymod])
        logger.info('Compute SWD for %d periods, with model vp/vs %.2f.' % (x.size, vpvs))
        return data

    @staticmethod
    def return_rfdata(h, vs, vpvs=1.73, pars=dict(), x=None):
        """Dictionary ref -> [x, y] of the P and S receiver functions (src/SynthObs.py:60-101)."""
        if x is None:
            x = np.linspace(-5, 35, 201)
        h = np.array(h)
        vs = np.array(vs)
        gauss = pars.get('gauss', 1.0)
        water = pars.get('water', 0.001)
        p = pars.get('p', 6.4)
        nsv = pars.get('nsv', None)
        vp = vs * vpvs
        rho = vp * 0.32 + 0.77
        data = {}
        for cls in (Targets.PReceiverFunction, Targets.SReceiverFunction):
            target = cls(x=x, y=None)
            target.moddata.plugin.set_modelparams(gauss=gauss, water=water, p=p, nsv=nsv)
            xmod, ymod = target.moddata.plugin.run_model(h=h, vp=vp, vs=vs, rho=rho)
            data[target.ref] = np.array([xmod, ymod])
        logger.info('Compute RF with gauss: %.2f, waterlevel: %.4f, slowness: %.2f' % (gauss, water, p))
        return data

    @staticmethod
    def save_data(data, outfile=None):
        """ASCII files, one per reference (src/SynthObs.py:103-118)."""
        if outfile is None:
            outfile = 'syn_%s.dat'
        if '%s' not in outfile:
            name, ext = os.path.splitext(outfile)
            outfile = name + '_%s.' + ext
        for ref in data.keys():
            x, y = data[ref]
            with open(outfile % ref, 'w') as f:
                for i in range(len(x)):
                    f.write('%.4f\t%.4f\n' % (x[i], y[i]))

    @staticmethod
    def save_model(h, vs, vpvs=1.73, outfile=None):
        """ASCII model table (src/SynthObs.py:120-135)."""
        h = np.array(h)
        vs = np.array(vs)
        vp = vs * vpvs
        rho = vp * 0.32 + 0.77
        if outfile is None:
            outfile = 'syn_mod.dat'
        target = Targets.PReceiverFunction(x=np.arange(10), y=None)
        target.moddata.plugin.write_startmodel(h, vp, vs, rho, outfile)

    @staticmethod
    def compute_expnoise(data_obs, corr=0.85, sigma=0.0125):
        """Exponentially correlated noise (src/SynthObs.py:137-145)."""
        idx = np.fromfunction(lambda i, j: (abs((i + j) - 2 * i)), (data_obs.size, data_obs.size))
        Ce = sigma ** 2 * corr ** idx
        return rstate.multivariate_normal(np.zeros(data_obs.size), Ce)

    @staticmethod
    def compute_gaussnoise(data_obs, corr=0.85, sigma=0.0125):
        """Gaussian correlated noise (src/SynthObs.py:147-155)."""
        idx = np.fromfunction(lambda i, j: (abs((i + j) - 2 * i)), (data_obs.size, data_obs.size))
        Ce = sigma ** 2 * corr ** (idx ** 2)
        return rstate.multivariate_normal(np.zeros(data_obs.size), Ce)
