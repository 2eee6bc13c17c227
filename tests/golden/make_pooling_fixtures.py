"""tests/golden/make_pooling_fixtures.py -- golden vectors for the host-side mirrors, produced by
RUNNING THE REFERENCE (build container only; see refshim.py):

  ref_final_distribution.npz  PlotFromStorage.save_final_distribution (src/Plotting.py:161-258,
                              get_outliers :113-154) on small synthetic per-chain files: the inputs
                              (per-chain arrays) and the pooled c_*.npy / outliers the reference wrote
  ref_synthobs_noise.npz      SynthObs.compute_expnoise / compute_gaussnoise (src/SynthObs.py:137-155)
                              from a freshly imported module (its RandomState(333))
"""
import os
import sys
import tempfile
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
warnings.simplefilter("ignore")
import refshim  # noqa: E402

BH = refshim.import_reference()
from BayHunter import Targets, utils  # noqa: E402
from BayHunter import Plotting, SynthObs  # noqa: E402


def make_pooling():
    rng = np.random.default_rng(99)
    tmp = tempfile.mkdtemp()
    data = os.path.join(tmp, "data")
    os.makedirs(data)
    x = np.linspace(1, 40, 20)
    t = -5 + 0.2 * np.arange(201)
    jt = Targets.JointTarget(targets=[Targets.RayleighDispersionPhase(x, np.ones(20) * 3.5),
                                      Targets.PReceiverFunction(t, np.zeros(201))])
    cfgfile = os.path.join(data, "fx_config.pkl")
    utils.save_config(jt, cfgfile, priors=dict(mantle=None, layers=(1, 5)),
                      initparams=dict(iter_burnin=100, iter_main=50, nchains=6))
    out = {}
    nch, maxl, T = 6, 6, 2
    for c in range(nch):
        n = int(rng.integers(25, 60))
        k = rng.integers(2, maxl + 1, n)
        models = np.full((n, 2 * maxl), np.nan)
        for i in range(n):
            models[i, :k[i]] = np.sort(rng.uniform(2, 5, k[i]))
            models[i, k[i]:2 * k[i]] = np.sort(rng.uniform(0, 60, k[i]))
        base = 400.0 if c not in (2, 4) else (330.0 if c == 2 else 395.0)     # chain 2 is an outlier
        arrs = dict(models=models, likes=(base + rng.normal(0, 3, n)).astype(np.float32),
                    misfits=rng.uniform(0.01, 0.1, (n, T + 1)), noise=rng.uniform(0, 0.1, (n, 2 * T)),
                    vpvs=rng.uniform(1.5, 2.0, n).astype(np.float32))
        for name, a in arrs.items():
            for phase in (1, 2):
                np.save(os.path.join(data, "c%.3d_p%d%s" % (c, phase, name)), a)
            out["in_c%d_%s" % (c, name)] = a
    pmod = sys.modules["BayHunter.Plotting"]
    pmod.rstate = np.random.RandomState(333)
    pfs = pmod.PlotFromStorage(cfgfile)
    pfs.save_final_distribution(maxmodels=150, dev=0.05)
    for name in ("models", "likes", "misfits", "noise", "vpvs"):
        out["out_" + name] = np.load(os.path.join(data, "c_%s.npy" % name))
    out["outliers"] = np.atleast_1d(np.loadtxt(os.path.join(data, "outliers.dat"), usecols=[0], dtype=int))
    out["maxmodels"] = np.array([150]); out["dev"] = np.array([0.05])
    np.savez_compressed(os.path.join(HERE, "ref_final_distribution.npz"), **out)
    print("ref_final_distribution.npz: pooled %d models, outliers %s" % (out["out_likes"].size, out["outliers"]))


def make_noise():
    mod = sys.modules["BayHunter.SynthObs"]
    mod.rstate = np.random.RandomState(333)
    cls = mod.SynthObs
    out = {}
    for n in (20, 201):
        y = np.zeros(n)
        out["exp_%d" % n] = cls.compute_expnoise(y, corr=0.6, sigma=0.02)
        out["gauss_%d" % n] = cls.compute_gaussnoise(y, corr=0.9, sigma=0.005)
    np.savez_compressed(os.path.join(HERE, "ref_synthobs_noise.npz"), **out)
    print("ref_synthobs_noise.npz")


if __name__ == "__main__":
    make_pooling()
    make_noise()
