"""Developer probe: the tutorial inversion's lock-step iteration (rdispph + prf with the Gauss law, transdimensional)
at several ensemble sizes: ms per iteration and the per-kernel times of the last evaluation.

  python tools/tutorial_probe.py [nchains ...]"""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bayhunter_b200 import Targets, utils, SynthObs, SingleChain as sc

here = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples")
priors, initparams = utils.load_params(os.path.join(here, "config.ini"))
h, vs, vpvs = [5, 23, 8, 0], [2.7, 3.6, 3.8, 4.4], 1.73
xsw = np.linspace(1, 41, 21); xrf = np.linspace(-5, 35, 201)
ysw = SynthObs.return_swddata(h, vs, vpvs=vpvs, x=xsw)['rdispph'][1]
yrf = SynthObs.return_rfdata(h, vs, vpvs=vpvs, x=xrf)['prf'][1]
t1 = Targets.RayleighDispersionPhase(xsw, ysw, yerr=np.full(ysw.size, 0.012))
t2 = Targets.PReceiverFunction(xrf, yrf)
t2.moddata.plugin.set_modelparams(gauss=1., water=0.01, p=6.4)
jt = Targets.JointTarget(targets=[t1, t2])
priors.update({'mohoest': (38, 4), 'rfnoise_corr': 0.98, 'swdnoise_corr': 0.})
settings = dict(kv.split("=") for a in sys.argv[1:] if "=" in a for kv in a.split(","))
settings = {k: int(v) for k, v in settings.items()}
for n in [int(a) for a in sys.argv[1:] if "=" not in a] or [512]:
    ens = sc.ChainEnsemble(jt, priors, initparams, nchains=n, seed=7)
    ens.init()
    ens.engine.set(profile=0, **settings)
    ens.run(300)
    ens.state()
    t0 = time.perf_counter(); ens.run(1000); ens.state(); dt = time.perf_counter() - t0
    ens.engine.set(profile=1)
    ens.run(3); ens.state()
    print(json.dumps(dict(nchains=n, settings=settings, ms_per_iteration=dt, chain_iterations_per_s=n * 1000 / dt,
                          kernels={k: round(v, 4) for k, v in ens.engine.last_kernel_ms().items()},
                          mean_rows=float(ens.state()["k"].mean()) + 1)), flush=True)
    ens.close()
