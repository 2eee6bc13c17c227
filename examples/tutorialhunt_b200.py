"""BayHunter's tutorial/tutorialhunt.py with the imports switched to bayhunter_b200.

    python examples/tutorialhunt_b200.py                      # one GPU
    torchrun --nproc-per-node 2 examples/tutorialhunt_b200.py # chains split over two GPUs

Same steps as the reference script: synthetic observed data for the st3 model (Rayleigh phase
dispersion + P receiver function) with correlated noise, priors / initparams from an .ini file,
`MCMC_Optimizer(...).mp_inversion()`, then the pooled posterior.  What differs: `nchains` is an
ensemble of hundreds of chains that advance in lock step on the GPU instead of a handful of
processes, and plotting is left to BayHunter's PlotFromStorage (the files written here are its
input format).
"""
import logging
import os
import os.path as op
import sys

import numpy as np

sys.path.insert(0, op.dirname(op.dirname(op.abspath(__file__))))
from bayhunter_b200 import Targets, utils, MCMC_Optimizer, SynthObs, chains   # noqa: E402

logging.basicConfig(format=' %(processName)-12s: %(levelname)-8s |  %(message)s', level=logging.INFO)
rank = int(os.environ.get("RANK", 0))
if "LOCAL_RANK" in os.environ:
    import torch
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))

here = op.dirname(op.abspath(__file__))
priors, initparams = utils.load_params(op.join(here, 'config.ini'))
if len(sys.argv) > 1:
    initparams['savepath'] = sys.argv[1]

# ------------------------------------------------------------ synthetic observed data (st3 model)
h, vs, vpvs = [5, 23, 8, 0], [2.7, 3.6, 3.8, 4.4], 1.73
xsw = np.linspace(1, 41, 21)
xrf = np.linspace(-5, 35, 201)
_ysw = SynthObs.return_swddata(h, vs, vpvs=vpvs, x=xsw)['rdispph'][1]
_yrf = SynthObs.return_rfdata(h, vs, vpvs=vpvs, x=xrf)['prf'][1]
noise = [0.0, 0.012, 0.98, 0.005]                     # corr1, sigma1, corr2, sigma2
ysw_err = SynthObs.compute_expnoise(_ysw, corr=noise[0], sigma=noise[1])
yrf_err = SynthObs.compute_gaussnoise(_yrf, corr=noise[2], sigma=noise[3])
ysw, yrf = _ysw + ysw_err, _yrf + yrf_err

# -------------------------------------------------------------------------------- targets
target1 = Targets.RayleighDispersionPhase(xsw, ysw, yerr=np.abs(ysw_err) + 1e-3)
target2 = Targets.PReceiverFunction(xrf, yrf)
target2.moddata.plugin.set_modelparams(gauss=1., water=0.01, p=6.4)
targets = Targets.JointTarget(targets=[target1, target2])

priors.update({'mohoest': (38, 4), 'rfnoise_corr': 0.98, 'swdnoise_corr': 0.})

# ---------------------------------------------------------------------------- inversion
optimizer = MCMC_Optimizer(targets, initparams=initparams, priors=priors, random_seed=7)
seconds = optimizer.mp_inversion()
st = optimizer.state
iters = initparams['iter_burnin'] + initparams['iter_main']
lo, hi = optimizer.chain_range
print("rank %d: chains %d..%d, %d iterations each in %.1f s = %.0f chain-iterations/s; median logL %.1f; "
      "laws %s" % (rank, lo, hi - 1, iters, seconds, (hi - lo) * iters / seconds, np.median(st["logL"]),
                   [t.covariance_law() if t.get_covariance else "-" for t in targets.targets]), flush=True)

if "RANK" in os.environ:
    import torch.distributed as dist
    dist.init_process_group("gloo")
    dist.barrier()                                     # every rank has written its chain files
if rank == 0:
    data = op.join(initparams['savepath'], 'data')
    post = chains.save_final_distribution(data, maxmodels=20000, dev=0.05)
    from bayhunter_b200 import Model
    z_probe = (2.0, 15.0, 33.0, 50.0)
    vs_at = []
    for m in post['models'][::max(1, len(post['models']) // 2000)]:
        vp_, vs_, h_ = Model.get_vp_vs_h(m, 1.73, None)
        top = np.concatenate(([0], np.cumsum(h_)[:-1]))
        vs_at.append([vs_[np.searchsorted(top, d, side='right') - 1] for d in z_probe])
    print("pooled %d models from %d chains (%d outlier chains by median likelihood, see outliers.dat); "
          "median vs at %s km: %s (truth 2.7, 3.6, 3.8, 4.4)"
          % (len(post['likes']), initparams['nchains'] - len(post['outliers']), len(post['outliers']),
             z_probe, np.round(np.median(vs_at, axis=0), 2)), flush=True)
