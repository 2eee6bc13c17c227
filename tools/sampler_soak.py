import sys, time; sys.path.insert(0,'.')
import numpy as np
from tests.test_gpu_sampler import _st3_targets
from bayhunter_b200 import SingleChain as sc
priors = dict(vs=(2, 5), z=(0, 60), layers=(1, 20), vpvs=(1.4, 2.1), mantle=None, mohoest=None,
              rfnoise_corr=(0.35, 0.75), rfnoise_sigma=(1e-5, 0.05), swdnoise_corr=0., swdnoise_sigma=(1e-5, 0.1))
ip = dict(iter_burnin=2000, iter_main=1000, thickmin=0.1, acceptance=(40, 45))
B = 8192
ens = sc.ChainEnsemble(_st3_targets(seed=3), priors, ip, nchains=B, seed=5, chain_seeds=np.arange(B) % 1000)
t0 = time.time(); ens.init(); t1 = time.time(); ens.run(3000); t2 = time.time()
st = ens.state()
print("init %.1f s (host draws), 3000 iterations %.1f s -> %.0f chain-iterations/s, %.2f ms/iteration" % (t1 - t0, t2 - t1, B * 3000 / (t2 - t1), (t2 - t1) / 3))
acc = st["accepted"].sum(1); prop = st["proposed"].sum(1)
print("iiter", st["iiter"].min(), st["iiter"].max(), "overflow", st["overflow"][0], "nstored == accepted + 1:", bool((st["nstored"] == acc + 1).all()))
print("acceptance rate median %.3f, evaluated fraction %.3f, layers median %d max %d, logL median %.1f, finite %s"
      % (np.median(acc / prop), prop.sum() / (B * 3000.), np.median(st["k"]) - 1, st["k"].max() - 1, np.median(st["logL"]), bool(np.isfinite(st["logL"]).all())))
print("propdist median", np.median(st["propdist"], axis=0))
