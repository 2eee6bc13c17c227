"""Worker of tests/test_gpu_round2.py::test_sampler_graph_replay_changes_nothing: a short transdimensional
lock-step run, its final state digested.  BH_SAMPLER_GRAPH (read once per process by bh_sampler_run) selects
the plain enqueue or the CUDA-graph replay."""
import hashlib
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bayhunter_b200 import Targets as T_, SingleChain as sc, synthetic   # noqa: E402
from oracle import joint_oracle as jo                                     # noqa: E402

st3 = synthetic.ST3
h, vs = st3["h"], st3["vs"]
vp = vs * st3["vpvs"]
rho = vp * 0.32 + 0.77
per = np.linspace(1, 40, 12)
xrf = -5.0 + 0.2 * np.arange(101)
jt = T_.JointTarget([T_.RayleighDispersionPhase(per, jo.surfdisp(h, vp, vs, rho, "rdispph", per)[1]),
                     T_.RayleighDispersionGroup(per, jo.surfdisp(h, vp, vs, rho, "rdispgr", per)[1]),
                     T_.PReceiverFunction(xrf, jo.recfunc(h, vp, vs, rho, xrf)[1])])
priors = dict(vs=(2, 5), z=(0, 60), layers=(1, 15), vpvs=(1.4, 2.1), swdnoise_corr=0., swdnoise_sigma=(1e-5, 0.05),
              rfnoise_corr=0.9, rfnoise_sigma=(1e-5, 0.05))
ens = sc.ChainEnsemble(jt, priors, dict(iter_burnin=300, iter_main=300, thickmin=0.1, acceptance=(40, 45), rcond=1e-5),
                       nchains=96, seed=3, max_accepted=400)
ens.init()
ens.run(150)          # one plain iteration + a replayed chunk (+ a second capture) when graphs are on
ens.run(450)
st = ens.state()
dig = hashlib.sha256()
for k in ("models", "k", "vpvs", "noise", "logL", "misfits", "propdist", "accepted", "proposed", "iiter", "nstored"):
    dig.update(np.ascontiguousarray(st[k]).tobytes())
print("DIGEST", dig.hexdigest(), int(st["iiter"][0]), int(st["accepted"].sum()), float(np.median(st["logL"])))
ens.close()
