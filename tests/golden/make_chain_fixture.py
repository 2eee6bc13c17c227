"""tests/golden/make_chain_fixture.py -- BASELINE config 1 as a golden vector: whole chains of the
REFERENCE's SingleChain (tutorial set-up: Rayleigh phase dispersion K = 21 + P receiver function
n = 201, 4 chains x 2048 iterations), driven by replayed random variates (see make_sampler_fixtures.py).

    python tests/golden/make_chain_fixture.py        (build container only)

Writes ref_chains_config1.npz: observed data, priors / initparams, per chain the initial state, the
variates of every iteration, and what the reference chain looked like afterwards (final model, vpvs,
noise, likelihood, proposal widths, counters, number of accepted models and their likelihoods).
The GPU test replays the same variates through bh_sampler_* and must end in the same state.
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
warnings.simplefilter("ignore")
import refshim  # noqa: E402
from make_sampler_fixtures import ReplayRNG, build_targets  # noqa: E402  (imports the reference through refshim)
from BayHunter.SingleChain import SingleChain  # noqa: E402

PRIORS = dict(vs=(2.0, 5.0), z=(0, 60), layers=(1, 20), vpvs=(1.4, 2.1), mantle=None, mohoest=None,
              rfnoise_corr=(0.35, 0.75), rfnoise_sigma=(1e-5, 0.05), swdnoise_corr=0., swdnoise_sigma=(1e-5, 0.05))
INIT = dict(iter_burnin=1536, iter_main=512, propdist=(0.015, 0.015, 0.015, 0.005, 0.005), acceptance=(40, 45),
            thickmin=0.1, lvz=None, hvz=None, rcond=1e-5)
NCHAINS, ITER = 4, 2048


def main():
    rng = np.random.default_rng(2024)
    joint, obs = build_targets(rng)
    T = len(joint.targets)
    maxlayers = PRIORS["layers"][1] + 1
    out = {}
    for ref, x, y in obs:
        out["obs_%s_x" % ref] = x
        out["obs_%s_y" % ref] = y
    out["refs"] = np.array([o[0] for o in obs])
    L2 = 2 * maxlayers

    def pad(m):
        k = m.size // 2
        v = np.zeros(L2)
        v[:k] = m[:k]; v[maxlayers:maxlayers + k] = m[k:]
        return v, k
    keys = ("init_model", "init_k", "init_vpvs", "init_noise", "draws", "fin_model", "fin_k", "fin_vpvs", "fin_noise",
            "fin_logL", "fin_propdist", "fin_accepted", "fin_proposed", "n_accepted", "acc_likes", "acc_iters")
    acc = {k: [] for k in keys}
    for c in range(NCHAINS):
        ip = dict(INIT); ip.update(nchains=1, station="fx", savepath="/tmp/bh_fx", maxmodels=50000)
        nmodels = ITER + 8
        f32 = np.float32
        nref = int(ITER * max(ip["acceptance"]) / 100.)          # what the reference's constructor expects
        shared = [np.full(n, np.nan, dtype=f32) for n in (nref * maxlayers * 2, nref * (T + 1), nref,
                                                         nref * T * 2, nref)]
        chain = SingleChain(targets=joint, chainidx=0, initparams=ip, modelpriors=dict(PRIORS),
                            sharedmodels=shared[0], sharedmisfits=shared[1], sharedlikes=shared[2],
                            sharednoise=shared[3], sharedvpvs=shared[4], random_seed=100 + c)
        # the reference sizes its arrays by the acceptance target (and raises IndexError beyond): give the
        # replay one row per iteration
        chain.nmodels = nmodels
        chain.chainmodels = np.full((nmodels, maxlayers * 2), np.nan, dtype=f32)
        chain.chainmisfits = np.full((nmodels, T + 1), np.nan, dtype=f32)
        chain.chainlikes = np.full(nmodels, np.nan, dtype=f32)
        chain.chainnoise = np.full((nmodels, T * 2), np.nan, dtype=f32)
        chain.chainvpvs = np.full(nmodels, np.nan, dtype=f32)
        chain.chainiter = np.ones(nmodels) * np.nan
        chain.n = 0
        chain.iiter = -chain.iter_phase1
        chain.append_currentmodel()
        chain.modelmods = ['vsmod', 'zvmod', 'birth', 'death']
        chain.noisemods = [] if len(chain.noiseinds) == 0 else ['noise']
        chain.vpvsmods = [] if type(chain.priors['vpvs']) == float else ['vpvs']
        chain.modifications = chain.modelmods + chain.noisemods + chain.vpvsmods
        chain.accepted = np.zeros(len(chain.propdist)); chain.proposed = np.zeros(len(chain.propdist))
        chain.tnull = 0.0
        m0, k0 = pad(np.array(chain.currentmodel))
        acc["init_model"].append(m0); acc["init_k"].append(k0); acc["init_vpvs"].append(float(chain.currentvpvs))
        acc["init_noise"].append(np.array(chain.currentnoise))
        replay = ReplayRNG(); chain.rstate = replay
        drng = np.random.default_rng(5000 + c)
        draws = np.zeros((ITER, 4))
        for it in range(ITER):
            draws[it] = (drng.random(), drng.random(), drng.standard_normal(), drng.random())
            replay.set(draws[it])
            chain.iterate()
        m1, k1 = pad(np.array(chain.currentmodel))
        acc["draws"].append(draws); acc["fin_model"].append(m1); acc["fin_k"].append(k1)
        acc["fin_vpvs"].append(float(chain.currentvpvs)); acc["fin_noise"].append(np.array(chain.currentnoise))
        acc["fin_logL"].append(float(chain.currentlikelihood)); acc["fin_propdist"].append(np.array(chain.propdist))
        acc["fin_accepted"].append(np.array(chain.accepted)); acc["fin_proposed"].append(np.array(chain.proposed))
        acc["n_accepted"].append(chain.n)
        likes = np.full(nmodels, np.nan); likes[:chain.n] = chain.chainlikes[:chain.n]
        iters = np.full(nmodels, -99999); iters[:chain.n] = chain.chainiter[:chain.n]
        acc["acc_likes"].append(likes); acc["acc_iters"].append(iters)
        print("chain %d: %d accepted of %d, final logL %.3f, layers %d" % (c, chain.n - 1, ITER, chain.currentlikelihood, k1 - 1))
    for k, v in acc.items():
        out[k] = np.array(v)
    out["iters"] = np.array([INIT["iter_burnin"], INIT["iter_main"]])
    np.savez_compressed(os.path.join(HERE, "ref_chains_config1.npz"), **out)
    print("ref_chains_config1.npz: %.0f KB" % (os.path.getsize(os.path.join(HERE, "ref_chains_config1.npz")) / 1e3))


if __name__ == "__main__":
    main()
