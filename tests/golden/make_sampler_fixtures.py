"""tests/golden/make_sampler_fixtures.py -- golden MCMC steps produced by RUNNING THE REFERENCE.

    python tests/golden/make_sampler_fixtures.py      (build container only)

The reference's own SingleChain (src/SingleChain.py, imported unmodified through
refshim.py) is driven iteration by iteration.  Its numpy RandomState is replaced
by a replay object that hands out, per iteration, the four variates
(u_mod, u_idx, gauss, u_acc) the device sampler consumes (sampler_core.cuh: Draw),
at the very calls where the reference draws:

    rstate.choice(<list of modification names>)  -> list[int(u_mod * n)]
    rstate.randint(lo, hi)                        -> lo + int(u_idx * (hi - lo))
    rstate.choice(<ndarray noiseinds>)            -> arr[int(u_idx * n)]
    rstate.uniform(low=zmin, high=zmax)           -> low + (high - low) * u_idx   (birth depth)
    rstate.normal(0, scale)                       -> scale * gauss
    rstate.uniform(0, 1)                          -> u_acc                        (acceptance)

Every recorded step holds the chain state before the step, the variates, what the
reference proposed (model, vpvs, noise, validity, the (h, vp, vs) it evaluated, the
likelihood/misfits it got), its log acceptance probability, its decision and the
state after -- everything produced by the reference's code.  Forward values come from
the reference's rfmini C++ and from oracle/surf96_oracle.c (see make_reference_fixtures.py).

Writes ref_sampler_steps.npz.
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
warnings.simplefilter("ignore")
import refshim  # noqa: E402

BH = refshim.import_reference()
from BayHunter import Targets  # noqa: E402
from BayHunter.SingleChain import SingleChain  # noqa: E402
from BayHunter import SingleChain as SC_module  # noqa: E402,F401

ST3 = dict(h=np.array([5., 23., 8., 0.]), vs=np.array([2.7, 3.6, 3.8, 4.4]), vpvs=1.73)
MODS = ["vsmod", "zvmod", "birth", "death", "noise", "vpvs"]


class ReplayRNG(object):
    """Stands in for numpy.random.RandomState inside SingleChain.iterate."""

    def __init__(self):
        self.d = None
        self.modify = None

    def set(self, draws):
        self.d = draws
        self.modify = None

    def choice(self, a):
        if isinstance(a, np.ndarray):                      # noiseinds
            return a[min(int(self.d[1] * a.size), a.size - 1)]
        a = list(a)
        self.modify = a[min(int(self.d[0] * len(a)), len(a) - 1)]
        return self.modify

    def randint(self, low, high=None):
        low, high = int(low), int(high)
        return low + min(int(self.d[1] * (high - low)), high - low - 1)

    def uniform(self, *args, **kw):
        if args:                                           # rstate.uniform(0, 1): acceptance
            assert tuple(args) == (0, 1)
            return self.d[3]
        low, high = kw["low"], kw["high"]                  # birth depth
        return low + (high - low) * self.d[1]

    def normal(self, loc, scale):
        assert loc == 0
        return scale * self.d[2]


def build_targets(rng):
    h, vs = ST3["h"], ST3["vs"]
    vp = vs * ST3["vpvs"]
    rho = vp * 0.32 + 0.77
    periods = np.linspace(1, 40, 20)
    x_rf = -5.0 + 0.2 * np.arange(201)
    targets, obs = [], []
    for ref, cls in (("rdispph", Targets.RayleighDispersionPhase), ("prf", Targets.PReceiverFunction)):
        x = x_rf if ref == "prf" else periods
        probe = cls(x, np.zeros(x.size))
        _, y = probe.moddata.plugin.run_model(h, vp, vs, rho)
        y = y + rng.normal(0, 0.005 if ref == "prf" else 0.012, y.size)
        targets.append(cls(x, y))
        obs.append((ref, x, y))
    return Targets.JointTarget(targets=targets), obs


SETUPS = [
    # name, priors, initparams
    ("default", dict(vs=(2.0, 5.0), z=(0, 60), layers=(1, 20), vpvs=(1.4, 2.1), mantle=None, mohoest=None,
                     rfnoise_corr=(0.35, 0.75), rfnoise_sigma=(1e-5, 0.05), swdnoise_corr=0.,
                     swdnoise_sigma=(1e-5, 0.1)),
     dict(iter_burnin=1500, iter_main=1000, propdist=(0.025, 0.025, 0.015, 0.005, 0.005),
          acceptance=(40, 45), thickmin=0.1, lvz=None, hvz=None, rcond=1e-5)),
    ("constrained", dict(vs=(2.0, 5.0), z=(0, 60), layers=(1, 8), vpvs=1.73, mantle=[4.2, 1.8], mohoest=None,
                         rfnoise_corr=(0.35, 0.75), rfnoise_sigma=(1e-5, 0.05), swdnoise_corr=0.,
                         swdnoise_sigma=(1e-5, 0.1)),
     dict(iter_burnin=1200, iter_main=800, propdist=(0.06, 0.4, 0.08, 0.004, 0.005),
          acceptance=(40, 45), thickmin=0.3, lvz=0.1, hvz=0.25, rcond=1e-5)),
]


def run_setup(name, priors, initparams, seed, keep_prob):
    rng = np.random.default_rng(seed)
    joint, obs = build_targets(rng)
    ip = dict(initparams)
    ip.update(nchains=1, station="fx", savepath="/tmp/bh_fx", maxmodels=50000)
    maxlayers = int(priors["layers"][1]) + 1
    T = len(joint.targets)
    iterations = ip["iter_burnin"] + ip["iter_main"]
    nmodels = int(iterations * max(ip["acceptance"]) / 100.)
    f32 = np.float32
    shared = [np.full(n, np.nan, dtype=f32) for n in (nmodels * maxlayers * 2, nmodels * (T + 1), nmodels,
                                                     nmodels * T * 2, nmodels)]
    chain = SingleChain(targets=joint, chainidx=0, initparams=ip, modelpriors=dict(priors),
                        sharedmodels=shared[0], sharedmisfits=shared[1], sharedlikes=shared[2],
                        sharednoise=shared[3], sharedvpvs=shared[4], random_seed=seed)
    # what run_chain sets up before its loop (SingleChain.py:591-603)
    chain.iiter = -chain.iter_phase1
    chain.modelmods = ['vsmod', 'zvmod', 'birth', 'death']
    chain.noisemods = [] if len(chain.noiseinds) == 0 else ['noise']
    chain.vpvsmods = [] if type(chain.priors['vpvs']) == float else ['vpvs']
    chain.modifications = chain.modelmods + chain.noisemods + chain.vpvsmods
    chain.accepted = np.zeros(len(chain.propdist))
    chain.proposed = np.zeros(len(chain.propdist))
    chain.tnull = 0.0
    init = dict(model=np.array(chain.currentmodel), vpvs=float(chain.currentvpvs), noise=np.array(chain.currentnoise),
                logL=float(chain.currentlikelihood), misfits=np.array(chain.currentmisfits))
    replay = ReplayRNG()
    chain.rstate = replay

    rec_eval = {}
    orig_eval = joint.evaluate

    def eval_spy(h, vp, vs, noise, **kw):
        orig_eval(h=h, vp=vp, vs=vs, noise=noise, **kw)
        rec_eval.update(h=np.array(h), vp=np.array(vp), vs=np.array(vs), noise=np.array(noise),
                        logL=float(joint.proposallikelihood), misfits=np.array(joint.proposalmisfits))
    joint.evaluate = eval_spy
    rec_alpha = {}
    orig_alpha = chain.get_acceptance_probability

    def alpha_spy(modify):
        a = orig_alpha(modify)
        rec_alpha.update(alpha=float(a), dvs2=float(getattr(chain, "dvs2", 0.0)) if modify in ("birth", "death") else 0.0)
        return a
    chain.get_acceptance_probability = alpha_spy

    steps = []
    draw_rng = np.random.default_rng(seed + 1000)
    while chain.iiter < chain.iter_phase2:
        d = np.array([draw_rng.random(), draw_rng.random(), draw_rng.standard_normal(), draw_rng.random()])
        replay.set(d)
        rec_eval.clear(); rec_alpha.clear()
        before = dict(model=np.array(chain.currentmodel), vpvs=float(chain.currentvpvs),
                      noise=np.array(chain.currentnoise), logL=float(chain.currentlikelihood),
                      misfits=np.array(chain.currentmisfits), propdist=np.array(chain.propdist),
                      accepted=np.array(chain.accepted), proposed=np.array(chain.proposed),
                      iiter=int(chain.iiter), n=int(chain.n))
        chain.iterate()
        valid = bool(rec_eval)
        accepted = chain.n > before["n"]
        keep = (draw_rng.random() < keep_prob or before["iiter"] % 1000 == 0
                or before["iiter"] < -chain.iter_phase1 + 40)
        if keep:
            steps.append(dict(before=before, draws=d, modify=MODS.index(replay.modify), valid=valid,
                              ev=dict(rec_eval), alpha=rec_alpha.get("alpha", np.nan), dvs2=rec_alpha.get("dvs2", 0.0),
                              accepted=accepted,
                              after=dict(model=np.array(chain.currentmodel), vpvs=float(chain.currentvpvs),
                                         noise=np.array(chain.currentnoise), logL=float(chain.currentlikelihood),
                                         propdist=np.array(chain.propdist), n=int(chain.n))))
    print("%s: %d iterations, kept %d steps, accepted models %d, final propdist %s" %
          (name, iterations, len(steps), chain.n, np.round(chain.propdist, 4)))
    return steps, obs, init, chain, maxlayers


def pack(steps, maxlayers, T):
    N = len(steps)
    L2 = 2 * maxlayers

    def model_pad(m):            # [vs(k) | z(k)] -> vs in the first half, z in the second (ABI layout)
        k = m.size // 2
        out = np.zeros(L2)
        out[:k] = m[:k]
        out[maxlayers:maxlayers + k] = m[k:]
        return out, k
    o = dict(
        b_model=np.zeros((N, L2)), b_k=np.zeros(N, np.int32), b_vpvs=np.zeros(N), b_noise=np.zeros((N, 2 * T)),
        b_logL=np.zeros(N), b_misfits=np.zeros((N, T + 1)), b_propdist=np.zeros((N, 5)),
        b_accepted=np.zeros((N, 5), np.int64), b_proposed=np.zeros((N, 5), np.int64), b_iiter=np.zeros(N, np.int64),
        draws=np.zeros((N, 4)), modify=np.zeros(N, np.int32), valid=np.zeros(N, np.int32),
        p_h=np.full((N, maxlayers + 1), np.nan), p_vp=np.full((N, maxlayers + 1), np.nan),
        p_vs=np.full((N, maxlayers + 1), np.nan), p_nlay=np.zeros(N, np.int32), p_noise=np.zeros((N, 2 * T)),
        p_logL=np.zeros(N), p_misfits=np.zeros((N, T + 1)), alpha=np.full(N, np.nan), dvs2=np.zeros(N),
        accepted=np.zeros(N, np.int32),
        a_model=np.zeros((N, L2)), a_k=np.zeros(N, np.int32), a_vpvs=np.zeros(N), a_noise=np.zeros((N, 2 * T)),
        a_logL=np.zeros(N), a_propdist=np.zeros((N, 5)))
    for i, s in enumerate(steps):
        b, a = s["before"], s["after"]
        o["b_model"][i], o["b_k"][i] = model_pad(b["model"])
        o["b_vpvs"][i] = b["vpvs"]; o["b_noise"][i] = b["noise"]; o["b_logL"][i] = b["logL"]
        o["b_misfits"][i] = b["misfits"]; o["b_propdist"][i] = b["propdist"]
        o["b_accepted"][i] = b["accepted"]; o["b_proposed"][i] = b["proposed"]; o["b_iiter"][i] = b["iiter"]
        o["draws"][i] = s["draws"]; o["modify"][i] = s["modify"]; o["valid"][i] = s["valid"]
        if s["valid"]:
            ev = s["ev"]
            n = ev["h"].size
            o["p_nlay"][i] = n
            o["p_h"][i, :n] = ev["h"]; o["p_vp"][i, :n] = ev["vp"]; o["p_vs"][i, :n] = ev["vs"]
            o["p_noise"][i] = ev["noise"]; o["p_logL"][i] = ev["logL"]; o["p_misfits"][i] = ev["misfits"]
            o["alpha"][i] = s["alpha"]; o["dvs2"][i] = s["dvs2"]
        o["accepted"][i] = s["accepted"]
        o["a_model"][i], o["a_k"][i] = model_pad(a["model"])
        o["a_vpvs"][i] = a["vpvs"]; o["a_noise"][i] = a["noise"]; o["a_logL"][i] = a["logL"]
        o["a_propdist"][i] = a["propdist"]
    return o


def main():
    out = {}
    for si, (name, priors, ip) in enumerate(SETUPS):
        steps, obs, init, chain, maxlayers = run_setup(name, priors, ip, seed=4242 + si, keep_prob=0.2)
        T = len(obs)
        for k, v in pack(steps, maxlayers, T).items():
            out["%s/%s" % (name, k)] = v
        for ref, x, y in obs:
            out["%s/obs_%s_x" % (name, ref)] = x
            out["%s/obs_%s_y" % (name, ref)] = y
        out["%s/refs" % name] = np.array([o[0] for o in obs])
        out["%s/laws" % name] = np.array([t.get_covariance.__name__ for t in chain.targets.targets])
        # configuration, flattened
        pr = chain.priors
        cfg = dict(layers=np.array(pr["layers"], float), vs=np.array(pr["vs"], float), z=np.array(pr["z"], float),
                   vpvs=np.atleast_1d(np.array(pr["vpvs"], float)),
                   mantle=np.array(pr["mantle"] if pr["mantle"] is not None else [], float),
                   thickmin=np.array([chain.thickmin]),
                   lvz=np.array([] if chain.lowvelperc is None else [chain.lowvelperc], float),
                   hvz=np.array([] if chain.highvelperc is None else [chain.highvelperc], float),
                   propdist0=np.array(ip["propdist"], float), acceptance=np.array(ip["acceptance"], float),
                   iters=np.array([ip["iter_burnin"], ip["iter_main"]]),
                   noise_fixed=np.array([not isinstance(p, (tuple, list, np.ndarray)) for p in chain.noisepriors]),
                   noise_lo=np.array([p[0] if isinstance(p, (tuple, list, np.ndarray)) else p for p in chain.noisepriors], float),
                   noise_hi=np.array([p[1] if isinstance(p, (tuple, list, np.ndarray)) else p for p in chain.noisepriors], float))
        for k, v in cfg.items():
            out["%s/cfg_%s" % (name, k)] = v
    out["setups"] = np.array([s[0] for s in SETUPS])
    np.savez_compressed(os.path.join(HERE, "ref_sampler_steps.npz"), **out)
    print("ref_sampler_steps.npz: %.1f KB" % (os.path.getsize(os.path.join(HERE, "ref_sampler_steps.npz")) / 1e3))


if __name__ == "__main__":
    main()
