"""GPU tests of the device sampler (bh_sampler_* through the C ABI)."""
import os
import pickle

import numpy as np
import pytest

from tests.test_sampler import _fixture, fixture_config

pytestmark = pytest.mark.gpu

LAWS = {"get_covariance_exp": "exp", "get_covariance_nocorr": "white",
        "get_covariance_nocorr_scalederr": "white_scaled", "get_covariance_gauss": "gauss"}


def _fixture_targets(fx, name):
    from bayhunter_b200 import Targets
    cls = {"rdispph": Targets.RayleighDispersionPhase, "prf": Targets.PReceiverFunction}
    ts = []
    for ref in fx["%s/refs" % name]:
        ref = str(ref)
        ts.append(cls[ref](fx["%s/obs_%s_x" % (name, ref)], fx["%s/obs_%s_y" % (name, ref)]))
    return Targets.JointTarget(ts)


def _fixture_dicts(fx, name):
    g = lambda k: fx["%s/cfg_%s" % (name, k)]
    v = g("vpvs")
    priors = dict(layers=tuple(int(x) for x in g("layers")), vs=tuple(g("vs")), z=tuple(g("z")),
                  vpvs=float(v[0]) if v.size == 1 else tuple(v), mantle=list(g("mantle")) if g("mantle").size else None,
                  mohoest=None)
    nf, lo, hi = g("noise_fixed"), g("noise_lo"), g("noise_hi")
    pri = [float(lo[i]) if nf[i] else (float(lo[i]), float(hi[i])) for i in range(nf.size)]
    priors.update(swdnoise_corr=pri[0], swdnoise_sigma=pri[1], rfnoise_corr=pri[2], rfnoise_sigma=pri[3])
    ip = dict(iter_burnin=int(g("iters")[0]), iter_main=int(g("iters")[1]), propdist=tuple(g("propdist0")),
              acceptance=tuple(g("acceptance")), thickmin=float(g("thickmin")[0]),
              lvz=float(g("lvz")[0]) if g("lvz").size else None, hvz=float(g("hvz")[0]) if g("hvz").size else None,
              rcond=1e-5)
    return priors, ip


@pytest.mark.parametrize("name", ["default", "constrained"])
def test_kernels_replay_reference_steps(name, golden_dir):
    """Every recorded step of the REFERENCE chain becomes one chain of a batch: its state before
    the step is loaded, the recorded variates are forced, one lock-step iteration runs on the GPU
    (propose kernel -> engine -> accept kernel).  Proposal and prior verdict must be identical;
    the proposal's log-likelihood within 1e-6 relative (forward-model tolerance); the decision
    identical wherever it is not within that tolerance of the threshold; the state after the step
    (model, vpvs, noise, proposal widths, counters) the reference's."""
    from bayhunter_b200 import SingleChain as sc
    fx = _fixture(golden_dir)
    g = lambda k: fx["%s/%s" % (name, k)]
    jt = _fixture_targets(fx, name)
    priors, ip = _fixture_dicts(fx, name)
    N = g("modify").size
    ens = sc.ChainEnsemble(jt, priors, ip, nchains=N, seed=3, max_accepted=4)
    assert [t.covariance_law() for t in jt.targets] == [LAWS[str(x)] for x in fx["%s/laws" % name]]
    cfg_ref, _ = fixture_config(fx, name)
    for f in ("layers_min", "layers_max", "vs_min", "vs_max", "z_min", "z_max", "vpvs_fixed", "vpvs_min", "vpvs_max",
              "has_mantle", "mantle_vs", "mantle_vpvs", "thickmin", "has_lvz", "has_hvz", "lvz", "hvz",
              "iter_burnin", "iter_main"):
        assert getattr(ens.config, f) == getattr(cfg_ref, f), f
    L = ens.maxlayers
    ens.set_state(g("b_model"), g("b_k"), g("b_vpvs"), g("b_noise"), logL=g("b_logL"), misfits=g("b_misfits"),
                  propdist=g("b_propdist"), accepted=g("b_accepted"), proposed=g("b_proposed"), iiter=g("b_iiter"))
    ens.force_draws(g("draws"))
    ens.run(1)
    prop, st = ens.proposal(), ens.state()
    assert np.array_equal(prop["modify"], g("modify"))
    assert np.array_equal(prop["valid"], g("valid"))
    v = g("valid") == 1
    assert np.array_equal(prop["k"][v], g("p_nlay")[v])
    assert np.array_equal(prop["noise"][v], g("p_noise")[v])
    assert np.array_equal(prop["dvs2"][v], g("dvs2")[v])
    for i in np.where(v)[0]:
        n = int(g("p_nlay")[i])
        assert np.array_equal(prop["models"][i, :n], g("p_vs")[i, :n]), i
    rel = np.abs(prop["logL"][v] - g("p_logL")[v]) / np.maximum(1.0, np.abs(g("p_logL")[v]))
    assert rel.max() <= 1e-6, rel.max()
    assert np.allclose(prop["misfits"][v], g("p_misfits")[v], rtol=1e-6, atol=1e-12)
    # decisions: identical unless log(u) is within the likelihood tolerance of alpha
    margin = np.abs(np.log(g("draws")[:, 3]) - g("alpha"))
    tol = 2e-6 * np.maximum(1.0, np.abs(g("p_logL")))
    clear = v & (margin > tol)
    took = st["iiter"] * 0 + (st["nstored"] > 0)
    assert clear.sum() >= 0.98 * v.sum()
    assert np.array_equal(took[clear] == 1, g("accepted")[clear] == 1)
    assert not took[~v].any()
    assert np.array_equal(st["iiter"], g("b_iiter") + 1)
    same = clear | ~v
    for i in np.where(same)[0]:
        ka = int(g("a_k")[i])
        assert st["k"][i] == ka
        assert np.array_equal(st["models"][i, :ka], g("a_model")[i, :ka]), i
        assert np.array_equal(st["models"][i, L:L + ka], g("a_model")[i, L:L + ka]), i
    assert np.array_equal(st["vpvs"][same], g("a_vpvs")[same])
    assert np.array_equal(st["noise"][same], g("a_noise")[same])
    assert np.array_equal(st["propdist"][same], g("a_propdist")[same])
    acc = g("accepted")[same] == 1
    assert np.allclose(st["logL"][same][acc], g("a_logL")[same][acc], rtol=1e-6)
    # accepted rows landed in the chain arrays in the reference's layout
    arr = ens.chain_arrays()
    i = int(np.where(same & (g("accepted") == 1))[0][0])
    ka = int(g("a_k")[i])
    row = arr["models"][i, 0]
    assert np.array_equal(row[:ka], g("a_model")[i, :ka].astype(np.float32))
    assert np.array_equal(row[ka:2 * ka], g("a_model")[i, L:L + ka].astype(np.float32))
    assert np.isnan(row[2 * ka:]).all() and arr["iters"][i, 0] == g("b_iiter")[i]
    assert np.isnan(arr["likes"][i, 1:]).all()


def _st3_targets(seed=0, rf=True):
    from bayhunter_b200 import SurfDisp, RFminiModRF, Targets
    rng = np.random.default_rng(seed)
    h = np.array([5., 23., 8., 0.]); vs = np.array([2.7, 3.6, 3.8, 4.4]); vp = vs * 1.73
    rho = vp * 0.32 + 0.77
    x = np.linspace(1, 40, 21)
    _, y = SurfDisp(x, "rdispph").run_model(h, vp, vs, rho)
    ts = [Targets.RayleighDispersionPhase(x, y + rng.normal(0, 0.012, x.size))]
    if rf:
        t = -5 + 0.2 * np.arange(201)
        _, r = RFminiModRF(t, "prf").run_model(h, vp, vs, rho)
        ts.append(Targets.PReceiverFunction(t, r + rng.normal(0, 0.005, t.size)))
    return Targets.JointTarget(ts)


PRIORS = dict(vs=(2, 5), z=(0, 60), layers=(1, 12), vpvs=(1.4, 2.1), mantle=None, mohoest=None,
              rfnoise_corr=(0.35, 0.75), rfnoise_sigma=(1e-5, 0.05), swdnoise_corr=0., swdnoise_sigma=(1e-5, 0.1))


def test_trajectories_do_not_depend_on_the_batch():
    """Counter-based variates keyed by the GLOBAL chain index: running chains 0..47 in one batch or
    as two batches (0..23, 24..47 -- what two GPUs would do) gives identical chains."""
    from bayhunter_b200 import SingleChain as sc
    ip = dict(iter_burnin=300, iter_main=100, thickmin=0.1, acceptance=(40, 45))
    seeds = np.arange(100, 148)
    a = sc.ChainEnsemble(_st3_targets(), PRIORS, ip, nchains=48, seed=11, chain_seeds=seeds)
    a.run(120)
    sa, ca = a.state(), a.chain_arrays()
    parts = []
    for lo in (0, 24):
        e = sc.ChainEnsemble(_st3_targets(), PRIORS, ip, nchains=24, first_chain=lo, seed=11, chain_seeds=seeds[lo:lo + 24])
        e.run(120)
        parts.append((e.state(), e.chain_arrays()))
    for key in ("models", "k", "vpvs", "noise", "logL", "propdist", "accepted", "proposed", "nstored"):
        assert np.array_equal(sa[key], np.concatenate([p[0][key] for p in parts])), key
    for key in ("models", "likes", "iters"):
        assert np.array_equal(ca[key], np.concatenate([p[1][key] for p in parts]), equal_nan=True), key
    assert sa["proposed"].sum() > 48 * 60 and sa["accepted"].sum() > 48 * 10
    assert (sa["iiter"] == -300 + 120).all()


def test_inversion_converges_and_writes_reference_files(tmp_path):
    """MCMC_Optimizer drop-in: a short joint inversion of synthetic st3 data on the GPU.  The chains
    must climb in likelihood, keep BayHunter's acceptance control in range, recover the truth model
    within posterior spread, and leave the reference's files."""
    from bayhunter_b200 import Model
    from bayhunter_b200.mcmcOptimizer import MCMC_Optimizer
    jt = _st3_targets(seed=5)
    ip = dict(nchains=96, iter_burnin=3000, iter_main=1500, thickmin=0.1, acceptance=(40, 45), station="st3",
              savepath=str(tmp_path), maxmodels=500)
    opt = MCMC_Optimizer(jt, initparams=ip, priors=PRIORS, random_seed=1)
    first = None
    opt.ensemble.init()
    first = opt.ensemble.state()["logL"].copy()
    opt.mp_inversion()
    st = opt.state
    assert (st["iiter"] == 1500).all() and int(st["overflow"][0]) == 0
    assert np.median(st["logL"]) > np.median(first) + 100
    rate = st["accepted"].sum(axis=1) / np.maximum(1, st["proposed"].sum(axis=1))
    assert 0.2 < np.median(rate) < 0.6, np.median(rate)
    assert (st["propdist"] != np.array(ip.get("propdist", (0.025, 0.025, 0.015, 0.005, 0.005)))).any()
    data = os.path.join(str(tmp_path), "data")
    cfg = pickle.load(open(os.path.join(data, "st3_config.pkl"), "rb"))
    assert cfg["targetrefs"] == ["rdispph", "prf"] and cfg["initparams"]["nchains"] == 96
    best = np.argsort(st["logL"])[-48:]            # the converged half (reference: outlier chains are dropped)
    vs_at = []
    for c in best:
        m = np.load(os.path.join(data, "c%.3d_p2models.npy" % c))
        l = np.load(os.path.join(data, "c%.3d_p2likes.npy" % c))
        v = np.load(os.path.join(data, "c%.3d_p2vpvs.npy" % c))
        n = np.load(os.path.join(data, "c%.3d_p2noise.npy" % c))
        mf = np.load(os.path.join(data, "c%.3d_p2misfits.npy" % c))
        assert m.shape == (l.size, 2 * 13) and v.shape == l.shape and n.shape == (l.size, 4) and mf.shape == (l.size, 3)
        assert 0 < l.size <= 500
        for row in m[:: max(1, l.size // 20)]:
            vp, vs, h = Model.get_vp_vs_h(row, 1.73, None)
            z = np.concatenate(([0], np.cumsum(h)[:-1]))
            vs_at.append([vs[np.searchsorted(z, d, side="right") - 1] for d in (2.0, 15.0, 50.0)])
    vs_at = np.array(vs_at)
    med = np.median(vs_at, axis=0)
    assert abs(med[0] - 2.7) < 0.35 and abs(med[1] - 3.6) < 0.25 and abs(med[2] - 4.4) < 0.35, med


def test_whole_chains_of_the_reference_are_reproduced(golden_dir):
    """BASELINE config 1 (tutorial set-up: Rayleigh phase K = 21 + P-RF n = 201, 4 chains x 2048
    iterations) as a golden vector: tests/golden/make_chain_fixture.py ran the REFERENCE's SingleChain
    with replayed variates; the same variates through the device sampler must walk the same chains --
    same number of accepted models at the same iterations, same final model / vpvs / noise bit for bit,
    likelihoods within the forward-model tolerance, same proposal widths and counters."""
    from bayhunter_b200 import SingleChain as sc, Targets
    fx = np.load(os.path.join(golden_dir, "ref_chains_config1.npz"))
    cls = {"rdispph": Targets.RayleighDispersionPhase, "prf": Targets.PReceiverFunction}
    jt = Targets.JointTarget([cls[str(r)](fx["obs_%s_x" % r], fx["obs_%s_y" % r]) for r in fx["refs"]])
    priors = dict(vs=(2.0, 5.0), z=(0, 60), layers=(1, 20), vpvs=(1.4, 2.1), mantle=None, mohoest=None,
                  rfnoise_corr=(0.35, 0.75), rfnoise_sigma=(1e-5, 0.05), swdnoise_corr=0., swdnoise_sigma=(1e-5, 0.05))
    ip = dict(iter_burnin=int(fx["iters"][0]), iter_main=int(fx["iters"][1]), propdist=(0.015, 0.015, 0.015, 0.005, 0.005),
              acceptance=(40, 45), thickmin=0.1, lvz=None, hvz=None, rcond=1e-5)
    C, ITER = fx["draws"].shape[0], fx["draws"].shape[1]
    ens = sc.ChainEnsemble(jt, priors, ip, nchains=C, seed=1, max_accepted=ITER + 8)
    L = ens.maxlayers
    models = [np.concatenate((fx["init_model"][c, :fx["init_k"][c]], fx["init_model"][c, L:L + fx["init_k"][c]]))
              for c in range(C)]
    ens.init(models, fx["init_vpvs"], fx["init_noise"])
    for it in range(ITER):
        ens.force_draws(fx["draws"][:, it, :])
        ens.run(1)
    st = ens.state()
    arr = ens.chain_arrays()
    assert np.array_equal(st["nstored"], fx["n_accepted"])
    assert np.array_equal(st["k"], fx["fin_k"])
    for c in range(C):
        k = int(fx["fin_k"][c])
        assert np.array_equal(st["models"][c, :k], fx["fin_model"][c, :k]), c
        assert np.array_equal(st["models"][c, L:L + k], fx["fin_model"][c, L:L + k]), c
        n = int(fx["n_accepted"][c])
        assert np.array_equal(arr["iters"][c, :n], fx["acc_iters"][c, :n].astype(np.int32)), c
        rel = np.abs(arr["likes"][c, :n] - fx["acc_likes"][c, :n]) / np.maximum(1.0, np.abs(fx["acc_likes"][c, :n]))
        assert rel.max() <= 2e-6, (c, rel.max())                                  # float32 chain arrays
    assert np.array_equal(st["vpvs"], fx["fin_vpvs"]) and np.array_equal(st["noise"], fx["fin_noise"])
    assert np.array_equal(st["propdist"], fx["fin_propdist"])
    assert np.array_equal(st["accepted"], fx["fin_accepted"].astype(np.int64))
    assert np.array_equal(st["proposed"], fx["fin_proposed"].astype(np.int64))
    assert np.allclose(st["logL"], fx["fin_logL"], rtol=1e-6)
    assert (st["iiter"] == int(fx["iters"][1])).all()
