"""CPU tests of the device sampler's core (bayhunter_b200/csrc/sampler_core.cuh compiled for the
host, tests/host_sim) against steps RECORDED FROM THE REFERENCE's own SingleChain
(tests/golden/ref_sampler_steps.npz, made by tests/golden/make_sampler_fixtures.py), plus the
host-side pieces: configuration, initial draws, chain files."""
import ctypes
import os

import numpy as np
import pytest

from tests.conftest import ROOT

D = ctypes.POINTER(ctypes.c_double)


def _fixture(golden_dir):
    return np.load(os.path.join(golden_dir, "ref_sampler_steps.npz"))


def fixture_config(fx, name):
    """bh_sampler_config of a fixture setup (flattened priors/initparams in the npz)."""
    from bayhunter_b200 import _lib
    g = lambda k: fx["%s/cfg_%s" % (name, k)]
    c = _lib.BhSamplerConfig()
    c.layers_min, c.layers_max = int(g("layers")[0]), int(g("layers")[1])
    c.vs_min, c.vs_max = g("vs")
    c.z_min, c.z_max = g("z")
    v = g("vpvs")
    c.vpvs_fixed = int(v.size == 1)
    c.vpvs_min, c.vpvs_max = (v[0], v[0]) if v.size == 1 else (v[0], v[1])
    if g("mantle").size:
        c.has_mantle, c.mantle_vs, c.mantle_vpvs = 1, g("mantle")[0], g("mantle")[1]
    nf = g("noise_fixed")
    for i in range(2 * _lib.MAX_TARGETS):
        c.noise_fixed[i] = 1
    for i in range(nf.size):
        c.noise_fixed[i] = int(nf[i]); c.noise_min[i] = g("noise_lo")[i]; c.noise_max[i] = g("noise_hi")[i]
    c.thickmin = g("thickmin")[0]
    if g("lvz").size:
        c.has_lvz, c.lvz = 1, g("lvz")[0]
    if g("hvz").size:
        c.has_hvz, c.hvz = 1, g("hvz")[0]
    for i in range(5):
        c.propdist[i] = g("propdist0")[i]
    c.acceptance[0], c.acceptance[1] = g("acceptance")
    c.iter_burnin, c.iter_main = int(g("iters")[0]), int(g("iters")[1])
    c.max_accepted = 64
    c.seed = 1
    return c, nf.size // 2


def test_philox_known_answers(host_sim):
    """Philox4x32-10 known-answer vectors of the Random123 distribution (kat_vectors)."""
    def ph(ctr, key):
        c = (ctypes.c_uint32 * 4)(*ctr); k = (ctypes.c_uint32 * 2)(*key)
        host_sim.sampler_sim_philox(c, k)
        return list(c)
    assert ph([0] * 4, [0] * 2) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert ph([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert ph([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_draw_statistics(host_sim):
    host_sim.sampler_sim_draw.argtypes = [ctypes.c_ulonglong, ctypes.c_ulonglong, ctypes.c_longlong, D]
    out = np.zeros(4)
    n = 40000
    d = np.zeros((n, 4))
    for i in range(n):
        host_sim.sampler_sim_draw(12345, i % 200, -100 + i // 200, out.ctypes.data_as(D))
        d[i] = out
    for col in (0, 1, 3):
        assert 0 <= d[:, col].min() and d[:, col].max() < 1
        assert abs(d[:, col].mean() - 0.5) < 0.01 and abs(d[:, col].var() - 1 / 12) < 0.003
    assert abs(d[:, 2].mean()) < 0.02 and abs(d[:, 2].std() - 1) < 0.02
    assert abs(np.corrcoef(d.T) - np.eye(4)).max() < 0.03
    # counter based: same (seed, chain, iteration) -> same variates
    a = np.zeros(4); b = np.zeros(4)
    host_sim.sampler_sim_draw(7, 3, -5, a.ctypes.data_as(D)); host_sim.sampler_sim_draw(7, 3, -5, b.ctypes.data_as(D))
    assert np.array_equal(a, b)


@pytest.mark.parametrize("name", ["default", "constrained"])
def test_sampler_core_reproduces_reference_steps(name, host_sim, golden_dir):
    """Every recorded step of the reference chain: same modification, same proposal (bit for bit),
    same prior verdict, same log acceptance probability (given the reference's proposal
    likelihood), same decision, same proposal widths afterwards."""
    from bayhunter_b200 import _lib
    fx = _fixture(golden_dir)
    cfg, T = fixture_config(fx, name)
    L = cfg.layers_max + 1
    g = lambda k: fx["%s/%s" % (name, k)]
    host_sim.sampler_sim_alpha.restype = ctypes.c_double
    host_sim.sampler_sim_alpha.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, D, ctypes.c_double,
                                           ctypes.c_double, ctypes.c_double]
    host_sim.sampler_sim_propose.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, D, D, D,
                                             ctypes.POINTER(ctypes.c_int), D, D, ctypes.POINTER(ctypes.c_int), D, D]
    N = g("modify").size
    seen = np.zeros(6, int); n_adjust = n_invalid = n_acc = 0
    for i in range(N):
        model = g("b_model")[i].copy(); k = ctypes.c_int(int(g("b_k")[i]))
        vpvs = ctypes.c_double(g("b_vpvs")[i]); noise = g("b_noise")[i].copy()
        propdist = g("b_propdist")[i].copy(); draws = g("draws")[i].copy()
        modify = ctypes.c_int(-1); dvs2 = ctypes.c_double(0.0); rows = np.zeros(L * 4)
        iiter = int(g("b_iiter")[i])
        valid = host_sim.sampler_sim_propose(ctypes.byref(cfg), T, iiter, propdist.ctypes.data_as(D),
                                             draws.ctypes.data_as(D), model.ctypes.data_as(D), ctypes.byref(k),
                                             ctypes.byref(vpvs), noise.ctypes.data_as(D), ctypes.byref(modify),
                                             ctypes.byref(dvs2), rows.ctypes.data_as(D))
        assert modify.value == g("modify")[i], i
        assert valid == g("valid")[i], (i, modify.value)
        seen[modify.value] += 1
        accepted = np.array(g("b_accepted")[i]); proposed = np.array(g("b_proposed")[i])
        after_model, after_k = g("b_model")[i], g("b_k")[i]
        after_vpvs, after_noise = g("b_vpvs")[i], g("b_noise")[i]
        if valid:
            n = int(g("p_nlay")[i])
            r = rows.reshape(L, 4)[:n]
            assert k.value == n
            assert np.array_equal(r[:, 0], g("p_vs")[i, :n])                      # vs
            assert np.array_equal(r[:, 0] * r[:, 1], g("p_vp")[i, :n])            # vp = vs * vpvs (mantle rule)
            assert np.array_equal(r[:, 3], g("p_h")[i, :n])                       # h
            assert np.array_equal(r[:, 2], np.concatenate(([0.0], np.cumsum(g("p_h")[i, :n])[:-1])))
            assert np.array_equal(noise, g("p_noise")[i])
            alpha = host_sim.sampler_sim_alpha(ctypes.byref(cfg), T, modify.value, propdist.ctypes.data_as(D),
                                               dvs2.value, float(g("p_logL")[i]), float(g("b_logL")[i]))
            assert dvs2.value == g("dvs2")[i]
            ra = g("alpha")[i]
            assert alpha == ra or abs(alpha - ra) <= 1e-12 * max(1.0, abs(ra)), (i, alpha, ra)
            par = [0, 1, 2, 2, 3, 4][modify.value]
            proposed[par] += 1
            acc = bool(np.log(draws[3]) < alpha)
            assert acc == bool(g("accepted")[i]), i
            if acc:
                accepted[par] += 1
                n_acc += 1
                after_model, after_k, after_vpvs, after_noise = model, k.value, vpvs.value, noise
            if iiter % 1000 == 0 and np.all(proposed != 0):
                host_sim.sampler_sim_adjust(ctypes.byref(cfg), T, propdist.ctypes.data_as(D),
                                            accepted.ctypes.data_as(ctypes.POINTER(ctypes.c_longlong)),
                                            proposed.ctypes.data_as(ctypes.POINTER(ctypes.c_longlong)))
                n_adjust += 1
        else:
            n_invalid += 1
            assert not g("accepted")[i]
        ka = int(g("a_k")[i])
        assert int(after_k) == ka
        assert np.array_equal(np.asarray(after_model)[:ka], g("a_model")[i][:ka])
        assert np.array_equal(np.asarray(after_model)[L:L + ka], g("a_model")[i][L:L + ka])
        assert after_vpvs == g("a_vpvs")[i] and np.array_equal(after_noise, g("a_noise")[i])
        assert np.array_equal(propdist, g("a_propdist")[i]), (i, propdist, g("a_propdist")[i])
    expect_mods = [0, 1, 2, 3, 4, 5] if name == "default" else [0, 1, 2, 3, 4]
    assert all(seen[m] > 5 for m in expect_mods), seen
    assert n_invalid > 10 and n_acc > 50
    if name == "default":
        assert n_adjust >= 2
        assert not np.array_equal(g("a_propdist")[-1], g("cfg_propdist0"))


def test_initial_draws_follow_reference_sequence():
    """InitialStateDrawer consumes numpy RandomState exactly like SingleChain's init draws
    (vpvs, then vs / z nuclei with rejection, then the free noise parameters)."""
    from bayhunter_b200 import SingleChain as sc, Targets
    x = np.linspace(1, 40, 20)
    t = -5 + 0.2 * np.arange(201)
    jt = Targets.JointTarget([Targets.RayleighDispersionPhase(x, np.ones(20) * 3.5),
                              Targets.PReceiverFunction(t, np.zeros(201))])
    priors = dict(sc.DEFAULT_PRIORS); priors.update(layers=(2, 10), vs=(2, 5))
    ip = dict(sc.DEFAULT_INITPARAMS); ip.update(thickmin=0.5)
    m, vpvs, noise = sc.InitialStateDrawer(jt, priors, ip, 17).draw()
    rs = np.random.RandomState(17)
    assert vpvs == rs.uniform(low=1.5, high=2.1)
    while True:
        vs = np.sort(rs.uniform(low=2, high=5, size=3)); z = np.sort(rs.uniform(low=0, high=60, size=3))
        zd = (z[:-1] + z[1:]) / 2
        h = zd - np.concatenate(([0], zd[:-1]))
        if np.all(h >= 0.5):
            break
    assert np.array_equal(m, np.concatenate((vs, z)))
    # swd corr fixed at 0 -> not drawn; swd sigma, rf corr, rf sigma drawn in that order
    assert noise[0] == 0.0
    assert noise[1] == rs.uniform(low=1e-5, high=0.1)
    assert noise[2] == rs.uniform(low=0.35, high=0.75) and noise[3] == rs.uniform(low=1e-5, high=0.05)
    sc.set_target_covariance(jt, priors, None)
    assert [tt.covariance_law() for tt in jt.targets] == ["white", "exp"]
    c = sc.make_config(jt, priors, ip, seed=5)
    assert c.noise_fixed[0] == 1 and c.noise_fixed[1] == 0 and c.max_accepted == 6145     # one row per iteration fits
    assert sc.make_config(jt, priors, ip, seed=5, nchains=8).max_accepted == 6145
    big = dict(ip); big.update(iter_burnin=100000, iter_main=50000)
    cb = sc.make_config(jt, priors, big, seed=5, nchains=65536)
    assert cb.max_accepted * 65536 * 4 * (2 * 11 + 3 * 2 + 4) <= (32 << 30) and cb.max_accepted >= 64
    assert (c.layers_min, c.layers_max, c.vpvs_fixed, c.has_mantle) == (2, 10, 0, 0)


def test_chain_files_match_reference_weighting(tmp_path):
    """save_chain_files == np.repeat by dwell time, split at iteration 0, thinned (SingleChain.py:614-690)."""
    from bayhunter_b200 import SingleChain as sc
    S, L, T = 12, 3, 2
    iters = np.array([-50, -40, -12, -3, 0, 4, 5, 30, 0, 0, 0, 0], dtype=np.int32)
    n = 8
    rng = np.random.default_rng(0)
    arr = dict(models=rng.random((S, 2 * L)).astype(np.float32), misfits=rng.random((S, T + 1)).astype(np.float32),
               likes=rng.random(S).astype(np.float32), noise=rng.random((S, 2 * T)).astype(np.float32),
               vpvs=rng.random(S).astype(np.float32), iters=iters)
    saved = sc.save_chain_files(arr, n, 7, str(tmp_path), maxmodels=10, final_iter=40)
    w1 = np.diff(np.concatenate((iters[:4], [0])))          # 10, 28, 9, 3
    w2 = np.diff(np.concatenate((iters[4:8], [40])))        # 4, 1, 25, 10
    thin = int(np.ceil(w2.sum() / 10.))
    assert saved == len(np.repeat(np.arange(4), w2)[::thin])
    for phase, rows, w in ((1, np.arange(4), w1), (2, np.arange(4, 8), w2)):
        idx = np.repeat(rows, w)[::thin]
        for name in ("models", "likes", "misfits", "noise", "vpvs"):
            got = np.load(tmp_path / ("c007_p%d%s.npy" % (phase, name)))
            assert np.array_equal(got, arr[name][idx])


def test_save_final_distribution_matches_reference(tmp_path, golden_dir):
    """chains.save_final_distribution on the same per-chain files == what the reference's
    PlotFromStorage.save_final_distribution wrote (outlier chain, per-chain random subset, pooled
    arrays), fixture made by tests/golden/make_pooling_fixtures.py."""
    from bayhunter_b200 import chains
    fx = np.load(os.path.join(golden_dir, "ref_final_distribution.npz"))
    for c in range(6):
        for name in ("models", "likes", "misfits", "noise", "vpvs"):
            for phase in (1, 2):
                np.save(tmp_path / ("c%.3d_p%d%s" % (c, phase, name)), fx["in_c%d_%s" % (c, name)])
    out = chains.save_final_distribution(str(tmp_path), maxmodels=int(fx["maxmodels"][0]), dev=float(fx["dev"][0]),
                                         rstate=np.random.RandomState(333))
    assert list(out["outliers"]) == list(fx["outliers"])
    assert np.loadtxt(tmp_path / "outliers.dat", usecols=[0], dtype=int).tolist() == int(fx["outliers"][0])
    for name in ("models", "likes", "misfits", "noise", "vpvs"):
        got = np.load(tmp_path / ("c_%s.npy" % name))
        assert np.array_equal(got, fx["out_" + name], equal_nan=True), name


def test_synthobs_noise_matches_reference(golden_dir):
    """SynthObs.compute_expnoise / compute_gaussnoise reproduce the reference's draws from its
    module RandomState(333)."""
    from bayhunter_b200 import SynthObs as so_mod
    import bayhunter_b200.SynthObs as mod
    fx = np.load(os.path.join(golden_dir, "ref_synthobs_noise.npz"))
    mod.rstate = np.random.RandomState(333)
    for n in (20, 201):
        y = np.zeros(n)
        assert np.array_equal(so_mod.compute_expnoise(y, corr=0.6, sigma=0.02), fx["exp_%d" % n])
        assert np.array_equal(so_mod.compute_gaussnoise(y, corr=0.9, sigma=0.005), fx["gauss_%d" % n])


def test_load_params_reads_the_reference_ini_dialect(tmp_path):
    """utils.load_params on an .ini with the keys of BayHunter's tutorial/config.ini (values evaluated
    like src/utils.py:44-70: expressions, None, comma lists; station / savepath stay strings)."""
    from bayhunter_b200 import utils
    ini = tmp_path / "config.ini"
    ini.write_text("[modelpriors]\nvpvs = 1.4, 2.1\nlayers = 1, 20\nmohoest = None\nrfnoise_corr = 0.9\n"
                   "swdnoise_corr = 0.\nrfnoise_sigma = 1e-5, 0.05   # comment\n\n[initparams]\nnchains = 5\n"
                   "iter_burnin = (2048 * 16)\npropdist = 0.015, 0.015, 0.015, 0.005, 0.005\nacceptance = 40, 45\n"
                   "lvz = None\nrcond= 1e-5\nstation = 'test'\nsavepath = 'results'\n[datapaths]\nx = y\n")
    priors, ip = utils.load_params(str(ini))
    assert priors == dict(vpvs=[1.4, 2.1], layers=[1, 20], mohoest=None, rfnoise_corr=0.9, swdnoise_corr=0.0,
                          rfnoise_sigma=[1e-5, 0.05])
    assert ip == dict(nchains=5, iter_burnin=32768, propdist=[0.015, 0.015, 0.015, 0.005, 0.005], acceptance=[40, 45],
                      lvz=None, rcond=1e-5, station='test', savepath='results')
